from .siam_unet import Siam_UNet  # noqa: F401
from .predict import Predict, Session  # noqa: F401
