from .siam_unet import Siam_UNet  # noqa: F401
from .predict import Predict  # noqa: F401
