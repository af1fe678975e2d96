from .unet import Unet  # noqa: F401
from .predict import Predict, Session  # noqa: F401
