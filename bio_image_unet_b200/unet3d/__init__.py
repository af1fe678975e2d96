from .unet3d import UNet3D  # noqa: F401
from .predict import Predict, Session  # noqa: F401
