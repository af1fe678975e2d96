from .multi_output_unet3d import MultiOutputUnet3D  # noqa: F401
from .predict import Predict, Session  # noqa: F401
