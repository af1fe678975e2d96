from .multi_output_unet import MultiOutputUnet  # noqa: F401
from .predict import Predict  # noqa: F401
