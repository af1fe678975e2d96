from .multi_output_nested_unet import MultiOutputNestedUNet, MultiOutputNestedUNet_3Levels  # noqa: F401
from .multi_output_unet import MultiOutputUnet  # noqa: F401
from .predict import Predict  # noqa: F401
