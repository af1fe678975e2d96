from .unet import Unet  # noqa: F401
from .attention_unet import AttentionUnet, AttentionBlock  # noqa: F401
from .unet_v0 import Unet_v0  # noqa: F401
from .predict import Predict, Session  # noqa: F401
