from .utils import get_device, save_as_tif, init_weights  # noqa: F401
