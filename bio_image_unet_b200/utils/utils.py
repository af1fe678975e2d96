"""Device pick and result writer of the prediction path (utils/utils.py:8-22,56-78 in the reference)."""
import torch
from torch import nn

from .. import tiff


def get_device(print_device=False):
    """'auto' device of the Predict classes. The reference returns cuda:0 whenever torch was *built* with CUDA
    (utils/utils.py:63); this engine is CUDA-only, so a missing GPU is an error instead of a silent CPU run."""
    if not torch.cuda.is_available():
        raise RuntimeError('bio_image_unet_b200: no CUDA device available (the engine has no CPU fallback)')
    device = torch.device('cuda:0')
    if print_device:
        print(f'Using device: {device}')
    return device


def save_as_tif(imgs, filename, normalize=False):
    """float16 TIFF; `normalize` is accepted and ignored exactly like utils/utils.py:8-22."""
    tiff.imwrite(filename, imgs.astype('float16'))


def init_weights(m):
    if isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, nonlinearity='leaky_relu')
