"""Import the UNMODIFIED reference package: from /root/reference in the authoring container, else from the
git-ignored install baseline/_ref that __graft_entry__.build() makes (that copy travels to the GPU box, where it is
the CPU arm of bench.py: `--impl reference` and `cpu_baseline`, kind "reference").

The reference needs tifffile, skimage, albumentations, matplotlib and napari at import time; none is installed
here. tifffile gets an in-memory stand-in (the Predict classes read/write through it); the other four are only
touched by training / GUI code and get attribute-swallowing stubs. Test infrastructure: never imported by the
product package; only tests/golden/make_golden.py and bench.py's CPU legs use it.
"""
import os
import sys
import types

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pick_root():
    for cand in (os.environ.get('BIU_REFERENCE_ROOT'), '/root/reference', os.path.join(_REPO, 'baseline', '_ref')):
        if cand and os.path.isdir(os.path.join(cand, 'bio_image_unet')):
            return cand
    return os.environ.get('BIU_REFERENCE_ROOT', '/root/reference')


REFERENCE_ROOT = _pick_root()

TIFF_STORE = {}   # filename -> ndarray (what the reference "wrote" / will "read")


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + '.' + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + '()')


class _Page:
    def __init__(self, arr):
        self.shape = arr.shape


class _TiffFile:
    def __init__(self, name):
        arr = TIFF_STORE[name]
        self.pages = [_Page(a) for a in (arr if arr.ndim == 3 else arr[None])]


class _TiffWriter:
    def __init__(self, name, bigtiff=False):
        self.name = name
        self.frames = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        TIFF_STORE[self.name] = np.stack(self.frames) if len(self.frames) > 1 else self.frames[0]
        return False

    def write(self, arr, contiguous=True):
        self.frames.append(np.array(arr))


def _imread(name, key=None):
    arr = TIFF_STORE[name]
    if key is None:
        return np.array(arr)
    return np.array(arr[key] if arr.ndim == 3 else arr)


def _imwrite(name, arr, **kw):
    TIFF_STORE[name] = np.array(arr)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'bio_image_unet'))


def import_reference():
    """Returns the imported ``bio_image_unet`` reference package (with stubs installed)."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    if 'bio_image_unet' in sys.modules and getattr(sys.modules['bio_image_unet'], '__biu_ref__', False):
        return sys.modules['bio_image_unet']
    tf = types.ModuleType('tifffile')
    tf.imread, tf.imwrite, tf.TiffFile, tf.TiffWriter = _imread, _imwrite, _TiffFile, _TiffWriter
    tf2 = types.ModuleType('tifffile.tifffile')
    tf2.TiffFile = _TiffFile
    tf.tifffile = tf2
    sys.modules['tifffile'] = tf
    sys.modules['tifffile.tifffile'] = tf2
    for name in ('skimage', 'skimage.morphology', 'skimage.draw', 'skimage.measure', 'skimage.filters',
                 'skimage.transform', 'skimage.io', 'albumentations', 'matplotlib', 'matplotlib.pyplot',
                 'napari', 'qtpy', 'qtpy.QtWidgets', 'qtpy.QtCore'):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import bio_image_unet  # noqa: E402
    bio_image_unet.__biu_ref__ = True
    return bio_image_unet
