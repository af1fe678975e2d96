"""Minimal TIFF reader / writer for the call sites of the prediction path
(unet/predict.py:64-65, siam_unet/predict.py:79-81,102,110-115,123, utils/utils.py:21-22).

``tifffile`` is used when it is installed; otherwise this module reads and writes uncompressed, single-sample
(grayscale) classic TIFF / BigTIFF stacks of uint8/uint16/uint32/int*/float16/float32/float64 pages stored in strips.
"""
import struct

import numpy as np

try:  # pragma: no cover - depends on the environment
    import tifffile as _tifffile
except Exception:  # noqa: BLE001
    _tifffile = None

_TYPE_SIZES = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}
_TYPE_FMT = {1: 'B', 2: 'c', 3: 'H', 4: 'I', 6: 'b', 8: 'h', 9: 'i', 11: 'f', 12: 'd', 16: 'Q', 17: 'q', 18: 'Q'}


def _dtype_of(bits, sample_format, bo):
    kind = {1: 'u', 2: 'i', 3: 'f'}.get(sample_format, 'u')
    return np.dtype(f'{bo}{kind}{bits // 8}')


class _Page:
    def __init__(self, shape, dtype, offsets, counts):
        self.shape, self.dtype, self._offsets, self._counts = shape, dtype, offsets, counts


class TiffFile:
    """Parses the IFD chain; ``pages[i].shape`` and ``asarray(key)`` are what the Predict classes need."""

    def __init__(self, path):
        self.path = path
        with open(path, 'rb') as f:
            data = f.read(16)
            bo = {b'II': '<', b'MM': '>'}.get(data[:2])
            if bo is None:
                raise ValueError(f'{path}: not a TIFF file')
            magic = struct.unpack(bo + 'H', data[2:4])[0]
            self._big = magic == 43
            if magic not in (42, 43):
                raise ValueError(f'{path}: bad TIFF magic {magic}')
            off = struct.unpack(bo + 'Q', data[8:16])[0] if self._big else struct.unpack(bo + 'I', data[4:8])[0]
            self.pages = []
            while off:
                off = self._read_ifd(f, off, bo)

    def _read_ifd(self, f, off, bo):
        f.seek(off)
        big = self._big
        n = struct.unpack(bo + ('Q' if big else 'H'), f.read(8 if big else 2))[0]
        esz = 20 if big else 12
        raw = f.read(n * esz)
        nxt = struct.unpack(bo + ('Q' if big else 'I'), f.read(8 if big else 4))[0]
        tags = {}
        for i in range(n):
            e = raw[i * esz:(i + 1) * esz]
            tag, typ = struct.unpack(bo + 'HH', e[:4])
            cnt = struct.unpack(bo + ('Q' if big else 'I'), e[4:12] if big else e[4:8])[0]
            val = e[12:20] if big else e[8:12]
            size = _TYPE_SIZES.get(typ, 1) * cnt
            if size > len(val):
                pos = struct.unpack(bo + ('Q' if big else 'I'), val)[0]
                here = f.tell()
                f.seek(pos)
                val = f.read(size)
                f.seek(here)
            if typ in _TYPE_FMT and typ != 2:
                tags[tag] = struct.unpack(bo + _TYPE_FMT[typ] * cnt, val[:size])
        width, height = tags[256][0], tags[257][0]
        bits = tags.get(258, (1,))[0]
        if tags.get(259, (1,))[0] != 1:
            raise NotImplementedError(f'{self.path}: compressed TIFF pages are not supported by the built-in reader')
        if tags.get(277, (1,))[0] != 1:
            raise NotImplementedError(f'{self.path}: only single-sample (grayscale) pages are supported')
        if 273 not in tags:
            raise NotImplementedError(f'{self.path}: tiled TIFF pages are not supported by the built-in reader')
        dtype = _dtype_of(bits, tags.get(339, (1,))[0], bo)
        counts = tags.get(279) or (width * height * dtype.itemsize,)
        self.pages.append(_Page((height, width), dtype, tags[273], counts))
        return nxt

    def asarray(self, key=None):
        idx = range(len(self.pages)) if key is None else ([key] if np.isscalar(key) else list(key))
        out = []
        with open(self.path, 'rb') as f:
            for i in idx:
                p = self.pages[i]
                buf = bytearray()
                for o, c in zip(p._offsets, p._counts):
                    f.seek(o)
                    buf += f.read(c)
                out.append(np.frombuffer(bytes(buf), dtype=p.dtype).reshape(p.shape).astype(p.dtype.newbyteorder('=')))
        if key is not None and np.isscalar(key):
            return out[0]
        return out[0] if len(out) == 1 else np.stack(out)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class TiffWriter:
    """Appends uncompressed pages; switches to BigTIFF when asked."""

    def __init__(self, path, bigtiff=False):
        self.path, self._big = path, bool(bigtiff)
        self._f = open(path, 'wb')
        self._f.write(b'II' + (struct.pack('<HHHQ', 43, 8, 0, 0) if self._big else struct.pack('<HI', 42, 0)))
        self._link_pos = 8 if self._big else 4   # where the offset of the next IFD has to be patched in

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._f:
            self._f.close()
            self._f = None

    def write(self, arr, contiguous=True, **_ignored):
        arr = np.asarray(arr)
        pages = arr.reshape((-1,) + arr.shape[-2:]) if arr.ndim > 2 else arr[None]
        for page in pages:
            self._write_page(np.ascontiguousarray(page))

    def _write_page(self, page):
        f, big = self._f, self._big
        dt = page.dtype.newbyteorder('<') if page.dtype.byteorder == '>' else page.dtype
        page = page.astype(dt, copy=False)
        fmt = {'u': 1, 'i': 2, 'f': 3, 'b': 1}[dt.kind]
        if f.tell() % 2:
            f.write(b'\0')
        data_off = f.tell()
        f.write(page.tobytes())
        if f.tell() % 2:
            f.write(b'\0')
        ifd_off = f.tell()
        h, w = page.shape
        ent = [(256, 4, w), (257, 4, h), (258, 3, dt.itemsize * 8), (259, 3, 1), (262, 3, 1),
               (273, 16 if big else 4, data_off), (277, 3, 1), (278, 4, h), (279, 16 if big else 4, page.nbytes),
               (339, 3, fmt)]
        if big:
            f.write(struct.pack('<Q', len(ent)))
            for tag, typ, val in ent:
                f.write(struct.pack('<HHQQ', tag, typ, 1, val))
            nxt = f.tell()
            f.write(struct.pack('<Q', 0))
        else:
            if ifd_off >= 2 ** 32 - page.nbytes:
                raise ValueError('file exceeds 4 GiB: open the TiffWriter with bigtiff=True')
            f.write(struct.pack('<H', len(ent)))
            for tag, typ, val in ent:
                f.write(struct.pack('<HHI', tag, typ, 1) + (struct.pack('<HH', val, 0) if typ == 3 else struct.pack('<I', val)))
            nxt = f.tell()
            f.write(struct.pack('<I', 0))
        end = f.tell()
        f.seek(self._link_pos)
        f.write(struct.pack('<Q' if big else '<I', ifd_off))
        f.seek(end)
        self._link_pos = nxt


def imread(path, key=None):
    if _tifffile is not None:
        return _tifffile.imread(path, key=key)
    return TiffFile(path).asarray(key)


def imwrite(path, data, **kwargs):
    if _tifffile is not None:
        return _tifffile.imwrite(path, data, **kwargs)
    if kwargs.get('compression'):
        raise NotImplementedError('the built-in TIFF writer does not compress; install tifffile')
    data = np.asarray(data)
    with TiffWriter(path, bigtiff=data.nbytes > 2 ** 32 - 2 ** 25) as tw:
        tw.write(data)


def page_count_and_shape(path):
    if _tifffile is not None:
        with _tifffile.TiffFile(path) as t:
            return len(t.pages), tuple(t.pages[0].shape)
    t = TiffFile(path)
    return len(t.pages), tuple(t.pages[0].shape)
