"""ctypes front-end of ``skimage_port.c`` (CPU oracle — test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libskimage_port.so")
_SRC = os.path.join(_HERE, "skimage_port.c")


def build(force=False):
    """Compile the C restatement with gcc (called by ``__graft_entry__.build``)."""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        tmp = _SO + ".%d.tmp" % os.getpid()
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", tmp, _SRC])
        os.replace(tmp, _SO)
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        i64 = ctypes.c_int64
        vp = ctypes.c_void_p
        lib.sk_label.restype = i64
        lib.sk_label.argtypes = [vp, i64, i64, i64, ctypes.c_int, vp]
        lib.sk_watershed.restype = None
        lib.sk_watershed.argtypes = [vp, vp, vp, i64, i64, vp]
        lib.sk_reconstruction_erosion.restype = None
        lib.sk_reconstruction_erosion.argtypes = [vp, vp, i64, i64, vp]
        lib.tiseg_align_foreground.restype = None
        lib.tiseg_align_foreground.argtypes = [vp, vp, i64, i64, ctypes.c_int]
        _lib = lib
    return _lib


def label(img, background=0, connectivity=None, return_num=False):
    """``skimage.measure.label`` for 2-D integer / bool images (int64 output, ids 1..K in raster
    order of first pixel; ``connectivity=None`` means full, i.e. 8-neighbourhood)."""
    a = np.ascontiguousarray(np.asarray(img).astype(np.int64))
    assert a.ndim == 2
    conn = 2 if connectivity is None else int(connectivity)
    out = np.empty_like(a)
    k = _load().sk_label(a.ctypes.data, a.shape[0], a.shape[1], int(background), conn, out.ctypes.data)
    return (out, int(k)) if return_num else out


def watershed(image, markers, mask=None):
    """``skimage.segmentation.watershed(image, markers, mask=mask)`` (connectivity 1,
    compactness 0, no watershed line).  int32 output like the library."""
    im = np.ascontiguousarray(np.asarray(image, dtype=np.float64))
    mk = np.ascontiguousarray(np.asarray(markers).astype(np.int32))
    ms = np.ones(im.shape, np.uint8) if mask is None else np.ascontiguousarray((np.asarray(mask) != 0).astype(np.uint8))
    out = np.empty(im.shape, np.int32)
    _load().sk_watershed(im.ctypes.data, mk.ctypes.data, ms.ctypes.data, im.shape[0], im.shape[1], out.ctypes.data)
    return out


def reconstruction_erosion(seed, mask):
    """``skimage.morphology.reconstruction(seed, mask, method='erosion')`` (float64 output)."""
    s = np.ascontiguousarray(np.asarray(seed, dtype=np.float64))
    m = np.ascontiguousarray(np.asarray(mask, dtype=np.float64))
    out = np.empty_like(s)
    _load().sk_reconstruction_erosion(s.ctypes.data, m.ctypes.data, s.shape[0], s.shape[1], out.ctypes.data)
    return out


def align_foreground(pred, foreground, time):
    """``tiseg/models/utils/postprocess.py:123-155``; mutates and returns ``pred`` (int64)."""
    assert pred.dtype == np.int64 and pred.flags.c_contiguous
    fg = np.ascontiguousarray((np.asarray(foreground) > 0).astype(np.uint8))
    _load().tiseg_align_foreground(pred.ctypes.data, fg.ctypes.data, pred.shape[0], pred.shape[1], int(time))
    return pred
