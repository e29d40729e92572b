"""ctypes binding of libyolo1_b200.so -- the C ABI declared in include/yolo1_b200.h.

The library is the product: there is no CPU or eager-PyTorch fallback.  If the shared object is missing
(or fails to load) every entry point raises, loudly, with the build command.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libyolo1_b200.so")
ABI_VERSION = 1

DTYPE_F32, DTYPE_BF16 = 0, 1
COORD_REFERENCE, COORD_PAPER = 0, 1

_c = ctypes
_i64p = _c.POINTER(_c.c_int64)
_vp = _c.c_void_p

# name -> (restype, argtypes); mirrors include/yolo1_b200.h one to one (tests/test_abi.py checks this list
# against the header and against the symbols the .so exports)
SIGNATURES = {
    "yolo1_abi_version": (_c.c_int, []),
    "yolo1_error_string": (_c.c_char_p, [_c.c_int]),
    "yolo1_loss_workspace_bytes": (_c.c_size_t, [_c.c_int64, _c.c_int, _c.c_int, _c.c_int]),
    "yolo1_loss_fwd_bwd": (_c.c_int, [_vp, _i64p, _c.c_int, _vp, _i64p, _vp, _i64p, _vp,
                                      _c.c_int64, _c.c_int, _c.c_int, _c.c_int,
                                      _c.c_float, _c.c_float, _c.c_float, _c.c_int,
                                      _vp, _c.c_size_t, _vp]),
    "yolo1_loss_fwd_bwd_ex": (_c.c_int, [_vp, _i64p, _c.c_int, _vp, _i64p, _vp, _i64p, _vp,
                                         _c.c_int64, _c.c_int, _c.c_int, _c.c_int,
                                         _c.c_float, _c.c_float, _c.c_float, _c.c_int,
                                         _vp, _c.c_size_t, _c.c_int, _vp]),
    "yolo1_loss_fwd_bwd_logits": (_c.c_int, [_vp, _i64p, _c.c_int, _vp, _i64p, _vp, _i64p, _vp,
                                             _c.c_int64, _c.c_int, _c.c_int, _c.c_int,
                                             _c.c_float, _c.c_float, _c.c_float, _c.c_int,
                                             _vp, _c.c_size_t, _vp]),
    "yolo1_loss_objects_workspace_bytes": (_c.c_size_t, [_c.c_int64, _c.c_int, _c.c_int, _c.c_int]),
    "yolo1_loss_fwd_bwd_objects": (_c.c_int, [_vp, _i64p, _c.c_int, _c.c_int, _vp, _vp, _vp, _vp, _i64p, _vp,
                                              _c.c_int64, _c.c_int, _c.c_int, _c.c_int,
                                              _c.c_float, _c.c_float, _c.c_float, _c.c_int,
                                              _vp, _c.c_size_t, _vp, _vp]),
    "yolo1_loss_fwd_bwd_objects_ex": (_c.c_int, [_vp, _i64p, _c.c_int, _c.c_int, _vp, _vp, _vp, _vp, _i64p, _vp,
                                                 _c.c_int64, _c.c_int, _c.c_int, _c.c_int,
                                                 _c.c_float, _c.c_float, _c.c_float, _c.c_int,
                                                 _vp, _c.c_size_t, _vp, _c.c_int, _vp]),
    "yolo1_scale_grad": (_c.c_int, [_vp, _c.c_int, _c.c_int64, _vp, _vp]),
    "yolo1_decode": (_c.c_int, [_vp, _i64p, _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_int,
                                _c.c_double, _vp, _vp, _vp, _vp, _vp]),
    "yolo1_nms": (_c.c_int, [_vp, _vp, _vp, _vp, _c.c_int64, _c.c_int, _c.c_float, _c.c_int,
                             _vp, _vp, _vp]),
    "yolo1_decode_nms": (_c.c_int, [_vp, _i64p, _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_int,
                                    _c.c_double, _c.c_float, _c.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "yolo1_boxes_to_pixels": (_c.c_int, [_vp, _c.c_int64, _c.c_float, _c.c_float, _vp, _vp]),
    "yolo1_encode_targets": (_c.c_int, [_vp, _vp, _vp, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _vp, _vp, _vp]),
    "yolo1_host_ctx_create": (_c.c_int, [_c.POINTER(_vp), _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int64]),
    "yolo1_host_ctx_destroy": (None, [_vp]),
    "yolo1_host_ctx_set_zero_copy": (_c.c_int, [_vp, _c.c_int]),
    "yolo1_host_pin": (_c.c_int, [_vp, _c.c_size_t]),
    "yolo1_host_unpin": (_c.c_int, [_vp]),
    "yolo1_loss_fwd_bwd_host": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _c.c_int64,
                                           _c.c_float, _c.c_float, _c.c_float, _c.c_int]),
    "yolo1_decode_nms_host": (_c.c_int, [_vp, _vp, _c.c_int64, _c.c_double, _c.c_float, _c.c_int,
                                         _vp, _vp, _vp, _vp]),
}

_lib = None


class Yolo1LibraryError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises Yolo1LibraryError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(SO_PATH):
            raise Yolo1LibraryError(
                "%s not found: the CUDA library is the only implementation of this package (no CPU "
                "fallback).  Build it with `make -C yolo_v1_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`." % SO_PATH)
        try:
            L = ctypes.CDLL(SO_PATH)
        except OSError as e:  # pragma: no cover
            raise Yolo1LibraryError("cannot load %s: %s" % (SO_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        got = L.yolo1_abi_version()
        if got != ABI_VERSION:
            raise Yolo1LibraryError("ABI version mismatch: library %d, binding %d" % (got, ABI_VERSION))
        _lib = L
    return _lib


def check(rc, what):
    """Turn a library return code into an exception (0 = success)."""
    if rc != 0:
        msg = lib().yolo1_error_string(int(rc))
        raise RuntimeError("%s failed: rc=%d (%s)" % (what, rc, msg.decode() if msg else "?"))


def strides4(t):
    """Element strides of a [N,S,S,D] tensor as an int64[4] ctypes array."""
    s = t.stride()
    return (_c.c_int64 * 4)(int(s[0]), int(s[1]), int(s[2]), int(s[3]))
