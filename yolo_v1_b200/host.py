"""Host-buffer path: the C ABI's `yolo1_*_host` entry points (include/yolo1_b200.h) for callers that hold
CPU tensors -- what the reference's call sites do when run with device='cpu' (v1Loss.py:10,
utils/utils.py:94).  The arithmetic still runs on the GPU: the library pipelines H2D copy, kernel and D2H
copy over chunks of the batch.  Pinned tensors (`tensor.pin_memory()`) reach full PCIe rate.
"""
import ctypes

import torch

from . import _lib

_COORD_MODES = {"reference": _lib.COORD_REFERENCE, "paper": _lib.COORD_PAPER}


class HostContext:
    """Owns the device staging buffers, streams and events of the host-buffer pipeline (one per thread)."""

    def __init__(self, S, B=2, C=20, device=0, chunk_images=0):
        self.S, self.B, self.C, self.D = int(S), int(B), int(C), 5 * int(B) + int(C)
        self.max_n = self.S * self.S * self.B
        self.device = int(device)
        self._h = ctypes.c_void_p()
        L = _lib.lib()
        _lib.check(L.yolo1_host_ctx_create(ctypes.byref(self._h), self.device, self.S, self.B, self.C,
                                           int(chunk_images)), "yolo1_host_ctx_create")

    def set_zero_copy(self, enable):
        """Pinned+mapped loss buffers: 2 (default) = read / written in place by one kernel; 1 = target streamed by
        the copy engine, pred / grad in place; 3 = only pred in place; 0 = staged H2D / kernel / D2H pipeline."""
        _lib.check(_lib.lib().yolo1_host_ctx_set_zero_copy(self._h, int(enable)), "yolo1_host_ctx_set_zero_copy")

    def close(self):
        if self._h:
            _lib.lib().yolo1_host_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_host(self, t, name):
        if t.is_cuda:
            raise ValueError("%s must be a host tensor" % name)
        if t.dim() != 4 or tuple(t.shape[1:]) != (self.S, self.S, self.D):
            raise ValueError("%s must be [N,%d,%d,%d], got %s" % (name, self.S, self.S, self.D, tuple(t.shape)))
        if t.dtype != torch.float32:
            t = t.float()
        return t.contiguous()

    def loss(self, pred, target, batch_size, l_coord=5.0, l_noobj=0.5, coord_mode="reference", want_grad=True,
             out_grad=None):
        """Returns (terms float32[5] CPU tensor, grad CPU tensor or None)."""
        pred = self._check_host(pred, "pred")
        target = self._check_host(target, "target")
        if pred.shape != target.shape:
            raise ValueError("pred and target shapes differ")
        N = pred.shape[0]
        grad = None
        if want_grad:
            grad = out_grad if out_grad is not None else torch.empty_like(pred)
        terms = torch.empty(5, dtype=torch.float32)
        rc = _lib.lib().yolo1_loss_fwd_bwd_host(
            self._h, pred.data_ptr(), target.data_ptr(), grad.data_ptr() if grad is not None else None,
            terms.data_ptr(), N, float(l_coord), float(l_noobj), 1.0 / float(batch_size),
            _COORD_MODES[coord_mode])
        _lib.check(rc, "yolo1_loss_fwd_bwd_host")
        return terms, grad

    def decode_nms(self, pred, thresh=0.3, nms_th=0.5, per_class=False, out=None):
        """pred host [N,S,S,D] -> dict(boxes [N,M,4], scores [N,M], cls [N,M] int32, counts [N] int32), host
        tensors, detections in descending score order, rows beyond counts zero."""
        pred = self._check_host(pred, "pred")
        N, M = pred.shape[0], self.max_n
        if out is None:
            out = dict(boxes=torch.empty((N, M, 4), dtype=torch.float32),
                       scores=torch.empty((N, M), dtype=torch.float32),
                       cls=torch.empty((N, M), dtype=torch.int32),
                       counts=torch.empty((N,), dtype=torch.int32))
        rc = _lib.lib().yolo1_decode_nms_host(
            self._h, pred.data_ptr(), N, float(thresh), float(nms_th), int(bool(per_class)),
            out["boxes"].data_ptr(), out["scores"].data_ptr(), out["cls"].data_ptr(), out["counts"].data_ptr())
        _lib.check(rc, "yolo1_decode_nms_host")
        return out


def pin(t):
    """Page-lock an existing host tensor in place (cudaHostRegister) -- yolo1_host_pin."""
    _lib.check(_lib.lib().yolo1_host_pin(t.data_ptr(), t.numel() * t.element_size()), "yolo1_host_pin")
    return t


def unpin(t):
    _lib.check(_lib.lib().yolo1_host_unpin(t.data_ptr()), "yolo1_host_unpin")
    return t
