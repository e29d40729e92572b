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
        the copy engine, pred / grad in place; 3 = only pred in place; 4 = pred and target in place, gradient through
        the copy engine; 0 = staged H2D / kernel / D2H pipeline (host_ctx.cu).  Values outside 0..4 are clamped."""
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

    @staticmethod
    def _check_out(t, shape, dtype, name):
        """Caller-provided output buffers go to the library as raw pointers: a wrong size, dtype or layout would
        let the D2H copy / the host-mapped kernel write past the buffer.  Validate before the call."""
        if not isinstance(t, torch.Tensor) or t.is_cuda:
            raise ValueError("%s must be a host tensor" % name)
        if tuple(t.shape) != tuple(shape) or t.dtype != dtype or not t.is_contiguous():
            raise ValueError("%s must be a contiguous %s host tensor of shape %s, got %s %s%s" %
                             (name, dtype, tuple(shape), t.dtype, tuple(t.shape),
                              "" if t.is_contiguous() else " (non-contiguous)"))
        return t

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
            grad = (self._check_out(out_grad, pred.shape, torch.float32, "out_grad") if out_grad is not None
                    else torch.empty_like(pred))
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
        else:
            for key, shape, dt in (("boxes", (N, M, 4), torch.float32), ("scores", (N, M), torch.float32),
                                   ("cls", (N, M), torch.int32), ("counts", (N,), torch.int32)):
                if key not in out:
                    raise ValueError("out must hold 'boxes', 'scores', 'cls' and 'counts'")
                self._check_out(out[key], shape, dt, "out[%r]" % key)
        rc = _lib.lib().yolo1_decode_nms_host(
            self._h, pred.data_ptr(), N, float(thresh), float(nms_th), int(bool(per_class)),
            out["boxes"].data_ptr(), out["scores"].data_ptr(), out["cls"].data_ptr(), out["counts"].data_ptr())
        _lib.check(rc, "yolo1_decode_nms_host")
        return out


ZERO_COPY_MODES = {0: "staged H2D / kernel / D2H pipeline (copy engines only)",
                   1: "target by copy engine, pred / grad in place",
                   2: "pred, target and grad in place (one kernel, sector reads)",
                   3: "pred in place, target and grad by copy engine",
                   4: "pred and target in place, grad by copy engine"}


def autotune_zero_copy(ctx, pred, target, out_grad, batch_size, modes=(2, 0), repeats=2, barrier=None,
                       reduce_max=None):
    """Pick the transfer mode of `ctx.loss` by measurement on the caller's own buffers and set it.

    Which mode wins depends on the host, not on the GPU: the in-place modes issue one small PCIe read per cell
    and every GPU of the box shares the host's ceiling on such requests, while the staged pipeline moves dense
    bursts on each GPU's own link (the in-place mode wins on one GPU and loses on eight).  Under
    torch.distributed pass `barrier` (called before each timed loop, so that all ranks load the host together)
    and `reduce_max` (ms -> max over ranks) so that every rank times the same contention and picks the same mode.
    Returns (best_mode, {mode: ms_per_call})."""
    import time
    table = {}
    for mode in modes:
        ctx.set_zero_copy(mode)
        ctx.loss(pred, target, batch_size=batch_size, out_grad=out_grad)          # warm: staging buffers, page maps
        if barrier is not None:
            barrier()
        t0 = time.perf_counter()
        for _ in range(repeats):
            ctx.loss(pred, target, batch_size=batch_size, out_grad=out_grad)
        ms = (time.perf_counter() - t0) * 1e3 / repeats
        table[mode] = float(reduce_max(ms)) if reduce_max is not None else ms
    best = min(table, key=table.get)
    ctx.set_zero_copy(best)
    return best, table


def gpu_numa_cpus(device=0):
    """CPUs of the NUMA node the GPU's PCIe slot hangs off (sysfs), or None when the platform does not say
    (numa_node = -1 on most virtualised hosts)."""
    import os
    try:
        p = torch.cuda.get_device_properties(device)
        addr = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % addr).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:
        return None


class near_gpu:
    """Context manager: run the body (pinned allocations and their first touch) on the CPUs next to the GPU so the
    pages land on that NUMA node; the previous affinity is restored on exit.  A no-op where sysfs has no answer."""

    def __init__(self, device=0):
        self.cpus = gpu_numa_cpus(device)
        self.saved = None

    def __enter__(self):
        import os
        if self.cpus:
            try:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.cpus)
            except Exception:
                self.saved = None
        return self

    def __exit__(self, *exc):
        import os
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


def pin(t):
    """Page-lock an existing host tensor in place (cudaHostRegister) -- yolo1_host_pin."""
    _lib.check(_lib.lib().yolo1_host_pin(t.data_ptr(), t.numel() * t.element_size()), "yolo1_host_pin")
    return t


def unpin(t):
    _lib.check(_lib.lib().yolo1_host_unpin(t.data_ptr()), "yolo1_host_unpin")
    return t
