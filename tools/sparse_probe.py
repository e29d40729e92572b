"""does reading only the needed sectors beat the dense 360 B/cell stream on HBM?
Times the dense TMA kernel (variant 0) against the sector-read kernels on DEVICE-resident config-3 tensors, for
cudaLimitMaxL2FetchGranularity = default / 32 / 64 / 128.   python tools/sparse_probe.py [variants...]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yolo_v1_b200 as y
from yolo_v1_b200 import synth

N, S = 65536, 14
variants = [int(v) for v in sys.argv[1:]] or [0, 100, 40, 41]
pred, target = synth.make_loss_inputs(N, S, seed=20241018 + 3000, device="cuda")
grad = torch.empty_like(pred)
ref_grad = torch.empty_like(pred)
terms = torch.empty(5, device="cuda")
ws = torch.empty(1 << 17, dtype=torch.uint8, device="cuda")
cells = N * S * S
_, _, ref_terms = y.yolo_loss_fused(pred, target, batch_size=N, out_grad=ref_grad, workspace=ws)
ref_terms = ref_terms.clone()
rt = ctypes.CDLL("libcudart.so.12")
LIMIT = 0x05   # cudaLimitMaxL2FetchGranularity


def set_gran(v):
    if v is None:
        return "default"
    rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(v))
    got = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(got), LIMIT)
    return "%d (rc %d, now %d)" % (v, rc, got.value)


for gran in (None, 32, 64, 128):
    tag = set_gran(gran)
    for v in variants:
        for want_grad in (True, False):
            def run():
                y.yolo_loss_fused(pred, target, batch_size=N, variant=v, want_grad=want_grad,
                                  out_grad=grad if want_grad else None, out_terms=terms, workspace=ws)
            try:
                for _ in range(3):
                    run()
            except RuntimeError as e:
                print("granularity %s variant %d: %s" % (tag, v, e))
                break
            ok = bool(torch.allclose(terms, ref_terms, rtol=1e-6)) and (not want_grad or bool(torch.equal(grad, ref_grad)))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(20):
                run()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            print("granularity %-22s variant %3d grad=%d: %.3f ms  %.2f Gcells/s  dense-equivalent %.0f GB/s  same result: %s"
                  % (tag, v, want_grad, ms, cells / ms / 1e6, cells * (360 if want_grad else 240) / ms / 1e6, ok), flush=True)
