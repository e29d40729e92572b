"""Planar-view (permuted NCHW) loss kernels at config-3 size: confidence-first (variant 50) against the dense
planar kernels (variant 51), fp32 / bf16, dense target / object lists / fused head.  python tools/tune_planar.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yolo_v1_b200 as y
from yolo_v1_b200 import synth

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for S, N in ((14, 65536), (7, 262144)):
    pred, target = synth.make_loss_inputs(N, S, seed=20241018 + 3000, device="cuda")
    cells = N * S * S
    objmask = target[..., 0] == 1
    idx = objmask.nonzero()
    bx = target[idx[:, 0], idx[:, 1], idx[:, 2], 2:6]
    cxcy = (bx[:, :2] + torch.stack([idx[:, 2], idx[:, 1]], 1).float()) / S
    boxes = torch.cat([cxcy, bx[:, 2:]], 1).contiguous()
    labels = target[idx[:, 0], idx[:, 1], idx[:, 2], 10:].argmax(1).to(torch.int32)
    offs = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
    offs[1:] = objmask.reshape(N, -1).sum(1).cumsum(0)
    terms = torch.empty(5, device="cuda")
    ws = torch.empty(1 << 17, dtype=torch.uint8, device="cuda")
    wso = torch.empty(int(y._lib.lib().yolo1_loss_objects_workspace_bytes(N, S, 2, 20)), dtype=torch.uint8, device="cuda")
    for dt in (torch.float32, torch.bfloat16):
        planar = pred.to(dt).permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
        zplanar = torch.logit(pred.clamp(1e-4, 1 - 1e-4)).to(dt).permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
        g = torch.empty_like(planar)
        esz = planar.element_size()
        for name, dense_b, sparse_b, fn in (
            ("dense target", 120 + 60 * esz, 120 + 32 * esz,
             lambda v: y.yolo_loss_fused(planar, target, batch_size=N, out_grad=g, out_terms=terms, workspace=ws, variant=v)),
            ("object lists", 4 + 60 * esz, 4 + 32 * esz,
             lambda v: y.yolo_loss_from_objects(planar, boxes, labels, offs, batch_size=N, out_grad=g, workspace=wso, variant=v)),
            ("fused head  ", 120 + 60 * esz, 120 + 32 * esz,
             # (the logits entry point has no variant argument: both columns time the library's default for this case)
             lambda v: y.yolo_loss_fused(zplanar, target, batch_size=N, out_grad=g, out_terms=terms, workspace=ws, variant=v,
                                         from_logits=True)),
        ):
            ms_s, ms_d = timeit(lambda: fn(50)), timeit(lambda: fn(51))
            print("S=%2d %s %-12s confidence-first %.3f ms = %5.0f GB/s of its %3d B/cell (%.2f of measured) | dense %.3f ms = %5.0f GB/s of %3d B/cell (%.2f)"
                  % (S, "fp32" if esz == 4 else "bf16", name, ms_s, cells * sparse_b / ms_s / 1e6, sparse_b,
                     cells * sparse_b / ms_s / 1e6 / peak, ms_d, cells * dense_b / ms_d / 1e6, dense_b,
                     cells * dense_b / ms_d / 1e6 / peak), flush=True)
