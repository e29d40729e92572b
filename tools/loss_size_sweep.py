"""Loss kernel time against call size (S=7, contiguous NHWC fp32): t = a + b * cells separates the fixed cost of a
launch (memset node, pipeline fill, last-CTA fix-up and tail) from the streaming rate -- what strong scaling (BASELINE
config 4: 1 M images over 8 GPUs = 131 072 images per call) pays.  python tools/loss_size_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yolo_v1_b200 as y
from yolo_v1_b200 import synth

S = 7
pred, target = synth.make_loss_inputs(1 << 20, S, seed=1, device="cuda")
grad = torch.empty_like(pred)
terms = torch.empty(5, device="cuda")
ws = torch.empty(1 << 17, dtype=torch.uint8, device="cuda")
for n in (1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
    for mode in ("reference", "paper"):
        def f():
            y.yolo_loss_fused(pred[:n], target[:n], batch_size=n, out_grad=grad[:n], out_terms=terms, workspace=ws,
                              coord_mode=mode)
        for _ in range(5):
            f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        reps = max(10, (1 << 22) // n)
        a.record()
        for _ in range(reps):
            f()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        cells = n * S * S
        print("images %8d cells %9d %-9s %.4f ms  %.0f GB/s  (%.1f us over %.5f ms/Mcell)" % (
            n, cells, mode, ms, cells * 360 / ms / 1e6, (ms - cells / 1e6 * 0.05472) * 1e3, 0.05472), flush=True)
