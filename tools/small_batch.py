"""Latency of small loss calls: the cluster kernel (default below 16 384 cells, csrc/loss_small.cu) against the
streaming kernels on the same call (variant 31), eager and under a CUDA graph.  python tools/small_batch.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yolo_v1_b200 as y
from yolo_v1_b200 import synth


def timeit(fn, n=300):
    for _ in range(10):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def graphed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    return g.replay


ws = torch.empty(1 << 17, dtype=torch.uint8, device="cuda")
for N, S in [(32, 7), (12, 14), (16, 14), (64, 7), (32, 14), (64, 14), (83, 14), (334, 7), (128, 14), (256, 14)]:
    p, t = synth.make_loss_inputs(N, S, seed=1, device="cuda")
    g, tm = torch.empty_like(p), torch.empty(5, device="cuda")
    planar = p.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    gp = torch.empty_like(planar)
    row = {}
    for name, variant in (("cluster", 0), ("streaming", 31)):
        if variant == 0 and N * S * S > 16384:
            continue
        f = lambda: y.yolo_loss_fused(p, t, batch_size=N, out_grad=g, out_terms=tm, workspace=ws, variant=variant)
        fpl = lambda: y.yolo_loss_fused(planar, t, batch_size=N, out_grad=gp, out_terms=tm, workspace=ws, variant=variant)
        row[name] = (timeit(f), timeit(graphed(f)), timeit(graphed(fpl)))
    print("N=%4d S=%2d cells=%6d  " % (N, S, N * S * S) + "  ".join(
        "%s: eager %.1f us, graph %.1f us, graph planar-view %.1f us" % ((k,) + v) for k, v in row.items()), flush=True)
