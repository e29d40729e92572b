"""Development aid: decode+NMS throughput over grid sizes, batch sizes and input distributions (CUDA events,
L2 flushed between iterations).  python tools/tune_decode.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_v1_b200 as y
from yolo_v1_b200 import synth

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for S, N, dist, th, nth in [(7, 4096, "uniform", 0.1, 0.5), (7, 65536, "uniform", 0.1, 0.5), (7, 65536, "sigmoid", 0.1, 0.5),
                            (7, 65536, "uniform", 0.005, 0.45), (14, 4096, "uniform", 0.1, 0.5), (14, 16384, "uniform", 0.1, 0.5),
                            (14, 16384, "sigmoid", 0.1, 0.5), (14, 16384, "sigmoid", 0.005, 0.45)]:
    pred = synth.make_decode_inputs(N, S, seed=2, device="cuda", dist=dist)
    M = S * S * 2
    outs = (torch.empty((N, M, 4), device="cuda"), torch.empty((N, M), dtype=torch.int32, device="cuda"),
            torch.empty((N, M), device="cuda"), torch.empty((N,), dtype=torch.int32, device="cuda"))
    for _ in range(3):
        y.decode_nms_batched(pred, th, nth, out=outs)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        y.decode_nms_batched(pred, th, nth, out=outs)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    _, _, _, cnt, _, cand = y.decode_nms_batched(pred, th, nth, return_keep=True)
    print("S=%2d N=%6d %-8s thresh=%.3f nms=%.2f  cand %.1f kept %.1f  median %.3f ms  %.2f M images/s  %.1f GB/s read"
          % (S, N, dist, th, nth, cand.float().mean().item(), cnt.float().mean().item(), ts[5], N / ts[5] / 1e3,
             N * S * S * 120 / ts[5] / 1e6), flush=True)
