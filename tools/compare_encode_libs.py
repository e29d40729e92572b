"""Development aid: encoder throughput of ONE given build of the library (path as argv[1]); A/B several builds inside
one gpurun call.   python tools/compare_encode_libs.py <lib.so>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_v1_b200 import _lib
_lib.SO_PATH = os.path.abspath(sys.argv[1])
import yolo_v1_b200 as y
res = []
for S, N, maxc in ((14, 65536, 7), (7, 262144, 7), (14, 65536, 1), (14, 65536, 13)):
    g = torch.Generator(device="cuda").manual_seed(1)
    counts = torch.randint(0, maxc, (N,), generator=g, device="cuda")
    offsets = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
    offsets[1:] = counts.cumsum(0)
    n = int(offsets[-1])
    boxes = torch.rand((max(n, 1), 4), generator=g, device="cuda")
    labels = torch.randint(0, 20, (max(n, 1),), generator=g, device="cuda", dtype=torch.int32)
    out = torch.empty((N, S, S, 30), device="cuda")
    for _ in range(3):
        y.encode_targets(boxes, labels, offsets, S, out=out, check=False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20):
        y.encode_targets(boxes, labels, offsets, S, out=out, check=False)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    res.append("S=%d<%dobj %.3fms=%.0fGB/s" % (S, maxc, ms, out.numel() * 4 / ms / 1e6))
print(os.path.basename(sys.argv[1]), "  ".join(res), flush=True)
