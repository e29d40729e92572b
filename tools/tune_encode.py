"""Development aid: throughput of the GPU target encoder and of the dense-target vs object-list loss calls it feeds.
python tools/tune_encode.py [S] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_v1_b200 as y

S = int(sys.argv[1]) if len(sys.argv) > 1 else 7
N = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
g = torch.Generator(device="cuda").manual_seed(1)
MAXC = int(sys.argv[3]) if len(sys.argv) > 3 else 7
counts = torch.randint(0, MAXC, (N,), generator=g, device="cuda")
offsets = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
offsets[1:] = counts.cumsum(0)
n = int(offsets[-1])
boxes = torch.rand((n, 4), generator=g, device="cuda")
labels = torch.randint(0, 20, (n,), generator=g, device="cuda", dtype=torch.int32)
out = torch.empty((N, S, S, 30), device="cuda")
for _ in range(3):
    y.encode_targets(boxes, labels, offsets, S, out=out, check=False)
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    y.encode_targets(boxes, labels, offsets, S, out=out, check=False)
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
byt = out.numel() * 4
print("encoder S=%d N=%d objects=%d: median %.3f ms  %.1f M images/s  %.0f GB/s written (%.2f GB target; the host would ship it over PCIe in %.0f ms)"
      % (S, N, n, ts[5], N / ts[5] / 1e3, byt / ts[5] / 1e6, byt / 1e9, byt / 55e9 * 1e3))
