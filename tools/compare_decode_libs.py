"""Development aid: decode+NMS throughput of ONE given build of the library (path as argv[1]), for A/B runs of several
builds inside one gpurun call (box-to-box differences are larger than most code changes):
    for f in tools/_exp/lib_*.so; do python tools/compare_decode_libs.py $f; done
Prints M images/s for S=7 N=4096 uniform, N=65536 uniform, N=65536 sigmoid, S=14 N=16384 uniform and sigmoid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_v1_b200 import _lib
_lib.SO_PATH = os.path.abspath(sys.argv[1])
import yolo_v1_b200 as y
from yolo_v1_b200 import synth
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = []
for S, N, dist in [(7, 4096, "uniform"), (7, 65536, "uniform"), (7, 65536, "sigmoid"), (14, 16384, "uniform"),
                   (14, 16384, "sigmoid")]:
    pred = synth.make_decode_inputs(N, S, seed=2, device="cuda", dist=dist)
    M = S * S * 2
    outs = (torch.empty((N, M, 4), device="cuda"), torch.empty((N, M), dtype=torch.int32, device="cuda"),
            torch.empty((N, M), device="cuda"), torch.empty((N,), dtype=torch.int32, device="cuda"))
    for _ in range(3):
        y.decode_nms_batched(pred, 0.1, 0.5, out=outs)
    ts = []
    for _ in range(15):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y.decode_nms_batched(pred, 0.1, 0.5, out=outs); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    res.append("%.2f" % (N / ts[len(ts) // 2] / 1e3))
print(os.path.basename(sys.argv[1]), " ".join(res), flush=True)
