"""Two launches of the fused decode+NMS kernel for ncu: BASELINE config 2 (4096 images, S=7) and the same workload at
65 536 images.  ncu --set full --import-source on -k regex:decode_nms -o gpurun_out/prof python tools/profile_decode.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yolo_v1_b200 as y
from yolo_v1_b200 import synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 7
for N in ((4096, 65536) if S == 7 else (4096, 16384)):
    pred = synth.make_decode_inputs(N, S, seed=2, device="cuda")
    M = S * S * 2
    outs = (torch.empty((N, M, 4), device="cuda"), torch.empty((N, M), dtype=torch.int32, device="cuda"),
            torch.empty((N, M), device="cuda"), torch.empty((N,), dtype=torch.int32, device="cuda"))
    y.decode_nms_batched(pred, 0.1, 0.5, out=outs)
    torch.cuda.synchronize()
print("profile_decode ok")
