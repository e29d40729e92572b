"""Development aid: time every launch shape of the streaming loss kernel (yolo1_loss_fwd_bwd_ex variants)
on a BASELINE config with CUDA events.  Run on the GPU box:  python tools/tune_loss.py [S] [N]"""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_v1_b200 as y
from yolo_v1_b200 import synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 14
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dtype = torch.bfloat16 if (len(sys.argv) > 3 and sys.argv[3] == "bf16") else torch.float32
pred, target = synth.make_loss_inputs(N, S, seed=20241018 + 3000, device="cuda")
pred = pred.to(dtype)
grad = torch.empty_like(pred)
terms = torch.empty(5, device="cuda")
ws = torch.empty(1 << 17, dtype=torch.uint8, device="cuda")
cells = N * S * S
bytes_per_cell = 30 * (4 + 2 * pred.element_size())
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
planar = pred.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
gplanar = torch.empty_like(planar)
for name, p, g, variants in (("nhwc", pred, grad, [0, 5, 8, 9, 10, 11, 12, 13]), ("planar-view", planar, gplanar, [0, 1, -1])):
    for v in variants:
        for want_grad in (True, False):
            def run():
                y.yolo_loss_fused(p, target, batch_size=N, variant=v, want_grad=want_grad, out_grad=g if want_grad else None,
                                  out_terms=terms, workspace=ws)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            med = ts[len(ts) // 2]
            b = cells * (bytes_per_cell if want_grad else bytes_per_cell - 30 * pred.element_size())
            print("%-12s variant %2d grad=%d  median %.3f ms  min %.3f ms  %.2f Gcells/s  %.0f GB/s (%.3f of measured %.0f)"
                  % (name, v, want_grad, med, ts[0], cells / med / 1e6, b / med / 1e6, b / med / 1e6 / peak, peak), flush=True)
