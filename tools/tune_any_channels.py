import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_v1_b200 as y
from yolo_v1_b200 import synth
for B, C, N, S in [(2, 80, 131072, 7), (2, 20, 262144, 7), (1, 20, 262144, 7), (3, 80, 65536, 7)]:
    D = 5 * B + C
    pred, target = synth.make_loss_inputs(N, S, B=B, C=C, seed=1, device="cuda")
    grad = torch.empty_like(pred)
    for variant, name in ((0, "bulk-copy"), (-1, "strided")):
        for _ in range(3):
            y.yolo_loss_fused(pred, target, batch_size=N, B=B, C=C, out_grad=grad, variant=variant)
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); y.yolo_loss_fused(pred, target, batch_size=N, B=B, C=C, out_grad=grad, variant=variant); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        byt = N * S * S * D * 12
        print("B=%d C=%d D=%d cells=%d %-9s median %.3f ms  %.0f GB/s (%.3f of 6555)" % (B, C, D, N * S * S, name, ts[5], byt / ts[5] / 1e6, byt / ts[5] / 1e6 / 6555), flush=True)
