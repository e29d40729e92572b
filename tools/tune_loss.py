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
for name, p, g, variants in (("nhwc", pred, grad, [0, 5, 8, 13]), ("planar-view", planar, gplanar, [0, 1, 20, -1])):
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

# ---- fused sigmoid head: pred holds logits ----
logit = torch.logit(pred.float().clamp(1e-4, 1 - 1e-4)).to(pred.dtype)
lplanar = logit.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
for name, p_, g_ in (("nhwc", logit, grad), ("planar-view", lplanar, gplanar)):
    for _ in range(3):
        y.yolo_loss_fused(p_, target, batch_size=N, out_grad=g_, out_terms=terms, workspace=ws, from_logits=True)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y.yolo_loss_fused(p_, target, batch_size=N, out_grad=g_, out_terms=terms, workspace=ws, from_logits=True)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    b = cells * bytes_per_cell
    print("%-12s fused sigmoid head   median %.3f ms  min %.3f ms  %.2f Gcells/s  %.0f GB/s (%.3f of measured)"
          % (name, ts[5], ts[0], cells / ts[5] / 1e6, b / ts[5] / 1e6, b / ts[5] / 1e6 / peak), flush=True)

# ---- loss from object lists (no dense target): 248 algorithmic bytes per cell (fp32) ----
tgt_cells = (target[..., 0] == 1)
cnt = tgt_cells.reshape(N, -1).sum(1)
offsets = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
offsets[1:] = cnt.cumsum(0)
idx = tgt_cells.nonzero()                                   # (n, i, j) row-major = input order per image
bx = target[idx[:, 0], idx[:, 1], idx[:, 2], 2:6]
cxcy = (bx[:, :2] + torch.stack([idx[:, 2], idx[:, 1]], 1).float()) / S
boxes_l = torch.cat([cxcy, bx[:, 2:]], 1).contiguous()
labels_l = target[idx[:, 0], idx[:, 1], idx[:, 2], 10:].argmax(1).to(torch.int32)
ws2 = torch.empty(int(y._lib.lib().yolo1_loss_objects_workspace_bytes(N, S, 2, 20)), dtype=torch.uint8, device="cuda")
for name, p, v in (("nhwc", pred, 0), ("planar-view", planar, 0)):
    def run():
        y.yolo_loss_from_objects(p, boxes_l, labels_l, offsets, batch_size=N, variant=v, out_grad=grad if p is pred else gplanar, workspace=ws2)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):   # 20 calls back to back per sample: the call is five stream operations, one call alone is launch-bound
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / 20)
    ts.sort()
    med = ts[5]
    b = cells * (30 * 2 * pred.element_size() + 8)
    print("%-12s object-list targets v%d median %.3f ms  min %.3f ms  %.2f Gcells/s  %.0f GB/s of 248-byte cells (%.3f of measured)"
          % (name, v, med, ts[0], cells / med / 1e6, b / med / 1e6, b / med / 1e6 / peak), flush=True)
