"""One launch of each secondary kernel at config-3 size, for ncu (run plain first, then under ncu with -k regex:...).
python tools/profile_variants.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_v1_b200 as y
from yolo_v1_b200 import synth

N, S = 65536, 14
pred, target = synth.make_loss_inputs(N, S, seed=20241018 + 3000, device="cuda")
planar = pred.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
cells = target[..., 0] == 1
idx = cells.nonzero()
bx = target[idx[:, 0], idx[:, 1], idx[:, 2], 2:6]
cxcy = (bx[:, :2] + torch.stack([idx[:, 2], idx[:, 1]], 1).float()) / S
boxes = torch.cat([cxcy, bx[:, 2:]], 1).contiguous()
labels = target[idx[:, 0], idx[:, 1], idx[:, 2], 10:].argmax(1).to(torch.int32)
offsets = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
offsets[1:] = cells.reshape(N, -1).sum(1).cumsum(0)
for _ in range(3):
    y.yolo_loss_fused(planar, target, batch_size=N)                               # loss_ws_kernel (planar, in place)
    y.yolo_loss_fused(planar.to(torch.bfloat16), target, batch_size=N)            # loss_ws_kernel bf16
    y.yolo_loss_from_objects(pred, boxes, labels, offsets, batch_size=N)          # loss_tma_kernel<..., LIST>
    y.yolo_loss_fused(pred, target, batch_size=N, from_logits=True)               # loss_tma_kernel<..., SIG>
    y.encode_targets(boxes, labels, offsets, S, check=False)                      # encode_kernel
torch.cuda.synchronize()
print("profile_variants ok")
