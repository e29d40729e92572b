"""Small run of every kernel of libyolo1_b200.so for compute-sanitizer (memcheck / racecheck / synccheck), one
tool per gpurun call:   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yolo_v1_b200 as y
from yolo_v1_b200 import synth

for S, N in ((7, 37), (14, 9)):
    pred, target = synth.make_loss_inputs(N, S, seed=1, p_obj=0.2, device="cuda")
    planar = pred.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    for p in (pred, planar, pred.to(torch.bfloat16), planar.to(torch.bfloat16)):
        for variant in (0, 8, -1, 31, 50):      # 0: small-call cluster kernel here; 31: streaming; 50: confidence-first planar
            try:
                y.yolo_loss_fused(p, target, batch_size=N, variant=variant)
                y.yolo_loss_fused(p, target, batch_size=N, variant=variant, want_grad=False)
            except RuntimeError:
                assert variant == 50      # only defined for the planar view

        y.yolo_loss_fused(p, target, batch_size=N, from_logits=True)
    mod = y.YOLOLossV1(N, S, 2, 20)
    q = pred.clone().requires_grad_(True)
    (mod(q, target) * 2.0).backward()
    ctx = y.HostContext(S, chunk_images=4)
    hp, ht = pred.cpu().pin_memory(), target.cpu().pin_memory()
    hg = torch.empty_like(hp).pin_memory()
    for mode in (0, 1, 2, 3):
        ctx.set_zero_copy(mode)
        ctx.loss(hp, ht, batch_size=N, out_grad=hg)
    dp, _ = synth.make_tie_free_decode_inputs(N, S, seed=2)
    ctx.decode_nms(dp, 0.1, 0.5)
    ctx.close()
    b, c, s, k = y.decode_nms_batched(dp.cuda(), 0.1, 0.5)
    y.decode_nms_batched(dp.cuda(), 0.005, 0.45, class_agnostic=False, return_keep=True)
    bx, sc, cl, cnt = y.decode_batched(dp.cuda(), 0.1)
    y.nms_batched(bx, sc, cnt, 0.5, cls=cl, per_class=True)
    y.boxes_to_pixels(b)
    y.decoder(dp[:1], grid_num=S, thresh=0.1)
    boxes = torch.rand(50, 4)
    labels = torch.randint(0, 20, (50,))
    offs = torch.tensor([0, 3, 3, 10, 25, 50])
    y.encode_targets(boxes.cuda(), labels.cuda(), offs.cuda(), S)
    bb, ll, oo = y.pack_objects([torch.rand(3, 4) * 0.8 + 0.1 for _ in range(N)], [torch.randint(0, 20, (3,)) for _ in range(N)])
    y.yolo_loss_from_objects(pred, bb.cuda(), ll.cuda(), oo.cuda(), batch_size=N)
    y.yolo_loss_from_objects(planar, bb.cuda(), ll.cuda(), oo.cuda(), batch_size=N, variant=31)
g = torch.Generator().manual_seed(0)
xy = torch.rand(1000, 2, generator=g) * 0.8
y.nms(torch.cat([xy, xy + 0.1], 1), torch.rand(1000, generator=g), 0.5)
torch.cuda.synchronize()
print("sanitize_smoke ok")
