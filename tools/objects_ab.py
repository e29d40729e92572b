"""Object-list loss at config-3 size: pre-pass + ownership map (variant 61) against owners found inside the streaming
kernel (variant 60), calls back to back.   python tools/objects_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_v1_b200 as y
from yolo_v1_b200 import synth, _lib


def lists_of(target, S):
    cells = target[..., 0] == 1
    idx = cells.nonzero()
    bx = target[idx[:, 0], idx[:, 1], idx[:, 2], 2:6]
    cxcy = (bx[:, :2] + torch.stack([idx[:, 2], idx[:, 1]], 1).float()) / S
    boxes = torch.cat([cxcy, bx[:, 2:]], 1).contiguous()
    labels = target[idx[:, 0], idx[:, 1], idx[:, 2], 10:].argmax(1).to(torch.int32)
    offsets = torch.zeros(target.shape[0] + 1, dtype=torch.int64, device="cuda")
    offsets[1:] = cells.reshape(target.shape[0], -1).sum(1).cumsum(0)
    return boxes, labels, offsets


def timed(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for N, S, dt in [(65536, 14, torch.float32), (65536, 14, torch.bfloat16), (262144, 7, torch.float32)]:
    pred, target = synth.make_loss_inputs(N, S, seed=20241018 + 3000, device="cuda")
    boxes, labels, offsets = lists_of(target, S)
    pred = pred.to(dt)
    grad, terms = torch.empty_like(pred), torch.empty(5, device="cuda")
    ws = torch.zeros(int(_lib.lib().yolo1_loss_objects_workspace_bytes(N, S, 2, 20)), dtype=torch.uint8, device="cuda")
    out = {}
    for v in (61, 60, 61, 60):
        kw = dict(batch_size=N, variant=v, out_grad=grad, workspace=ws)
        ms = timed(lambda: y.yolo_loss_from_objects(pred, boxes, labels, offsets, **kw))
        out.setdefault(v, []).append(ms)
    cells = N * S * S
    bpc = 248 if dt == torch.float32 else 128
    print("N=%d S=%d %s objects=%d  pre-pass+map %s ms  in-kernel %s ms  (%.0f B/cell -> %.2f / %.2f TB/s)" % (
        N, S, str(dt).split(".")[1], boxes.shape[0], ["%.4f" % m for m in out[61]], ["%.4f" % m for m in out[60]], bpc,
        bpc * cells / min(out[61]) / 1e9, bpc * cells / min(out[60]) / 1e9), flush=True)
