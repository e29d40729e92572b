"""BASELINE config 5: end-to-end ResNet50-YOLOv1 448x448 bf16 (DDP) train step with the fused loss kernel.
    python tools/train_step_bench.py [--batch 64] [--S 7] [--steps 20]            (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py ...
Synthetic images and targets, random-init weights.  Prints one JSON line on rank 0: images/s (whole job), ms per
step and the share of the step spent in the loss kernel (CUDA events around the loss call)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from yolo_v1_b200 import synth
from yolo_v1_b200.trainstep import TrainStep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--S", type=int, default=7)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-fuse-head", action="store_true")
    ap.add_argument("--backbone", default="resnet50", choices=["resnet50", "densenet121"])
    ap.add_argument("--graph-loss", action="store_true", help="loss kernel + terms all-reduce replayed as one CUDA graph")
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ts = TrainStep(S=a.S, batch_size=a.batch, device="cuda:%d" % lr, ddp=world > 1, fuse_head=not a.no_fuse_head, backbone=a.backbone,
                   graph_loss=a.graph_loss)
    images = torch.randn(a.batch, 3, 448, 448, device="cuda").to(memory_format=torch.channels_last)
    _, target = synth.make_loss_inputs(a.batch, a.S, seed=1 + rank, device="cuda")
    for _ in range(a.warmup):
        loss = ts.step(images, target)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = ts.step(images, target)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    # the loss call alone, on a prediction of the same shape / layout / dtype
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        pred = ts.net(images)
    pred = pred.detach().requires_grad_(True)
    loss_call = ts.graphed if a.graph_loss else ts.loss
    for _ in range(3):
        loss_call(pred, target).backward()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(50):
        loss_call(pred, target).backward()
    l1.record()
    torch.cuda.synchronize()
    loss_ms = l0.elapsed_time(l1) / 50
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "config5: %s-YOLOv1 448x448 bf16 %s train step, S=%d, batch %d per GPU, fused loss%s" %
                          (a.backbone, "DDP x%d" % world if world > 1 else "single GPU", a.S, a.batch, "" if a.no_fuse_head else " + fused sigmoid head"),
                          "images_per_s": a.batch * world / (float(t) * 1e-3), "ms_per_step": float(t), "n_gpus": world,
                          "loss_fwd_bwd_ms": loss_ms, "graph_loss": bool(a.graph_loss), "loss_share_of_step": loss_ms / float(t), "loss": float(loss),
                          "pred_dtype": str(pred.dtype), "pred_strides": list(pred.stride())}))
    ts.close()      # a graph that captured the all-reduce goes before the process group
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
