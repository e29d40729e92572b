"""BASELINE config 4: 1 M-image synthetic loss + decode + NMS, batch-sharded across the GPUs of one box
(strong scaling: the total is fixed at 1 048 576 images, S=7), one NCCL all-reduce of the loss terms per step.
    python tools/config4_bench.py                                  (1 GPU: all 1 M images)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/config4_bench.py
One JSON line on rank 0."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import yolo_v1_b200 as y
from yolo_v1_b200 import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S = 7
    lo, hi = y.shard_range(a.images, rank, world)
    n = hi - lo
    # synthetic shard, generated in slabs to bound the generator's temporaries
    pred = torch.empty((n, S, S, 30), device=dev)
    target = torch.empty((n, S, S, 30), device=dev)
    slab = 1 << 16
    for s0 in range(0, n, slab):
        m = min(slab, n - s0)
        p, t = synth.make_loss_inputs(m, S, seed=20241018 + 4000 + (lo + s0), device=dev)
        pred[s0:s0 + m], target[s0:s0 + m] = p, t
    grad = torch.empty_like(pred)
    terms = torch.empty(5, device=dev)
    ws = torch.empty(1 << 17, dtype=torch.uint8, device=dev)
    M = S * S * 2
    outs = (torch.empty((n, M, 4), device=dev), torch.empty((n, M), dtype=torch.int32, device=dev),
            torch.empty((n, M), device=dev), torch.empty((n,), dtype=torch.int32, device=dev))
    gterms = torch.empty(5, device=dev)

    def step():
        y.yolo_loss_fused(pred, target, batch_size=n, out_grad=grad, out_terms=terms, workspace=ws)
        if world > 1:
            gterms.copy_(terms)
            dist.all_reduce(gterms)
        y.decode_nms_batched(pred, 0.1, 0.5, out=outs)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    l0.record()
    for _ in range(a.steps):
        y.yolo_loss_fused(pred, target, batch_size=n, out_grad=grad, out_terms=terms, workspace=ws)
    l1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps, l0.elapsed_time(l1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, loss_ms = float(t[0]), float(t[1])
    kept = int(outs[3].sum().item())
    if rank == 0:
        print(json.dumps({"config": "config4: %d images total (S=7), loss fwd+bwd + decode + NMS, batch-sharded x%d, strong scaling" % (a.images, world),
                          "n_gpus": world, "images_per_s": a.images / (ms * 1e-3), "ms_per_step": ms,
                          "loss_ms_per_step": loss_ms, "loss_cells_per_s": a.images * 49 / (loss_ms * 1e-3),
                          "loss_hbm_gbs_per_gpu": n * 49 * 360 / (loss_ms * 1e-3) / 1e9, "images_per_gpu": n,
                          "kept_detections_rank0": kept, "steps": a.steps}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
