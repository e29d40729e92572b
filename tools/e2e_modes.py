"""The host-buffer loss call (yolo1_loss_fwd_bwd_host) in every transfer mode at config-3 size, on 1..8 GPUs together.
    python tools/e2e_modes.py
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_modes.py
Every rank drives its own GPU from its own pinned buffers; the loops start together (barrier) and the reported time is
the slowest rank's, so a row shows what the HOST sustains with N GPUs pulling at once (the in-place
mode shares one host-wide ceiling on small PCIe reads, the staged pipeline moves dense bursts on each GPU's own link).
One table on rank 0: mode x {ms per call, M cells/s over all ranks}."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import yolo_v1_b200 as y
from yolo_v1_b200 import host as yhost
from yolo_v1_b200 import synth

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, S = 65536, 14
pred, target = synth.make_loss_inputs(N, S, seed=1 + rank, device=dev)
with yhost.near_gpu(lr) as numa:
    hp = torch.empty(pred.shape, pin_memory=True)
    ht = torch.empty(pred.shape, pin_memory=True)
    hg = torch.empty(pred.shape, pin_memory=True)
    hp.copy_(pred), ht.copy_(target)
torch.cuda.synchronize()
del pred, target


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(v):
    if world > 1:
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return v


if rank == 0:
    print("ranks %d, NUMA-bound pinned allocation: %s, host cores %d" % (world, bool(numa.cpus), os.cpu_count()), flush=True)
ctx = y.HostContext(S, device=lr)
for mode in (0, 1, 2, 3, 4):
    for grad in (True, False):
        ctx.set_zero_copy(mode)
        for _ in range(2):
            ctx.loss(hp, ht, batch_size=N, out_grad=hg if grad else None, want_grad=grad)
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.loss(hp, ht, batch_size=N, out_grad=hg if grad else None, want_grad=grad)
        ms = max_over_ranks((time.perf_counter() - t0) / 5 * 1e3)
        barrier()
        if rank == 0:
            print("ranks %d mode %d (%s) grad=%d: %.1f ms per call, %.0f M cells/s over all ranks" % (
                world, mode, yhost.ZERO_COPY_MODES[mode][:48], grad, ms, world * N * S * S / ms / 1e3), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
