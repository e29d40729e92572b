"""CPU-only, world_size 2 over gloo: the N>1 host logic -- contiguous batch shards, per-shard loss with
the per-rank batch size, one all_reduce(sum) of the 5-float terms vector.
The per-rank compute is played by the CPU oracle here (tests may call it; the product path uses the CUDA
kernels); what is under test is the sharding and the reduction plumbing in yolo_v1_b200/dist.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from yolo_v1_b200 import dist as ydist
from yolo_v1_b200 import synth

N, S = 37, 7     # odd on purpose: ranks get 19 and 18 images


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pred, target = synth.make_loss_inputs(N, S, seed=4242, p_obj=0.1)
        a, b = ydist.shard_range(N, rank, world)
        terms, _ = O.loss(pred[a:b].numpy(), target[a:b].numpy(), batch_size=b - a)
        t = torch.from_numpy(terms.copy())
        summed = ydist.all_reduce_terms(t.clone(), average=False)
        mean = ydist.all_reduce_terms(t.clone(), average=True)
        np.save(os.path.join(out_dir, "r%d.npy" % rank), np.stack([t.numpy(), summed.numpy(), mean.numpy()]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_terms(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(tmp_path, "r%d.npy" % r)) for r in range(world)]
    pred, target = synth.make_loss_inputs(N, S, seed=4242, p_obj=0.1)
    local = []
    for r in range(world):
        a, b = ydist.shard_range(N, r, world)
        t, _ = O.loss(pred[a:b].numpy(), target[a:b].numpy(), batch_size=b - a)
        local.append(t)
        assert np.array_equal(res[r][0], t)
    want = local[0] + local[1]
    for r in range(world):
        assert np.allclose(res[r][1], want, rtol=1e-6)
        assert np.allclose(res[r][2], want / world, rtol=1e-6)
    assert np.array_equal(res[0][1], res[1][1])      # every rank holds the same reduced vector


def test_all_reduce_is_identity_without_process_group():
    t = torch.arange(5, dtype=torch.float32)
    assert torch.equal(ydist.all_reduce_terms(t.clone()), t)
