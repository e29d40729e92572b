"""Two ranks over NCCL (needs 2 GPUs; skipped otherwise): every rank runs the fused loss on its contiguous shard,
the 5-float terms vector is all-reduced (yolo_v1_b200/dist.py), per-shard results match the oracle."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, S = 1000, 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import yolo_v1_b200 as y
    from yolo_v1_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    import datetime
    # a short collective timeout: a hang must fail this test in a minute, not in the watchdog's default ten
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=60))
    try:
        pred, target = synth.make_loss_inputs(N, S, seed=77, p_obj=0.1)
        a, b = y.shard_range(N, rank, world)
        mod = y.YOLOLossV1(b - a, S, 2, 20)
        p = pred[a:b].cuda().requires_grad_(True)
        local, gterms = y.sharded_loss(mod, p, target[a:b].cuda(), average=False)
        local.backward()
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([mod.last_terms.cpu().numpy(), gterms.cpu().numpy()]))
        np.save(os.path.join(out_dir, "g%d.npy" % rank), p.grad.cpu().numpy())
        if os.environ.get("YOLO1_SKIP_GRAPH_NCCL") == "1":
            return
        # the same exchange as ONE captured CUDA graph (loss kernel + NCCL all-reduce), replayed twice
        gl = y.GraphedLoss(b - a, S, 2, 20, average=False)
        for it in range(2):
            q = pred[a:b].cuda().requires_grad_(True)
            gl(q, target[a:b].cuda()).backward()
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, "gr%d.npy" % rank), gl.global_terms.cpu().numpy())
        np.save(os.path.join(out_dir, "gg%d.npy" % rank), q.grad.cpu().numpy())
        gl.release()      # a graph that captured a collective must be gone before the process group is destroyed
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharded_loss_over_nccl(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from yolo_v1_b200 import dist as ydist
    from yolo_v1_b200 import synth
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    pred, target = synth.make_loss_inputs(N, S, seed=77, p_obj=0.1)
    total = np.zeros(5, np.float64)
    for r in range(2):
        a, b = ydist.shard_range(N, r, 2)
        o_terms, o_grad = O.loss(pred[a:b].numpy(), target[a:b].numpy(), batch_size=b - a)
        got = np.load(os.path.join(tmp_path, "r%d.npy" % r))
        grad = np.load(os.path.join(tmp_path, "g%d.npy" % r))
        assert np.allclose(got[:5], o_terms, rtol=1e-5)
        assert np.abs(grad - o_grad).max() <= 1e-5 * np.abs(o_grad).max()
        total += o_terms
    for r in range(2):
        got = np.load(os.path.join(tmp_path, "r%d.npy" % r))
        assert np.allclose(got[5:], total, rtol=1e-5)
        if os.path.exists(os.path.join(tmp_path, "gr%d.npy" % r)):
            assert np.allclose(np.load(os.path.join(tmp_path, "gr%d.npy" % r)), total, rtol=1e-5)      # graph replay
            assert np.array_equal(np.load(os.path.join(tmp_path, "gg%d.npy" % r)), np.load(os.path.join(tmp_path, "g%d.npy" % r)))
        else:
            assert os.environ.get("YOLO1_SKIP_GRAPH_NCCL") == "1"
