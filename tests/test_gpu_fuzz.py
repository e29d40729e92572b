"""Randomised sweep of the whole hot path against the oracle: random batch sizes (including empty and ragged),
grid sizes, object densities, layouts, dtypes and thresholds.  Seeds are fixed, so a failure reproduces."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from yolo_v1_b200 import synth

pytestmark = pytest.mark.gpu


def _layout(t, kind):
    if kind == "planar":
        return t.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    if kind == "padded":        # channel-padded storage: strided kernel
        buf = torch.zeros(t.shape[0], t.shape[1], t.shape[2], t.shape[3] + 3, dtype=t.dtype, device=t.device)
        v = buf[..., 1:1 + t.shape[3]]
        v.copy_(t)
        return v
    return t


def test_loss_fuzz():
    import yolo_v1_b200 as y
    rng = np.random.RandomState(20241018)
    for it in range(60):
        S = int(rng.choice([1, 2, 3, 5, 7, 7, 7, 11, 14, 14, 16]))
        N = int(rng.choice([0, 1, 2, 3, 7, 16, 33, 64, 127, 130, 257]))
        p_obj = float(rng.choice([0.0, 0.01, 3.0 / (S * S), 0.3, 1.0]))
        kind = str(rng.choice(["nhwc", "planar", "padded"]))
        variant = "mixed" if rng.rand() < 0.5 else "encoder"
        bs = float(rng.choice([max(N, 1), 12, 64]))
        lc, ln = float(rng.choice([5.0, 1.0, 2.5])), float(rng.choice([0.5, 1.0, 0.1]))
        mode = int(rng.rand() < 0.25)
        pred, target = synth.make_loss_inputs(N, S, p_obj=p_obj, seed=1000 + it, variant=variant)
        o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), l_coord=lc, l_noobj=ln, batch_size=bs, coord_mode=mode)
        pc = _layout(pred.cuda(), kind)
        _, grad, terms = y.yolo_loss_fused(pc, target.cuda(), batch_size=bs, l_coord=lc, l_noobj=ln,
                                           coord_mode="paper" if mode else "reference")
        what = (it, S, N, p_obj, kind, variant, bs, lc, ln, mode)
        t = terms.cpu().numpy()
        assert np.all(np.abs(t - o_terms) <= 1e-5 * np.abs(o_terms) + 1e-7), (what, t, o_terms)
        if N:
            g = grad.cpu().numpy()
            assert np.abs(g - o_grad).max() <= 1e-5 * max(np.abs(o_grad).max(), 1e-12), what
        # the object-list form of the same call (encoder-style targets only: one box shared by both slots)
        if variant == "encoder" and N and kind != "padded":
            cells = (target[..., 0] == 1)
            idx = cells.nonzero()
            bx = target[idx[:, 0], idx[:, 1], idx[:, 2], 2:6]
            cxcy = (bx[:, :2] + torch.stack([idx[:, 2], idx[:, 1]], 1).float()) / S
            boxes = torch.cat([cxcy, bx[:, 2:]], 1).contiguous()
            labels = target[idx[:, 0], idx[:, 1], idx[:, 2], 10:].argmax(1).to(torch.int32) if len(idx) else torch.zeros(0, dtype=torch.int32)
            offs = torch.zeros(N + 1, dtype=torch.int64)
            offs[1:] = cells.reshape(N, -1).sum(1).cumsum(0)
            re_enc = O.encode(boxes.numpy(), labels.numpy(), offs.numpy(), S)
            if np.array_equal(re_enc[..., 0], target[..., 0].numpy()):      # centre round trip landed in the same cells
                o2_terms, o2_grad = O.loss(pred.numpy(), re_enc, l_coord=lc, l_noobj=ln, batch_size=bs, coord_mode=mode)
                _, g2, t2 = y.yolo_loss_from_objects(pc, boxes.cuda(), labels.cuda(), offs.cuda(), batch_size=bs,
                                                     l_coord=lc, l_noobj=ln, coord_mode="paper" if mode else "reference")
                assert np.all(np.abs(t2.cpu().numpy() - o2_terms) <= 1e-5 * np.abs(o2_terms) + 1e-7), what
                assert np.abs(g2.cpu().numpy() - o2_grad).max() <= 1e-5 * max(np.abs(o2_grad).max(), 1e-12), what


def test_decode_nms_fuzz():
    import yolo_v1_b200 as y
    rng = np.random.RandomState(7)
    for it in range(40):
        S = int(rng.choice([1, 2, 3, 5, 7, 7, 9, 14, 16, 22]))
        N = int(rng.choice([1, 2, 5, 17, 64, 100]))
        dist = str(rng.choice(["uniform", "sigmoid"]))
        th = float(rng.choice([0.005, 0.1, 0.3, 0.6, 0.99]))
        nth = float(rng.choice([0.0, 0.25, 0.45, 0.5, 0.9, 1.0]))
        per_class = bool(rng.rand() < 0.3)
        kind = str(rng.choice(["nhwc", "planar", "padded"]))
        pred, _ = synth.make_tie_free_decode_inputs(N, S, seed=500 + it, dist=dist)
        orc = O.decode_nms(pred.numpy(), thresh=th, nms_th=nth, per_class=per_class)
        got = y.decode_nms_batched(_layout(pred.cuda(), kind), th, nth, class_agnostic=not per_class, return_keep=True)
        what = (it, S, N, dist, th, nth, per_class, kind)
        assert np.array_equal(got[3].cpu().numpy(), orc["counts"]), what
        assert np.array_equal(got[5].cpu().numpy(), orc["cand_counts"]), what
        assert np.array_equal(got[4].cpu().numpy(), orc["keep_idx"]), what
        assert np.array_equal(got[0].cpu().numpy().view(np.uint32), orc["boxes"].view(np.uint32)), what
        assert np.array_equal(got[2].cpu().numpy().view(np.uint32), orc["scores"].view(np.uint32)), what
        assert np.array_equal(got[1].cpu().numpy(), orc["cls"]), what


def test_nms_fuzz_with_ties_and_duplicates():
    """Raw box sets through yolo1_nms: every size from tiny to past one CTA's thread count, heavily tied scores (the
    rank-checksum path re-ranks those sets with the tie rule), duplicated boxes, class-aware and class-agnostic."""
    import yolo_v1_b200 as y
    rng = np.random.RandomState(11)
    sizes = [1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 95, 96, 97, 98, 127, 128, 129, 255, 256, 257, 383, 384, 385, 392, 700]
    for it, n in enumerate(sizes * 2):
        xy = rng.rand(n, 2).astype(np.float32) * 0.7
        wh = (rng.rand(n, 2) * 0.4 + 0.02).astype(np.float32)
        b = np.concatenate([xy, xy + wh], 1)
        if n > 3:
            b[rng.randint(0, n, n // 4)] = b[rng.randint(0, n, n // 4)]        # exact duplicates (IoU = 1)
        levels = int(rng.choice([1, 2, 5, 50, 10 ** 6]))
        s = (rng.randint(0, levels, n) / float(levels)).astype(np.float32) + np.float32(0.01)
        cls = rng.randint(0, 3, n).astype(np.int32)
        thr = float(rng.choice([0.0, 0.3, 0.5, 0.75, 1.0]))
        per_class = bool(it % 3 == 0)
        M = n + int(rng.randint(0, 5))                                           # padded rows beyond the count
        bb, ss, cc = np.zeros((1, M, 4), np.float32), np.zeros((1, M), np.float32), np.zeros((1, M), np.int32)
        bb[0, :n], ss[0, :n], cc[0, :n] = b, s, cls
        keep, kc = y.nms_batched(torch.from_numpy(bb).cuda(), torch.from_numpy(ss).cuda(),
                                 torch.tensor([n], dtype=torch.int32), thr, cls=torch.from_numpy(cc).cuda(),
                                 per_class=per_class)
        want = O.nms(b, s, thr, cls=cls if per_class else None, per_class=per_class)
        what = (it, n, levels, thr, per_class)
        assert int(kc[0]) == len(want), what
        assert np.array_equal(keep[0, :len(want)].cpu().numpy(), want), what
