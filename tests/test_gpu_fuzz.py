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


def test_streaming_kernels_soak_with_poisoned_outputs():
    """compute-sanitizer is closed on this pool (profiles/sanitizer_r2.log), so memory safety of the bulk-copy
    pipelines is soaked instead: 10 000 randomised launches over every streaming kernel form --
    NHWC launch shapes, small-call cluster kernels, sector-read kernels, planar kernels (dense, warp-specialised,
    confidence-first), object lists, fused head, bf16, forward only, the target encoder and decode+NMS -- with sizes
    that hit every tile / tail / grid-stride boundary.  Every output buffer sits between NaN guard bands and is itself
    NaN-poisoned before the launch: after it the guards must be untouched and no NaN may be left inside, and every
    launch must agree with the strided one-thread-per-cell kernel on the same input (a sample of them with the
    oracle).  The inputs are drawn once per shape class and re-sliced, so the loop is launch-bound."""
    import yolo_v1_b200 as y
    rng = np.random.RandomState(99)
    G = 1024                                   # guard band, floats (16-byte multiples keep the fast paths eligible)
    pools = {}

    def pool(S):
        if S not in pools:
            n = {3: 4096, 7: 2048, 14: 512}[S]
            pred, target = synth.make_loss_inputs(n, S, seed=7000 + S, p_obj=3.0 / (S * S))
            dense_t = synth.make_loss_inputs(n, S, seed=7100 + S, p_obj=0.3, variant="mixed")[1]
            pools[S] = (pred.cuda(), target.cuda(), dense_t.cuda(), pred, target, dense_t)
        return pools[S]

    def guarded(shape, dtype, planar):
        N, S, _, D = shape
        numel = N * S * S * D
        buf = torch.full((numel + 2 * G,), float("nan"), dtype=dtype, device="cuda")
        inner = buf[G:G + numel]
        view = inner.view(N, D, S, S).permute(0, 2, 3, 1) if planar else inner.view(N, S, S, D)
        return buf, view

    forms = ([("nhwc", v) for v in (1, 2, 3, 5, 8, 13, 31, 0, 40, 41, 42)] +
             [("planar", v) for v in (0, 1, 20, 31, 50, 51)] + [("lists", 0), ("lists_planar", 31), ("logits", 0),
                                                               ("logits_planar", 0), ("encode", 0), ("decode", 0)])
    checked = 0
    for it in range(10000):
        form, variant = forms[it % len(forms)]
        S = int(rng.choice([7, 14, 14, 3]))
        cap = {3: 4096, 7: 2048, 14: 512}[S]
        N = int(rng.choice([1, 2, 3, int(rng.randint(1, 40)), int(rng.randint(40, cap + 1))]))
        lo = int(rng.randint(0, cap - N + 1))
        pc, tc, tdense, ph, th, tdh = pool(S)
        bf16 = form in ("nhwc", "planar") and rng.rand() < 0.25 and variant not in (40, 41, 42)
        want_grad = rng.rand() < 0.85
        tt = tdense if rng.rand() < 0.3 else tc
        pred = pc[lo:lo + N]
        target = tt[lo:lo + N]
        if form == "decode":
            M = S * S * 2
            bb = torch.full((N * M * 4 + 2 * G,), float("nan"), device="cuda")
            out = (bb[G:G + N * M * 4].view(N, M, 4), torch.full((N, M), -7, dtype=torch.int32, device="cuda"),
                   torch.full((N, M), float("nan"), device="cuda"), torch.full((N,), -7, dtype=torch.int32, device="cuda"))
            y.decode_nms_batched(pred, 0.1, 0.5, out=out)
            assert bool(torch.isnan(bb[:G]).all()) and bool(torch.isnan(bb[-G:]).all()), (it, "decode guards")
            assert not bool(torch.isnan(bb[G:-G]).any()) and not bool(torch.isnan(out[2]).any()), (it, "decode poison left")
            assert int(out[3].min()) >= 0 and int(out[1].min()) >= 0
            continue
        if form == "encode":
            k = int(rng.randint(0, 6))
            boxes = torch.rand(N * k, 4, device="cuda") * 0.98 + 0.01
            labels = torch.randint(0, 20, (N * k,), device="cuda", dtype=torch.int32)
            offs = torch.arange(0, N * k + 1, max(k, 1), device="cuda", dtype=torch.int64)[:N + 1] if k else \
                torch.zeros(N + 1, dtype=torch.int64, device="cuda")
            t1 = y.encode_targets(boxes, labels, offs, S)
            t2 = y.encode_targets(boxes, labels, offs, S)
            assert torch.equal(t1, t2) and not bool(torch.isnan(t1).any()), (it, "encode")
            assert int((t1[..., 0] == 1).sum()) <= N * k
            continue
        planar = form in ("planar", "lists_planar", "logits_planar")
        p_in = pred.to(torch.bfloat16) if bf16 else pred
        if planar:
            p_in = p_in.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
        buf, gview = guarded(pred.shape, p_in.dtype, planar)
        kw = dict(batch_size=N, want_grad=want_grad, out_grad=gview if want_grad else None)
        if form.startswith("lists"):
            objmask = tc[lo:lo + N, ..., 0] == 1          # encoder-style targets only
            idx = objmask.nonzero()
            bx = tc[lo:lo + N][idx[:, 0], idx[:, 1], idx[:, 2], 2:6]
            cxcy = (bx[:, :2] + torch.stack([idx[:, 2], idx[:, 1]], 1).float()) / S
            boxes = torch.cat([cxcy, bx[:, 2:]], 1).contiguous()
            labels = tc[lo:lo + N][idx[:, 0], idx[:, 1], idx[:, 2], 10:].argmax(1).to(torch.int32)
            offs = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
            offs[1:] = objmask.reshape(N, -1).sum(1).cumsum(0)
            _, grad, terms = y.yolo_loss_from_objects(p_in, boxes, labels, offs, variant=variant, **kw)
            target = y.encode_targets(boxes, labels, offs, S)
            ref_kw = {}
        elif form.startswith("logits"):
            _, grad, terms = y.yolo_loss_fused(p_in, target, from_logits=True, **kw)
            ref_kw = dict(from_logits=True)
        else:
            try:
                _, grad, terms = y.yolo_loss_fused(p_in, target, variant=variant, **kw)
            except RuntimeError:
                # a forced kernel form that the call is not eligible for (the sector-read kernels want 16-byte aligned
                # fp32 NHWC: an odd slice offset of a 7x7 pool is 8 bytes off; 50 wants whole-image planar tiles)
                assert variant in (40, 41, 42, 50), (it, form, variant)
                _, grad, terms = y.yolo_loss_fused(p_in, target, variant=0, **kw)
            ref_kw = {}
        _, g_ref, t_ref = y.yolo_loss_fused(p_in, target, batch_size=N, variant=-1, **ref_kw)
        what = (it, form, variant, S, N, lo, bf16, want_grad)
        assert bool(torch.isnan(buf[:G]).all()) and bool(torch.isnan(buf[-G:]).all()), (what, "guard band written")
        assert torch.allclose(terms, t_ref, rtol=2e-6, atol=1e-7), (what, terms, t_ref)
        if want_grad:
            assert not bool(torch.isnan(buf[G:-G]).any()), (what, "poison left in the gradient")
            tol = 2.0 ** -7 if bf16 else 2e-6
            scale = float(g_ref.float().abs().max()) + 1e-12
            assert float((gview.float() - g_ref.float()).abs().max()) <= tol * scale, what
        else:
            assert bool(torch.isnan(buf).all()), (what, "forward-only call wrote a gradient")
        if it % 200 == 0 and not bf16 and not form.startswith("logits"):
            src_t = target.cpu().numpy()
            o_terms, o_grad = O.loss(ph[lo:lo + N].numpy(), src_t, batch_size=N)
            assert np.all(np.abs(terms.cpu().numpy() - o_terms) <= 1e-5 * np.abs(o_terms) + 1e-7), what
            checked += 1
    assert checked >= 20
