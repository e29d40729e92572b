"""K2/K3 parity: CUDA decode + NMS (through the C ABI / the reference-shaped functions) against
  (1) the golden vectors produced by the reference's own decoder/nms (tests/golden/decode_cases.npz),
  (2) the CPU oracle on seeded tie-free synthetic inputs (BASELINE config 2 in full),
  (3) size-independent properties on large batches.
Bar: bit-exact -- boxes and scores compared as uint32 bit patterns, classes, counts and keep lists as
integers.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from yolo_v1_b200 import synth

pytestmark = pytest.mark.gpu


def _y():
    import yolo_v1_b200 as y
    return y


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _assert_batched_equal(got, orc, what):
    boxes, cls, probs, counts = [t.cpu().numpy() for t in got[:4]]
    assert np.array_equal(counts, orc["counts"]), what
    assert np.array_equal(_bits(boxes), _bits(orc["boxes"])), what     # padding rows are zero on both sides
    assert np.array_equal(_bits(probs), _bits(orc["scores"])), what
    assert np.array_equal(cls, orc["cls"]), what
    if len(got) > 4:
        assert np.array_equal(got[4].cpu().numpy(), orc["keep_idx"]), what
        assert np.array_equal(got[5].cpu().numpy(), orc["cand_counts"]), what


def test_golden_decoder_cases_from_the_reference(golden_dir):
    y = _y()
    z = np.load(os.path.join(golden_dir, "decode_cases.npz"))
    sets = sorted({k.split("/")[0] for k in z.files if k.endswith("/pred")})
    assert len(sets) >= 8
    for name in sets:
        S, th, nth, gt = z[name + "/params"]
        pred, counts = z[name + "/pred"], z[name + "/counts"]
        off = 0
        for n in range(pred.shape[0]):
            dev = 'cpu' if n % 2 else 'cuda'
            src = torch.from_numpy(pred[n:n + 1].copy())
            b, c, s = y.decoder(src.cuda() if n % 3 == 0 else src, grid_num=int(S), device=dev,
                                thresh=float(th), nms_th=float(nth), gt=bool(gt))
            assert np.array_equal(_bits(src.numpy()), _bits(pred[n:n + 1]))   # the caller's tensor is left intact (NaN-safe)
            k = int(counts[n])
            rb, rc, rs = z[name + "/boxes"][off:off + k], z[name + "/cls"][off:off + k], z[name + "/probs"][off:off + k]
            off += k
            b, s = b.cpu().numpy(), s.cpu().numpy()
            c = c.cpu().numpy().astype(np.float32)
            if gt:   # every score ties at 1.0: order unspecified upstream, compare canonically
                key = np.lexsort((c, b[:, 3], b[:, 2], b[:, 1], b[:, 0], -s))
                b, c, s = b[key], c[key], s[key]
            assert len(s) == k, (name, n)
            assert np.array_equal(_bits(b), _bits(rb)), (name, n)
            assert np.array_equal(c, rc), (name, n)
            assert np.array_equal(_bits(s), _bits(rs)), (name, n)


def test_golden_nms_cases_from_the_reference(golden_dir):
    y = _y()
    z = np.load(os.path.join(golden_dir, "decode_cases.npz"))
    names = sorted({k.split("/")[0] for k in z.files if k.endswith("/keep")})
    assert "nms_chain" in names and len(names) >= 4
    for name in names:
        keep = y.nms(torch.from_numpy(z[name + "/boxes"]), torch.from_numpy(z[name + "/scores"]),
                     float(z[name + "/thr"]))
        assert keep.dtype == torch.long and not keep.is_cuda
        assert np.array_equal(keep.numpy(), z[name + "/keep"]), name
    assert y.nms(torch.zeros(0, 4), torch.zeros(0)).numel() == 0


@pytest.mark.parametrize("S,N,dist,th,nth", [
    (7, 4096, "uniform", 0.1, 0.5),      # BASELINE config 2 exactly (eval.py:94 thresholds)
    (7, 512, "uniform", 0.005, 0.45),    # run_test_mAP thresholds (utils/utils.py:405)
    (7, 512, "sigmoid", 0.1, 0.5),
    (14, 256, "uniform", 0.1, 0.5),
    (14, 256, "sigmoid", 0.005, 0.45),
    (3, 64, "uniform", 0.3, 0.5),
])
def test_batched_decode_nms_bit_exact_vs_oracle(S, N, dist, th, nth):
    y = _y()
    pred, redrawn = synth.make_tie_free_decode_inputs(N, S, seed=2 + S, dist=dist)
    orc = O.decode_nms(pred.numpy(), thresh=th, nms_th=nth)
    got = y.decode_nms_batched(pred.cuda(), th, nth, return_keep=True)
    _assert_batched_equal(got, orc, (S, N, dist, th, nth, "redrawn=%d" % redrawn))
    # two-kernel path (yolo1_decode + yolo1_nms) gives the same keep lists as the fused kernel
    boxes, scores, cls, counts = y.decode_batched(pred.cuda(), th)
    assert np.array_equal(counts.cpu().numpy(), orc["cand_counts"])
    keep, kc = y.nms_batched(boxes, scores, counts, nth)
    assert np.array_equal(kc.cpu().numpy(), orc["counts"])
    assert np.array_equal(keep.cpu().numpy(), orc["keep_idx"])
    # per-class mode (north_star wording): equals the oracle's per-class mode
    orc_pc = O.decode_nms(pred.numpy(), thresh=th, nms_th=nth, per_class=True)
    got_pc = y.decode_nms_batched(pred.cuda(), th, nth, class_agnostic=False, return_keep=True)
    _assert_batched_equal(got_pc, orc_pc, (S, N, dist, "per-class"))


def test_strided_and_bf16_inputs():
    y = _y()
    pred, _ = synth.make_tie_free_decode_inputs(64, 7, seed=31)
    orc = O.decode_nms(pred.numpy(), thresh=0.1, nms_th=0.5)
    planar = pred.cuda().permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)   # backbone's NCHW view
    _assert_batched_equal(y.decode_nms_batched(planar, 0.1, 0.5, return_keep=True), orc, "planar")
    pb = pred.to(torch.bfloat16)
    orc_b = O.decode_nms(pb.float().numpy(), thresh=0.1, nms_th=0.5)
    got = y.decode_nms_batched(pb.cuda(), 0.1, 0.5)    # bf16 rounding can create score ties: canonical order on both sides
    assert np.array_equal(got[3].cpu().numpy(), orc_b["counts"])
    assert np.array_equal(_bits(got[0].cpu().numpy()), _bits(orc_b["boxes"]))


def test_sentinel_empty_and_edge_cases():
    y = _y()
    # nothing passes: the reference's all-zero sentinel (utils/utils.py:134-137)
    pred = torch.full((1, 7, 7, 30), 0.01)
    b, c, s = y.decoder(pred, thresh=0.3)
    assert b.shape == (1, 4) and c.shape == (1,) and s.shape == (1,)
    assert float(b.abs().sum()) == 0 and float(c.sum()) == 0 and float(s.sum()) == 0
    # conf <= 1e-4 but equal to the image max is still a candidate (:111-113)
    pred = torch.zeros(1, 7, 7, 30)
    pred[0, 3, 4, 0] = 5e-5
    pred[0, 3, 4, 2:6] = torch.tensor([0.5, 0.5, 0.2, 0.2])
    pred[0, 3, 4, 17] = 1.0
    ob, oc, os_ = O.decoder(pred.numpy(), thresh=1e-5)
    b, c, s = y.decoder(pred, thresh=1e-5)
    assert len(os_) == 1 and np.array_equal(_bits(b.numpy()), _bits(ob)) and int(c[0]) == 7 == int(oc[0])
    assert np.array_equal(_bits(s.numpy()), _bits(os_))
    # all-zero image: every slot equals the max (0) but 0 > thresh fails -> sentinel
    b, c, s = y.decoder(torch.zeros(1, 7, 7, 30), thresh=0.0)
    assert b.shape == (1, 4) and float(s.sum()) == 0
    # batched: an empty image reports count 0, others unaffected
    batch, _ = synth.make_tie_free_decode_inputs(8, 7, seed=5)
    batch[3] = 0.01
    orc = O.decode_nms(batch.numpy(), thresh=0.1, nms_th=0.5)
    assert orc["counts"][3] == 0
    _assert_batched_equal(y.decode_nms_batched(batch.cuda(), 0.1, 0.5, return_keep=True), orc, "empty image")
    # N = 0
    out = y.decode_nms_batched(torch.zeros(0, 7, 7, 30, device="cuda"), 0.1, 0.5)
    assert out[3].numel() == 0


def test_nms_semantics():
    y = _y()
    # chain A > B > C: A kills B, B would have killed C; C must survive (a suppressed box does not suppress)
    boxes = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.4, 0.0, 1.4, 1.0], [0.8, 0.0, 1.8, 1.0]])
    scores = torch.tensor([0.9, 0.8, 0.7])
    assert y.nms(boxes, scores, 0.4).tolist() == [0, 2] == O.nms(boxes.numpy(), scores.numpy(), 0.4).tolist()
    # exactly one survivor in a round (the case that crashes the reference on torch >= 0.5)
    boxes = torch.tensor([[0, 0, 1, 1], [0, 0, 1, 1.01], [2, 2, 3, 3]], dtype=torch.float32)
    scores = torch.tensor([0.5, 0.9, 0.1])
    assert y.nms(boxes, scores, 0.5).tolist() == [1, 2]
    # IoU exactly equal to the threshold survives (ovr <= threshold, utils/utils.py:180)
    boxes = torch.tensor([[0, 0, 2, 1], [1, 0, 3, 1]], dtype=torch.float32)   # IoU = 1/3
    thr = float(np.float32(1.0) / np.float32(3.0))
    assert y.nms(boxes, torch.tensor([0.9, 0.8]), thr).tolist() == [0, 1]
    assert y.nms(boxes, torch.tensor([0.9, 0.8]), thr - 1e-6).tolist() == [0]
    # degenerate boxes: 0/0 IoU is NaN -> the later box dies (NaN <= thr is false)
    boxes = torch.zeros(2, 4)
    assert y.nms(boxes, torch.tensor([0.9, 0.8]), 0.5).tolist() == O.nms(boxes.numpy(), np.array([0.9, 0.8]), 0.5).tolist() == [0]
    # ties -> lower index first (canonical order)
    boxes = torch.tensor([[0, 0, 1, 1], [5, 5, 6, 6], [9, 9, 10, 10]], dtype=torch.float32)
    assert y.nms(boxes, torch.tensor([0.5, 0.5, 0.5]), 0.5).tolist() == [0, 1, 2]
    # large random sets up to the 1024-box limit, against the oracle
    g = torch.Generator().manual_seed(0)
    for n in (1, 31, 32, 33, 98, 392, 1000, 1024):
        xy = torch.rand(n, 2, generator=g) * 0.8
        wh = torch.rand(n, 2, generator=g) * 0.3 + 0.01
        b = torch.cat([xy, xy + wh], 1)
        s = torch.randperm(n, generator=g).float() / n
        for thr in (0.25, 0.5):
            assert np.array_equal(y.nms(b, s, thr).numpy(), O.nms(b.numpy(), s.numpy(), thr)), (n, thr)


def test_nms_long_kill_chains():
    """Box i overlaps box i+1 above the threshold and box i+2 below it, scores descending: the greedy loop keeps
    0, 2, 4, ... -- a kill chain as deep as the set.  The sweep re-evaluates all boxes per pass ("sweep as
    a fixed point") and settles about two per pass here, so this is its worst case: every size class of the kernels
    (one-warp sweep up to 128 boxes, CTA-wide above), through the raw call and through the fused decode."""
    y = _y()
    for n in (2, 3, 33, 64, 97, 98, 128, 129, 392, 1024):
        x0 = np.arange(n, dtype=np.float32) * np.float32(0.25)          # IoU(i, i+1) = 0.6, IoU(i, i+2) = 1/3
        b = np.stack([x0, np.zeros(n, np.float32), x0 + 1, np.ones(n, np.float32)], 1)
        s = (1.0 - np.arange(n) / (2.0 * n)).astype(np.float32)
        want = O.nms(b, s, 0.5)
        assert want.tolist() == list(range(0, n, 2))
        assert y.nms(torch.from_numpy(b), torch.from_numpy(s), 0.5).tolist() == want.tolist(), n
        # chains that break and restart: every 7th box is far away
        b2 = b.copy()
        b2[::7, 1] += 5
        b2[::7, 3] += 5
        assert np.array_equal(y.nms(torch.from_numpy(b2), torch.from_numpy(s), 0.5).numpy(), O.nms(b2, s, 0.5)), n
    # the same chain inside a decoded image: 49 cells x 2 slots, all boxes on one row of overlapping squares
    S = 7
    pred = torch.zeros(1, S, S, 30)
    k = 0
    for i in range(S):
        for j in range(S):
            for bslot in range(2):
                pred[0, i, j, bslot] = 0.99 - 0.005 * k                  # descending confidences
                cx = 0.05 + 0.004 * k                                    # centres 0.004 apart, width 0.016: IoU 0.6 / 0.33
                pred[0, i, j, 2 + 4 * bslot] = (cx * S - j)
                pred[0, i, j, 3 + 4 * bslot] = (0.5 * S - i)
                pred[0, i, j, 4 + 4 * bslot] = 0.016
                pred[0, i, j, 5 + 4 * bslot] = 0.5
                k += 1
            pred[0, i, j, 10] = 1.0
    orc = O.decode_nms(pred.numpy(), thresh=0.1, nms_th=0.5)
    assert 40 <= int(orc["counts"][0]) <= 60
    _assert_batched_equal(y.decode_nms_batched(pred.cuda(), 0.1, 0.5, return_keep=True), orc, "chain image")


def test_nms_threshold_boundaries():
    """`ovr <= threshold` (utils/utils.py:180) is evaluated without the division in the kernel (iou_exceeds): pairs
    whose fp32 quotient is exactly the threshold, one ulp above and one ulp below must fall on the right side.
    Expected values: the reference's op sequence in numpy float32 (IEEE division)."""
    y = _y()
    rng = np.random.default_rng(7)
    n = 120000
    a = rng.integers(2, 4096, n).astype(np.float32)
    c = np.floor(rng.random(n) * a).astype(np.float32)              # 0 <= c < a: the two boxes overlap
    b = rng.integers(1, 4096, n).astype(np.float32)
    h = (rng.integers(1, 64, n) * 2.0 ** -6).astype(np.float32)
    sc = np.float32(2.0 ** -12)
    boxes = np.zeros((n, 2, 4), np.float32)
    boxes[:, 0, 2], boxes[:, 0, 3] = a * sc, h
    boxes[:, 1, 0], boxes[:, 1, 2], boxes[:, 1, 3] = c * sc, (c + b) * sc, h
    scores = np.tile(np.array([0.9, 0.8], np.float32), (n, 1))
    A, Bx = boxes[:, 0], boxes[:, 1]
    area = lambda q: (q[:, 2] - q[:, 0]) * (q[:, 3] - q[:, 1])
    w = np.maximum(np.minimum(Bx[:, 2], A[:, 2]) - np.maximum(Bx[:, 0], A[:, 0]), np.float32(0))
    hh = np.maximum(np.minimum(Bx[:, 3], A[:, 3]) - np.maximum(Bx[:, 1], A[:, 1]), np.float32(0))
    inter = w * hh
    ovr = inter / ((area(A) + area(Bx)) - inter)
    assert ovr.dtype == np.float32 and np.isfinite(ovr).all()
    bd, sd = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    counts = torch.full((n,), 2, dtype=torch.int32, device="cuda")
    picks = ovr[rng.integers(0, n, 12)]
    thrs = [0.5, 0.45, 0.25, float(np.float32(1) / np.float32(3)), 0.0, 1.0]
    for v in picks:
        thrs += [float(v), float(np.nextafter(v, np.float32(2))), float(np.nextafter(v, np.float32(-1)))]
    ties = 0
    for thr in thrs:
        t32 = np.float32(thr)
        want = np.where(ovr <= t32, 2, 1)
        ties += int((ovr == t32).sum())
        _, kc = y.nms_batched(bd, sd, counts, thr)
        assert np.array_equal(kc.cpu().numpy(), want), thr
    assert ties > 50   # the equality case is exercised


def test_nms_nan_and_inf_boxes():
    """NaN / infinite coordinates take the kernel's non-finite path; same keep lists as the oracle."""
    y = _y()
    g = torch.Generator().manual_seed(3)
    for n in (2, 7, 40, 98):
        xy = torch.rand(n, 2, generator=g) * 0.8
        b = torch.cat([xy, xy + torch.rand(n, 2, generator=g) * 0.3 + 0.01], 1)
        s = torch.randperm(n, generator=g).float() / n
        for bad in (float("nan"), float("inf"), -float("inf"), 3.0e38):
            bb = b.clone()
            bb[n // 2, 2] = bad
            bb[0, 1] = bad
            if n > 7:
                bb[5] = bad
            assert np.array_equal(y.nms(bb, s, 0.5).numpy(), O.nms(bb.numpy(), s.numpy(), 0.5)), (n, bad)


def test_helpers_match_reference_fixtures(golden_dir):
    import json
    y = _y()
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    fx = meta["iou_fixture"]    # utils/utils.py:506-525
    iou = y.compute_iou_matrix(torch.tensor(fx["b1"]).cuda(), torch.tensor(fx["b2"]).cuda())
    assert np.allclose(iou.cpu().numpy(), np.array(fx["iou"], np.float32), rtol=1e-6)
    cf = meta["convert_fixture"]
    out = y.convert_CxCyWH_to_X1Y1X2Y2(torch.tensor(cf["boxes"]), cf["S"])
    assert np.allclose(out.numpy(), np.array(cf["out"], np.float32), rtol=1e-6)


def test_host_buffer_path_equals_device_path():
    y = _y()
    pred, _ = synth.make_tie_free_decode_inputs(300, 7, seed=8)
    orc = O.decode_nms(pred.numpy(), thresh=0.1, nms_th=0.5)
    for chunk in (0, 7, 300):
        ctx = y.HostContext(7, chunk_images=chunk)
        out = ctx.decode_nms(pred, 0.1, 0.5)
        got = (out["boxes"], out["cls"], out["scores"], out["counts"])
        _assert_batched_equal(got, orc, ("host", chunk))
        # pinned input and outputs (full-rate copies), with either setting of the loss-only zero-copy switch
        pp = pred.pin_memory()
        for zc in (2, 0):
            ctx.set_zero_copy(zc)
            pin = dict(boxes=torch.full((300, 98, 4), -1.0).pin_memory(), scores=torch.full((300, 98), -1.0).pin_memory(),
                       cls=torch.full((300, 98), -1, dtype=torch.int32).pin_memory(),
                       counts=torch.full((300,), -1, dtype=torch.int32).pin_memory())
            out = ctx.decode_nms(pp, 0.1, 0.5, out=pin)
            _assert_batched_equal((out["boxes"], out["cls"], out["scores"], out["counts"]), orc, ("host pinned", chunk, zc))
        ctx.close()


def test_large_batch_properties():
    """65536 images (16 copies of a 4096-image tie-free block): every copy must reproduce the block's
    oracle-checked result bit for bit (images are independent), and detections are sorted and within counts."""
    y = _y()
    block, _ = synth.make_tie_free_decode_inputs(4096, 7, seed=2)
    orc = O.decode_nms(block.numpy(), thresh=0.1, nms_th=0.5)
    big = block.cuda().repeat(16, 1, 1, 1)
    boxes, cls, probs, counts = y.decode_nms_batched(big, 0.1, 0.5)
    ref = [torch.from_numpy(orc[k]).cuda() for k in ("boxes", "cls", "scores", "counts")]
    for r in range(16):
        sl = slice(r * 4096, (r + 1) * 4096)
        assert torch.equal(counts[sl], ref[3]) and torch.equal(boxes[sl], ref[0])
        assert torch.equal(probs[sl], ref[2]) and torch.equal(cls[sl], ref[1])
    M = probs.shape[1]
    valid = torch.arange(M, device="cuda")[None, :] < counts[:, None]
    d = probs[:, 1:] - probs[:, :-1]
    assert bool(((d <= 0) | ~valid[:, 1:]).all())          # descending scores inside the count
    assert bool((probs[~valid] == 0).all())
