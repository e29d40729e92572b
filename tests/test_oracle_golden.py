"""The CPU oracle (oracle/yolo1_oracle.c) against the golden vectors produced by the reference's own
code (oracle/make_golden.py).  This is the pin that makes the oracle trustworthy as the checker of the
CUDA path.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

LOSS_TOL = 2e-6      # oracle vs reference (fp32 autograd) -- the GPU gate itself is 1e-5


def _npz(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _loss_case_names(golden_dir):
    z = _npz(golden_dir, "loss_cases.npz")
    return sorted({k.split("/")[0] for k in z.files})


def test_loss_cases_match_reference(golden_dir):
    z = _npz(golden_dir, "loss_cases.npz")
    names = _loss_case_names(golden_dir)
    assert len(names) >= 10
    for name in names:
        S, B, C, lc, ln, bs = z[name + "/hyper"]
        terms, grad = O.loss(z[name + "/pred"], z[name + "/target"], int(B), int(C), lc, ln, bs)
        ref_loss, ref_grad = float(z[name + "/loss"]), z[name + "/grad"]
        assert abs(float(terms[4]) - ref_loss) <= LOSS_TOL * max(abs(ref_loss), 1e-12), name
        assert np.abs(grad - ref_grad).max() <= LOSS_TOL * max(np.abs(ref_grad).max(), 1e-12), name
        # components recombine into the total (v1Loss.py:104-105)
        tot = lc * terms[0] + terms[1] + ln * terms[2] + terms[3]
        assert abs(tot - terms[4]) <= 1e-5 * abs(terms[4]) + 1e-7, name


def test_loss_permuted_nchw_view_gives_same_result(golden_dir):
    """The backbone hands the loss a permuted NCHW view (OriginResNet.py:189); strides must not matter."""
    z = _npz(golden_dir, "loss_cases.npz")
    pred, target = z["rand_s7_n8/pred"], z["rand_s7_n8/target"]
    planar = np.ascontiguousarray(pred.transpose(0, 3, 1, 2)).transpose(0, 2, 3, 1)
    assert planar.strides != pred.strides
    t1, g1 = O.loss(pred, target)
    t2, g2 = O.loss(planar, target)
    assert np.array_equal(t1, t2)
    assert g2.strides == planar.strides and np.array_equal(g1, g2)


def test_loss_paper_mode_differs_and_quirk_boundary(golden_dir):
    z = _npz(golden_dir, "loss_cases.npz")
    pred, target = z["rand_s7_n8/pred"], z["rand_s7_n8/target"]
    t_ref, _ = O.loss(pred, target, coord_mode=0)
    t_pap, _ = O.loss(pred, target, coord_mode=1)
    assert abs(t_ref[0] - t_pap[0]) > 1e-3          # the row-slice quirk is not the paper formula
    assert np.allclose(t_ref[1:4], t_pap[1:4])      # only the location term changes


def test_decode_sets_bit_exact(golden_dir):
    z = _npz(golden_dir, "decode_cases.npz")
    sets = sorted({k.split("/")[0] for k in z.files if k.endswith("/pred")})
    assert len(sets) >= 8
    for name in sets:
        S, th, nth, gt = z[name + "/params"]
        pred, counts = z[name + "/pred"], z[name + "/counts"]
        off = 0
        for n in range(pred.shape[0]):
            b, c, s = O.decoder(pred[n], grid_num=int(S), thresh=th, nms_th=nth, gt=bool(gt))
            k = int(counts[n])
            rb, rc, rs = z[name + "/boxes"][off:off + k], z[name + "/cls"][off:off + k], z[name + "/probs"][off:off + k]
            off += k
            c = np.asarray(c, np.float32)
            if gt:   # all scores tie at 1.0: order unspecified upstream, compare canonically
                key = np.lexsort((c, b[:, 3], b[:, 2], b[:, 1], b[:, 0], -s))
                b, c, s = b[key], c[key], s[key]
            assert len(s) == k, (name, n)
            assert np.array_equal(b.view(np.uint32), rb.view(np.uint32)), (name, n)
            assert np.array_equal(c, rc), (name, n)
            assert np.array_equal(s.view(np.uint32), rs.view(np.uint32)), (name, n)


def test_batched_decode_nms_equals_per_image(golden_dir):
    z = _npz(golden_dir, "decode_cases.npz")
    pred = z["uni_s7/pred"]
    out = O.decode_nms(pred, thresh=0.1, nms_th=0.5)
    assert np.array_equal(out["counts"], z["uni_s7/counts"])
    off = 0
    for n in range(pred.shape[0]):
        k = int(out["counts"][n])
        assert np.array_equal(out["boxes"][n, :k], z["uni_s7/boxes"][off:off + k])
        assert np.array_equal(out["scores"][n, :k], z["uni_s7/probs"][off:off + k])
        assert np.array_equal(out["cls"][n, :k].astype(np.float32), z["uni_s7/cls"][off:off + k])
        off += k


def test_nms_cases(golden_dir):
    z = _npz(golden_dir, "decode_cases.npz")
    names = sorted({k.split("/")[0] for k in z.files if k.endswith("/keep")})
    assert "nms_chain" in names and len(names) >= 4
    for name in names:
        keep = O.nms(z[name + "/boxes"], z[name + "/scores"], float(z[name + "/thr"]))
        assert np.array_equal(keep, z[name + "/keep"]), name
    assert list(z["nms_chain/keep"]) == [0, 2, 3]   # a suppressed box does not suppress


def test_per_class_nms_equals_reference_nms_per_class_subset(golden_dir):
    """per_class=True must equal: run the (class-agnostic) nms once per class subset, merge by score."""
    z = _npz(golden_dir, "decode_cases.npz")
    pred = z["uni_s7/pred"]
    for n in range(4):
        b, s, c = O.decode_image(pred[n], thresh=0.1)
        merged = []
        for k in np.unique(c):
            idx = np.nonzero(c == k)[0]
            merged += [int(idx[i]) for i in O.nms(b[idx], s[idx], 0.5)]
        merged.sort(key=lambda i: (-s[i], i))
        assert merged == list(O.nms(b, s, 0.5, cls=c, per_class=True))


def test_reference_print_only_fixtures(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    fx = meta["iou_fixture"]   # utils/utils.py:506-525
    iou = O.iou_matrix(np.array(fx["b1"], np.float32), np.array(fx["b2"], np.float32))
    assert np.array_equal(iou, np.array(fx["iou"], np.float32))
    assert np.allclose(iou, [[0.24449877, 0.53832752, 0.0], [0.0, 0.0, 0.17006803]], atol=1e-7)
    cf = meta["convert_fixture"]   # utils/utils.py:59-75
    out = O.cxcywh_to_xyxy(np.array(cf["boxes"], np.float32), cf["S"])
    assert np.array_equal(out, np.array(cf["out"], np.float32))
    assert meta["voc_eval_fixture"]["mAP"] == pytest.approx(0.9166666666666666, abs=0)
    assert meta["loss_oracle_vs_reference_worst_rel"] < LOSS_TOL
