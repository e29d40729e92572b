"""Generate tests/golden/* by running the REAL reference (/root/reference, torch CPU) on seeded inputs.

Run in the build container only (the reference mount does not exist on the GPU box):

    python oracle/make_golden.py            # writes tests/golden/*.npz + golden_meta.json

What runs is the reference's own code: `YOLOLossV1.forward` + autograd (`v1Loss.py:22-118`),
`decoder`/`nms` (`utils/utils.py:94-184`, with the one-token torch>=0.5 shim described in
oracle/ref_loader.py), `compute_iou_matrix` (`utils/utils.py:10-57`),
`convert_CxCyWH_to_X1Y1X2Y2` (`utils/utils.py:59-75`), `voc_eval` (`utils/utils.py:240-319`) and
`run_test_mAP` (`utils/utils.py:389-418`, driven with an identity "network").
The script also cross-checks oracle/yolo1_oracle.c against every vector it writes and aborts on a
mismatch, so a successful run is itself the oracle pin.
"""
import contextlib
import copy
import io
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle.ref_loader import load_reference, quiet  # noqa: E402
from yolo_v1_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
RefLoss, U = load_reference()


# --------------------------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------------------------
def ref_loss(pred, target, S, B, C, lc, ln, bs):
    p = pred.clone().requires_grad_(True)
    mod = RefLoss(bs, S, B, C, _l_coord=lc, _l_noobj=ln, _device="cpu")
    with quiet():
        out = mod(p, target)
        out.backward()
    return float(out.item()), p.grad.detach().numpy().copy()


def hand_target(S, B, C, cells, boxes=None, cls=None):
    """cells: list of (n,i,j); builds an encoder-style target for N = max n + 1 images."""
    N = max([c[0] for c in cells], default=0) + 1
    t = torch.zeros(N, S, S, 5 * B + C)
    g = torch.Generator().manual_seed(99)
    for k, (n, i, j) in enumerate(cells):
        bx = boxes[k] if boxes is not None else (torch.rand(4, generator=g) * 0.8 + 0.1)
        t[n, i, j, :B] = 1
        for b in range(B):
            t[n, i, j, B + 4 * b:B + 4 * b + 4] = torch.as_tensor(bx)
        t[n, i, j, 5 * B + (cls[k] if cls is not None else (3 * k) % C)] = 1
    return t


def loss_cases():
    cases = {}

    def add(name, pred, target, S, B=2, C=20, lc=5.0, ln=0.5, bs=None):
        bs = pred.shape[0] if bs is None else bs
        cases[name] = dict(pred=pred, target=target, S=S, B=B, C=C, lc=lc, ln=ln, bs=bs)

    p, t = synth.make_loss_inputs(8, 7, seed=20241018)
    add("rand_s7_n8", p, t, 7)
    p, t = synth.make_loss_inputs(6, 14, seed=20241019, p_obj=6.0 / 196)
    add("rand_s14_n6", p, t, 14)
    p, t = synth.make_loss_inputs(3, 7, seed=20241020, p_obj=0.5, variant="mixed")
    add("dense_mixed_s7_n3_bs5", p, t, 7, lc=4.0, ln=0.3, bs=5)
    p, t = synth.make_loss_inputs(2, 14, seed=20241021, p_obj=0.5, variant="mixed")
    add("dense_mixed_s14_n2", p, t, 14)
    # quirk boundary: exactly K = 0, 1, 2, 3 objects in the call
    g = torch.Generator().manual_seed(7)
    cells = [(0, 1, 2), (0, 5, 5), (1, 0, 0)]
    for K in range(4):
        pred = torch.rand(2, 7, 7, 30, generator=g) * 0.98 + 0.01
        if K == 0:
            tgt = torch.zeros(2, 7, 7, 30)
        else:
            tgt = hand_target(7, 2, 20, cells[:K])
            if tgt.shape[0] < 2:
                tgt = torch.cat([tgt, torch.zeros(1, 7, 7, 30)])
        add("k%d_objects" % K, pred, tgt, 7)
    # both IoUs zero -> responsible slot 0, zero IoU gradient
    pred = torch.rand(1, 7, 7, 30, generator=g) * 0.98 + 0.01
    tgt = hand_target(7, 2, 20, [(0, 3, 3), (0, 4, 1), (0, 6, 6)],
                      boxes=[[0.1, 0.1, 0.02, 0.02]] * 3)
    for (i, j) in [(3, 3), (4, 1), (6, 6)]:
        pred[0, i, j, 2:6] = torch.tensor([0.9, 0.9, 0.02, 0.02])
        pred[0, i, j, 6:10] = torch.tensor([0.8, 0.85, 0.03, 0.02])
    add("zero_iou", pred, tgt, 7)
    # identical predictor boxes -> tie -> slot 0
    pred = torch.rand(1, 7, 7, 30, generator=g) * 0.98 + 0.01
    tgt = hand_target(7, 2, 20, [(0, 2, 2), (0, 2, 3), (0, 5, 0)])
    pred[0, :, :, 6:10] = pred[0, :, :, 2:6]
    add("tie_identical_boxes", pred, tgt, 7)
    return cases


def gen_loss(meta):
    out = {}
    worst = 0.0
    for name, c in loss_cases().items():
        loss, grad = ref_loss(c["pred"], c["target"], c["S"], c["B"], c["C"], c["lc"], c["ln"], c["bs"])
        pn, tn = c["pred"].numpy(), c["target"].numpy()
        terms, og = O.loss(pn, tn, c["B"], c["C"], c["lc"], c["ln"], c["bs"])
        rel_l = abs(float(terms[4]) - loss) / max(abs(loss), 1e-12)
        rel_g = float(np.abs(og - grad).max() / max(np.abs(grad).max(), 1e-12))
        worst = max(worst, rel_l, rel_g)
        print("loss case %-26s ref=%.7f oracle=%.7f rel_loss=%.2e rel_grad=%.2e objs=%d" % (
            name, loss, float(terms[4]), rel_l, rel_g, int((tn[..., 0] == 1).sum())))
        assert rel_l < 2e-6 and rel_g < 2e-6, "C oracle disagrees with the reference on %s" % name
        out[name + "/pred"] = pn
        out[name + "/target"] = tn
        out[name + "/grad"] = grad
        out[name + "/loss"] = np.float32(loss)
        out[name + "/hyper"] = np.array([c["S"], c["B"], c["C"], c["lc"], c["ln"], c["bs"]], np.float64)
    np.savez_compressed(os.path.join(GOLD, "loss_cases.npz"), **out)
    meta["loss_oracle_vs_reference_worst_rel"] = worst


# --------------------------------------------------------------------------------------------
# decode / nms
# --------------------------------------------------------------------------------------------
def ref_decoder(pred_img, S, thresh, nms_th, gt=False):
    b, c, s = U.decoder(pred_img.clone()[None], grid_num=S, B=2, device="cpu", thresh=thresh,
                        nms_th=nms_th, gt=gt)
    return b.numpy().astype(np.float32), c.numpy(), s.numpy().astype(np.float32)


def canon_rows(b, c, s):
    """Sort detections lexicographically by (score desc, x1, y1, x2, y2, cls): a canonical order for
    outputs whose reference order is tie dependent."""
    key = np.lexsort((c, b[:, 3], b[:, 2], b[:, 1], b[:, 0], -s))
    return b[key], c[key], s[key]


def gen_decode(meta):
    out = {}
    sets = {
        "uni_s7": (synth.make_tie_free_decode_inputs(24, 7, seed=2)[0], 7, 0.1, 0.5, False),
        "uni_s7_map": (synth.make_tie_free_decode_inputs(12, 7, seed=3)[0], 7, 0.005, 0.45, False),
        "sig_s7": (synth.make_tie_free_decode_inputs(16, 7, seed=4, dist="sigmoid")[0], 7, 0.1, 0.5, False),
        "uni_s14": (synth.make_tie_free_decode_inputs(3, 14, seed=5)[0], 14, 0.1, 0.5, False),
        "sig_s14_map": (synth.make_tie_free_decode_inputs(3, 14, seed=6, dist="sigmoid")[0], 14, 0.005, 0.45, False),
    }
    # nothing passes the threshold -> sentinel; conf <= 1e-4 everywhere but the image max is a candidate
    empty = torch.rand(2, 7, 7, 30, generator=torch.Generator().manual_seed(11)) * 0.2
    empty[1, :, :, :2] = 5e-5
    empty[1, 3, 3, 0] = 9e-5
    sets["empty_s7"] = (empty, 7, 0.3, 0.5, False)
    # tiny confidences with a high-probability class: only the max-conf slot is a candidate
    tiny = torch.rand(1, 7, 7, 30, generator=torch.Generator().manual_seed(12))
    tiny[0, :, :, :2] = 5e-5
    tiny[0, 2, 5, 1] = 1e-4 * 0.999
    sets["tiny_conf_s7"] = (tiny, 7, 1e-6, 0.5, False)
    # gt=True round trip of an encoder-style target (YOLODataLoader.py:249)
    _, tg = synth.make_loss_inputs(4, 7, seed=13, p_obj=0.15)
    sets["gt_roundtrip_s7"] = (tg, 7, 0.3, 0.5, True)
    # NaN semantics of the ATen ops the reference runs: torch.max(dim) lets a NaN class score win, so
    # the slot's score is NaN and `> thresh` drops it; `contain.max()` is NaN as soon as one confidence is, so the
    # `== max` rule selects nothing.  (NaN only where it cannot reach an emitted box: payload bits are not portable.)
    nanp = synth.make_tie_free_decode_inputs(4, 7, seed=7)[0].clone()
    nanp[0, 2, 3, 15] = float("nan")      # class c = 5 of one cell, c = 19 of another, c = 0 of a third
    nanp[0, 4, 4, 29] = float("nan")
    nanp[0, 0, 0, 10] = float("nan")
    nanp[1, 1, 1, 0] = float("nan")       # one NaN confidence among ordinary ones
    nanp[2, :, :, :2] *= 5e-5             # every confidence <= 1e-4 and one NaN: no candidate at all
    nanp[2, 5, 2, 1] = float("nan")
    nanp[3, :, :, :2] *= 5e-5             # the same without the NaN: exactly the maximum is a candidate
    sets["nan_s7"] = (nanp, 7, 1e-7, 0.5, False)
    for name, (pred, S, th, nth, gt) in sets.items():
        N = pred.shape[0]
        boxes, clss, probs, counts = [], [], [], []
        for n in range(N):
            b, c, s = ref_decoder(pred[n], S, th, nth, gt)
            ob, oc, os_ = O.decoder(pred[n].numpy(), grid_num=S, thresh=th, nms_th=nth, gt=gt)
            c = np.asarray(c, np.float32)
            if gt:
                # every GT score is exactly 1.0: the order among ties is unspecified in the reference
                # (unstable sort) -> compare as a set, store in canonical (lexicographic) order
                b, c, s = canon_rows(b, c, s)
                ob, oc, os_ = canon_rows(ob, np.asarray(oc, np.float32), os_)
            assert b.shape == ob.shape and np.array_equal(b.view(np.uint32), ob.view(np.uint32)), (name, n)
            assert np.array_equal(np.asarray(c, np.float64), np.asarray(oc, np.float64)), (name, n)
            assert np.array_equal(s.view(np.uint32), os_.view(np.uint32)), (name, n)
            boxes.append(b); clss.append(c); probs.append(s); counts.append(len(s))
        print("decode set %-16s N=%d S=%d thresh=%g nms=%g gt=%s kept/img=%.1f  bit-exact vs C oracle" % (
            name, N, S, th, nth, gt, float(np.mean(counts))))
        out[name + "/pred"] = pred.numpy()
        out[name + "/boxes"] = np.concatenate(boxes, 0)
        out[name + "/cls"] = np.concatenate(clss, 0)
        out[name + "/probs"] = np.concatenate(probs, 0)
        out[name + "/counts"] = np.asarray(counts, np.int32)
        out[name + "/params"] = np.array([S, th, nth, float(gt)], np.float64)
    # stand-alone nms() calls
    g = torch.Generator().manual_seed(21)
    for name, n, thr in [("nms_rand40_default", 40, 0.25), ("nms_rand200_05", 200, 0.5),
                         ("nms_rand7_single", 7, 0.0)]:
        xy = torch.rand(n, 2, generator=g) * 0.7
        wh = torch.rand(n, 2, generator=g) * 0.3 + 0.02
        bx = torch.cat([xy, xy + wh], 1)
        sc = torch.rand(n, generator=g)
        keep = U.nms(bx, sc, thr).numpy()
        okeep = O.nms(bx.numpy(), sc.numpy(), thr)
        assert np.array_equal(keep, okeep), name
        out[name + "/boxes"], out[name + "/scores"] = bx.numpy(), sc.numpy()
        out[name + "/keep"], out[name + "/thr"] = keep.astype(np.int64), np.float64(thr)
        print("nms case %-20s n=%d thr=%g kept=%d" % (name, n, thr, len(keep)))
    # NaN scores: torch.sort(descending=True) puts them first (utils/utils.py:161)
    xy = torch.rand(12, 2, generator=g) * 0.7
    wh = torch.rand(12, 2, generator=g) * 0.3 + 0.02
    bx = torch.cat([xy, xy + wh], 1)
    sc = torch.rand(12, generator=g)
    sc[3] = sc[7] = float("nan")
    keep = U.nms(bx, sc, 0.5).numpy()
    assert np.array_equal(keep, O.nms(bx.numpy(), sc.numpy(), 0.5)) and list(keep[:1]) == [3], keep
    out["nms_nan_scores/boxes"], out["nms_nan_scores/scores"] = bx.numpy(), sc.numpy()
    out["nms_nan_scores/keep"], out["nms_nan_scores/thr"] = keep.astype(np.int64), np.float64(0.5)
    print("nms case %-20s n=12 thr=0.5 kept=%d (NaN scores first)" % ("nms_nan_scores", len(keep)))
    # chain A > B > C: A kills B, B would have killed C, C must survive (iterated suppression)
    bx = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.45, 0.0, 1.45, 1.0], [0.9, 0.0, 1.9, 1.0], [3, 3, 4, 4.0]])
    sc = torch.tensor([0.9, 0.8, 0.7, 0.6])
    keep = U.nms(bx, sc, 0.3).numpy()
    assert np.array_equal(keep, O.nms(bx.numpy(), sc.numpy(), 0.3)) and list(keep) == [0, 2, 3], keep
    out["nms_chain/boxes"], out["nms_chain/scores"] = bx.numpy(), sc.numpy()
    out["nms_chain/keep"], out["nms_chain/thr"] = keep.astype(np.int64), np.float64(0.3)
    np.savez_compressed(os.path.join(GOLD, "decode_cases.npz"), **out)


# --------------------------------------------------------------------------------------------
# helpers + voc_eval + run_test_mAP
# --------------------------------------------------------------------------------------------
def gen_misc(meta):
    # utils/utils.py:506-525 fixture
    b1 = np.array([[10, 20, 100, 123], [200, 300, 300, 350]], np.float32)
    b2 = np.array([[50, 60, 150, 120], [0, 10, 123, 150], [170, 190, 310, 400]], np.float32)
    iou = U.compute_iou_matrix(torch.from_numpy(b1), torch.from_numpy(b2)).numpy()
    assert np.array_equal(iou.view(np.uint32), O.iou_matrix(b1, b2).view(np.uint32))
    meta["iou_fixture"] = dict(b1=b1.tolist(), b2=b2.tolist(), iou=[[float(v) for v in r] for r in iou])
    g = torch.Generator().manual_seed(31)
    bx = torch.rand(6, 4, generator=g)
    conv = U.convert_CxCyWH_to_X1Y1X2Y2(bx, 7, 2, "cpu").numpy()
    assert np.array_equal(conv.view(np.uint32), O.cxcywh_to_xyxy(bx.numpy(), 7).view(np.uint32))
    meta["convert_fixture"] = dict(S=7, boxes=[[float(v) for v in r] for r in bx.numpy()],
                                   out=[[float(v) for v in r] for r in conv])
    # utils/utils.py:321-324 test_eval fixture
    preds = {'cat': [['image01', 0.9, 20, 20, 40, 40], ['image01', 0.8, 20, 20, 50, 50],
                     ['image02', 0.8, 30, 30, 50, 50]], 'dog': [['image01', 0.78, 60, 60, 90, 90]]}
    target = {('image01', 'cat'): [[20, 20, 41, 41]], ('image01', 'dog'): [[60, 60, 91, 91]],
              ('image02', 'cat'): [[30, 30, 51, 51]]}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m = U.voc_eval(copy.deepcopy(preds), copy.deepcopy(target), VOC_CLASSES=['cat', 'dog'])
    meta["voc_eval_fixture"] = dict(mAP=m, log=buf.getvalue().strip().splitlines())
    print("voc_eval fixture mAP", m)


def gen_map(meta):
    """run_test_mAP (utils/utils.py:389-418) with an identity network over synthetic predictions and a
    synthetic GT built from perturbed top detections."""
    S = 7
    pred, _ = synth.make_tie_free_decode_inputs(96, S, seed=41)
    pred_in = pred.numpy().copy()   # the reference decoder overwrites x,y of candidate boxes IN PLACE
    rng = np.random.RandomState(5)  # (utils/utils.py:119,123): keep what the network "produced"

    gt = {}
    for n in range(pred.shape[0]):
        b, c, s = ref_decoder(pred[n], S, 0.005, 0.45)
        for k in range(min(3, len(s))):
            box = np.clip(b[k], 0, 1) * 448
            box = [int(v) for v in (box + rng.randint(-12, 13, size=4))]
            if box[2] <= box[0] or box[3] <= box[1]:
                continue
            gt.setdefault(("img%04d" % n, U.VOC_CLASSES[int(c[k])]), []).append(box)
    dataset = [(pred[n].clone(), torch.zeros(1), "/x/img%04d.jpg" % n) for n in range(pred.shape[0])]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
        m = U.run_test_mAP(lambda x: x, copy.deepcopy(gt), dataset, len(dataset), S=S, device="cpu")
    aps = [float(l.split(" ap ")[1].rstrip("-")) for l in buf.getvalue().splitlines() if "---class" in l]
    print("synthetic mAP case: mAP=%r over %d classes" % (m, len(aps)))
    assert np.array_equal(pred.numpy(), pred_in)
    np.savez_compressed(os.path.join(GOLD, "map_case.npz"), pred=pred_in)
    meta["map_case"] = dict(S=S, mAP=m, aps=aps, gt=[[k[0], k[1], v] for k, v in sorted(gt.items())])


def load_reference_encoder():
    """The reference's target encoder is a method of its dataset class (utils/YOLODataLoader.py:200-230); the module
    imports imgaug (absent here) at the top, so stub that import -- the encoder itself only uses torch."""
    import importlib.util
    import types
    for name in ("imgaug", "imgaug.augmenters"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["imgaug"].seed = lambda *a, **k: None
    sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]
    spec = importlib.util.spec_from_file_location("ref_yolo_dataloader", "/root/reference/utils/YOLODataLoader.py")
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(mod)
    return mod.yoloDataset


def gen_encoder(meta):
    DS = load_reference_encoder()
    out = {}
    rng = np.random.RandomState(17)
    for name, S, n_img in [("enc_s7", 7, 64), ("enc_s14", 14, 32), ("enc_s3_b1", 3, 8)]:
        B, C = (1, 5) if name.endswith("b1") else (2, 20)
        ds = object.__new__(DS)
        ds.S, ds.B, ds.C = S, B, C
        boxes, labels, offsets, targets = [], [], [0], []
        for n in range(n_img):
            k = 0 if n % 9 == 0 else rng.randint(1, 7)          # empty images too (:209-210)
            bx = rng.rand(k, 4).astype(np.float32)
            if k and n % 5 == 0:
                bx[0, :2] = [1.0, 1.0]                           # right / bottom edge
            if k and n % 7 == 0:
                bx[-1, :2] = [0.0, 0.25]                         # ij = -1 wraps to the last column (Python indexing)
            if k > 1 and n % 3 == 0:
                bx[1, :2] = bx[0, :2]                            # two objects in one cell: the last one wins
            if k and n % 4 == 0:
                bx[0, 0] = np.float32(3.0 / S)                   # exactly on a cell boundary
            lb = rng.randint(0, C, size=k)
            t = ds.encoder(torch.from_numpy(bx).reshape(-1, 4), torch.from_numpy(lb))
            boxes.append(bx.reshape(-1, 4)); labels.append(lb); offsets.append(offsets[-1] + k); targets.append(t.numpy())
        boxes = np.concatenate(boxes, 0).astype(np.float32)
        labels = np.concatenate(labels, 0).astype(np.int32)
        target = np.stack(targets).astype(np.float32)
        mine = O.encode(boxes, labels, offsets, S, B, C)
        assert np.array_equal(mine.view(np.uint32), target.view(np.uint32)), name
        out[name + "/boxes"], out[name + "/labels"] = boxes, labels
        out[name + "/offsets"], out[name + "/target"] = np.asarray(offsets, np.int64), target
        out[name + "/params"] = np.array([S, B, C], np.int64)
        print("encoder set %-10s images=%d objects=%d  bit-exact vs C oracle" % (name, n_img, len(labels)))
    np.savez_compressed(os.path.join(GOLD, "encoder_cases.npz"), **out)


def time_reference(meta):
    """Timing of the reference's own Python path in THIS container (context for DESIGN.md/BASELINE.md)."""
    torch.set_num_threads(os.cpu_count())
    p, t = synth.make_loss_inputs(32, 7, seed=20241018 + 1000)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        ref_loss(p, t, 7, 2, 20, 5.0, 0.5, 32)
        ts.append(time.perf_counter() - t0)
    loss_ms = float(np.median(ts) * 1e3)
    pd = synth.make_decode_inputs(64, 7, seed=2)
    t0 = time.perf_counter()
    for n in range(64):
        ref_decoder(pd[n], 7, 0.1, 0.5)
    dec = 64 / (time.perf_counter() - t0)
    meta["reference_python_timing_build_container"] = dict(
        cores=os.cpu_count(), loss_fwd_bwd_ms_n32_s7=loss_ms, loss_cells_per_s=32 * 49 / (loss_ms * 1e-3),
        objects=int((t[..., 0] == 1).sum()), decode_nms_images_per_s=dec, torch=torch.__version__)
    print("reference python timing:", meta["reference_python_timing_build_container"])


def main():
    os.makedirs(GOLD, exist_ok=True)
    O.build()
    meta = dict(generator="oracle/make_golden.py", torch=torch.__version__, numpy=np.__version__)
    gen_loss(meta)
    gen_decode(meta)
    gen_misc(meta)
    gen_map(meta)
    gen_encoder(meta)
    time_reference(meta)
    with open(os.path.join(GOLD, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
