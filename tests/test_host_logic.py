"""CPU-only: host-side logic that needs no GPU -- sharding arithmetic, interface shape of the
reference-compatible module and functions, input validation, synthetic generators."""
import inspect

import numpy as np
import pytest
import torch

import yolo_v1_b200 as y
from yolo_v1_b200 import synth


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 1048576, 1048577):
        for w in (1, 2, 3, 4, 8):
            spans = [y.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        y.shard_range(10, 4, 4)


def test_signatures_match_the_reference_call_sites():
    """v1Loss.py:10,22 -- utils/utils.py:94,150 (names, order and defaults)."""
    sig = inspect.signature(y.YOLOLossV1.__init__)
    names = list(sig.parameters)[1:10]
    assert names == ['_batch_size', '_S', '_B', '_clsN', '_l_coord', '_l_noobj', '_device', '_logger', '_vis']
    assert sig.parameters['_l_coord'].default == 5. and sig.parameters['_l_noobj'].default == 0.5
    assert sig.parameters['_device'].default == 'cuda:0'
    assert list(inspect.signature(y.YOLOLossV1.forward).parameters) == ['self', 'pred_tensor', 'target_tensor']
    d = inspect.signature(y.decoder)
    assert list(d.parameters) == ['pred', 'grid_num', 'B', 'device', 'thresh', 'nms_th', 'gt']
    assert [d.parameters[k].default for k in list(d.parameters)[1:]] == [7, 2, 'cpu', 0.3, 0.5, False]
    n = inspect.signature(y.nms)
    assert list(n.parameters) == ['bboxes', 'scores', 'threshold'] and n.parameters['threshold'].default == 0.25
    m = y.YOLOLossV1(32, 7, 2, 20)
    assert (m.S, m.B, m.C, m.batch_size, m.lambda_coord, m.lambda_noobj) == (7, 2, 20, 32, 5., .5)
    assert len(m.state_dict()) == 0
    with pytest.raises(ValueError):
        y.YOLOLossV1(32, 7, 2, 20, coord_mode="nope")


def test_compat_modules_import_like_the_reference():
    """`from v1Loss import YOLOLossV1` / `from utils.utils import decoder, nms` with yolo_v1_b200/compat on
    sys.path (how INTEGRATION.md wires train.py / eval.py)."""
    import importlib
    import os
    import sys
    compat = os.path.join(os.path.dirname(y.__file__), "compat")
    sys.path.insert(0, compat)
    saved = {k: sys.modules.pop(k, None) for k in ("utils", "utils.utils", "v1Loss")}
    try:
        v1 = importlib.import_module("v1Loss")
        uu = importlib.import_module("utils.utils")
        assert v1.YOLOLossV1 is y.YOLOLossV1
        assert uu.decoder is y.decoder and uu.nms is y.nms
        assert uu.compute_iou_matrix is y.compute_iou_matrix
        assert len(uu.VOC_CLASSES) == 20
    finally:
        sys.path.remove(compat)
        for k in ("utils", "utils.utils", "v1Loss"):
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]


_IMPORT_LINES = """
import sys, types
sys.dont_write_bytecode = True
sys.path[:0] = [{compat!r}, {root!r}, {ref!r}]
for m in ("imgaug", "imgaug.augmenters", "visdom", "torchsummary"):     # third-party packages absent from this image
    sys.modules.setdefault(m, types.ModuleType(m))
sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]
sys.modules["imgaug"].seed = lambda *a, **k: None
sys.modules["visdom"].Visdom = object
# --- train.py:12-14,20 / eval.py:15,17 / run_voc_mAP.py:3, verbatim ---
from v1Loss import YOLOLossV1
from utils.YOLODataLoader import yoloDataset
from utils.utils import *
from utils.visual import Visual
# ---
import yolo_v1_b200 as y
import utils.utils as uu, utils.YOLODataLoader as dl
assert YOLOLossV1 is y.YOLOLossV1 and decoder is y.decoder and nms is y.nms and run_test_mAP is y.run_test_mAP
for name in {helpers!r}:                      # helpers train.py / eval.py use that are NOT on the hot path
    assert callable(globals()[name]), name
assert uu._ref is not None and uu._ref.decoder is y.decoder and uu._ref.nms is y.nms   # reference-internal callers
assert getattr(dl, "decoder", y.decoder) is y.decoder      # a module that star-imports utils.utils gets ours
assert yoloDataset.__module__ == "utils.YOLODataLoader" and Visual.__module__ == "utils.visual"
print("ok")
"""


def _run_import_lines(ref_root, helpers):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(y.__file__)))
    code = _IMPORT_LINES.format(compat=os.path.join(root, "yolo_v1_b200", "compat"), root=root, ref=ref_root,
                                helpers=tuple(helpers))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_compat_extends_the_reference_utils_package(tmp_path):
    """with `compat` ahead of the reference on sys.path the shim must EXTEND the reference's `utils`
    package, not shadow it -- train.py:13 `from utils.YOLODataLoader import yoloDataset`, train.py:20
    `from utils.visual import Visual` and the non-hot-path helpers of `from utils.utils import *` keep working.
    Runs train.py's import lines against a stand-in reference tree (same file layout, no __init__.py)."""
    ref = tmp_path / "ref"
    (ref / "utils").mkdir(parents=True)
    (ref / "utils" / "utils.py").write_text(
        "import os, numpy as np\n"
        "def decoder(*a, **k):\n    raise AssertionError('the reference decoder must be overridden')\n"
        "def nms(*a, **k):\n    raise AssertionError('the reference nms must be overridden')\n"
        "def prep_test_data(file_path, little_test=None):\n    return file_path\n"
        "def create_logger(base_path, log_name):\n    return None\n"
        "def cv_resize(img, resize=448):\n    return img\n"
        "def bbox_un_norm(bboxes, img_size=(448, 448)):\n    return bboxes\n"
        "def run_test_mAP(*a, **k):\n    return decoder()\n")
    (ref / "utils" / "YOLODataLoader.py").write_text("from utils.utils import *\nclass yoloDataset:\n    pass\n")
    (ref / "utils" / "visual.py").write_text("class Visual:\n    pass\n")
    _run_import_lines(str(ref), ("prep_test_data", "create_logger", "cv_resize", "bbox_un_norm"))


def test_compat_against_the_real_reference_tree():
    """The same import lines against /root/reference itself (build container only)."""
    import os
    ref = os.environ.get("YOLO1_REFERENCE_ROOT", "/root/reference")
    if not os.path.isfile(os.path.join(ref, "utils", "utils.py")):
        pytest.skip("reference tree not mounted")
    _run_import_lines(ref, ("prep_test_data", "create_logger", "cv_resize", "bbox_un_norm", "make_eval_tensor",
                            "from_img_path_get_label_list", "draw_debug_rect"))


def test_helper_functions_cpu():
    b1 = torch.tensor([[10., 20., 100., 123.], [200., 300., 300., 350.]])
    b2 = torch.tensor([[50., 60., 150., 120.], [0., 10., 123., 150.], [170., 190., 310., 400.]])
    iou = y.compute_iou_matrix(b1, b2)          # utils/utils.py:506-525 print-only fixture
    assert np.allclose(iou.numpy(), [[0.24449877, 0.53832752, 0.0], [0.0, 0.0, 0.17006803]], atol=1e-7)
    with pytest.raises(TypeError):
        y.compute_iou_matrix([[0, 0, 1, 1]], b2)
    out = y.convert_CxCyWH_to_X1Y1X2Y2(torch.tensor([[0.5, 0.5, 0.2, 0.4]]), 7)
    assert np.allclose(out.numpy(), [[0.5 / 7 - 0.1, 0.5 / 7 - 0.2, 0.5 / 7 + 0.1, 0.5 / 7 + 0.2]])
    with pytest.raises(AssertionError):
        y.convert_CxCyWH_to_X1Y1X2Y2(torch.zeros(2, 3), 7)


def test_synth_targets_follow_the_encoder_format():
    pred, target = synth.make_loss_inputs(64, 7, seed=1)
    obj = target[..., 0] == 1
    assert 0 < int(obj.sum()) < 64 * 49
    assert torch.equal(target[..., 0], target[..., 1])                    # both confidence slots
    assert torch.equal(target[..., 2:6], target[..., 6:10])               # same box in both slots
    assert torch.equal(target[..., 10:].sum(-1), obj.float())             # one-hot class on object cells only
    assert float(target[~obj].abs().sum()) == 0
    assert float(pred.min()) >= 0.01 and float(pred.max()) <= 0.99
    p2, n = synth.make_tie_free_decode_inputs(128, 7, seed=2)
    assert synth.score_tie_images(p2).numel() == 0 and n >= 0


# ---- properties the decode+NMS kernel's design rests on (DESIGN.md, K2+K3), checked here without a GPU -------------
def test_fixed_point_passes_reach_the_greedy_keep_set():
    """Sweep as a fixed point: K(j) = no i < j with K(i) and M[i][j], re-evaluated for all j until nothing changes,
    gives exactly what the reference's greedy loop keeps (utils/utils.py:162-182), whatever the kill matrix."""
    rng = np.random.RandomState(3)
    for trial in range(300):
        n = int(rng.randint(1, 140))
        dens = rng.choice([0.0, 0.02, 0.1, 0.5, 1.0])
        M = np.triu(rng.rand(n, n) < dens, 1)                     # M[i][j], i < j: j dies if i is kept
        if trial % 10 == 0:                                       # a full chain: i kills i+1
            M = np.zeros((n, n), bool)
            M[np.arange(n - 1), np.arange(1, n)] = True
        greedy = np.ones(n, bool)
        for i in range(n):
            if greedy[i]:
                greedy[M[i]] = False
        K = np.ones(n, bool)
        passes = 0
        while True:
            newK = ~(M & K[:, None]).any(0)
            passes += 1
            if np.array_equal(newK, K):
                break
            K = newK
        assert np.array_equal(K, greedy), (trial, n, dens)
        assert passes <= n + 1


def test_pair_pretest_never_settles_a_pair_that_dies():
    """The pair loop's fp32 pre-test (decode_nms.cu, pair_margin): m = fma(inter, k, -(ta_a + ta_b)) < 0 must imply
    fl32(inter / u) <= thr, the reference's survival test (utils/utils.py:179-180), for the constants
    set_threshold() derives.  numpy float32 emulation (the fma in float64: inter * k is exact there)."""
    f = np.float32
    rng = np.random.RandomState(5)

    def check(thr, A, B):
        thr = f(thr)
        thr_lo = np.nextafter(f(np.float64(thr) * (1 - 2.0 ** -18)), f(0))
        k = np.nextafter(f(1.0 + np.float64(thr_lo)), f(np.inf))
        aa = (A[:, 2] - A[:, 0]) * (A[:, 3] - A[:, 1])
        ab = (B[:, 2] - B[:, 0]) * (B[:, 3] - B[:, 1])
        tame = (aa >= f(1e-30)) & (ab >= f(1e-30)) & (aa <= f(1e30)) & (ab <= f(1e30))
        ww = np.maximum(np.minimum(B[:, 2], A[:, 2]) - np.maximum(B[:, 0], A[:, 0]), f(0))
        hh = np.maximum(np.minimum(B[:, 3], A[:, 3]) - np.maximum(B[:, 1], A[:, 1]), f(0))
        inter = ww * hh
        R = (-(thr_lo * aa)) - (thr_lo * ab)
        settled = (inter.astype(np.float64) * np.float64(k) + R.astype(np.float64)) < 0
        with np.errstate(all="ignore"):
            survives = inter / ((aa + ab) - inter) <= thr
        assert not (tame & settled & ~survives).any(), thr
        return (tame & settled).sum(), (tame & survives).sum()

    n = 200_000
    for thr in (0.5, 0.45, 0.1, 1.0, 0.99999, 2.0 ** -20, 4.0, 0.3, 0.005):
        c, wh = rng.rand(n, 2).astype(f), rng.rand(n, 2).astype(f)
        A = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(f)
        B = (A + rng.randn(n, 4) * rng.choice([1e-3, 1e-2, 0.1, 0.5], (n, 1))).astype(f)
        settled, survive = check(thr, A, B)
        assert settled >= 0.999 * survive                       # ... and it settles practically every survivor
        check(thr, A * f(1e-14), B * f(1e-14))                  # tiny and huge boxes
        check(thr, A * f(1e14), B * f(1e14))
        s = np.exp(rng.uniform(-30, 30, n)).astype(f)            # IoU within +-40 ulp of the threshold
        t = np.clip(f(thr) * (1 + rng.randint(-40, 40, n) * 2.0 ** -23), 0, 1).astype(f)
        z = np.zeros(n, f)
        check(thr, np.stack([z, z, s, s], 1), np.stack([z, z, s, (s * t).astype(f)], 1))


def _area_skip(thr, nb=256):
    """set_threshold() of decode_nms.cu, restated: the smallest bucket distance d whose guaranteed area ratio
    G(d) = min_f lower(f + d - 1) / lower(f) reaches (1 + 1e-5) / thr (area keys = float bits >> 20: 8 per octave)."""
    need = (1.0 + 1e-5) / float(np.float32(thr))
    for d in range(1, nb + 1):
        g = min(np.ldexp(1.0 + ((f + d - 1) % 8) / 8.0, (f + d - 1) // 8) / (1.0 + f / 8.0) for f in range(8))
        if g >= need:
            return d
    return nb + 1


def test_area_pruning_never_skips_a_pair_that_dies():
    """decode_nms.cu, nms_phase (2): pairs whose area buckets are at least area_skip apart are never tested, on the
    grounds that IoU <= min(area) / max(area).  numpy float32 emulation of the reference's test (utils/utils.py:166-180)
    over random and adversarial (nested, nearly equal-ratio) boxes: no pair that dies may be skipped -- with clamped
    bucket indices too -- and at thr = 0.5 about two thirds of the pairs of a uniform image are skipped."""
    f = np.float32
    rng = np.random.RandomState(11)
    key0 = (127 - 24) << 3

    def check(thr, A, B):
        thr = f(thr)
        skip = _area_skip(thr)
        aa = (A[:, 2] - A[:, 0]) * (A[:, 3] - A[:, 1])
        ab = (B[:, 2] - B[:, 0]) * (B[:, 3] - B[:, 1])
        tame = (aa >= f(1e-30)) & (ab >= f(1e-30)) & (aa <= f(1e30)) & (ab <= f(1e30))
        ka = np.clip((aa.view(np.uint32) >> 20).astype(np.int64) - key0, 0, 255)
        kb = np.clip((ab.view(np.uint32) >> 20).astype(np.int64) - key0, 0, 255)
        skipped = np.abs(ka - kb) >= skip
        ww = np.maximum(np.minimum(B[:, 2], A[:, 2]) - np.maximum(B[:, 0], A[:, 0]), f(0))
        hh = np.maximum(np.minimum(B[:, 3], A[:, 3]) - np.maximum(B[:, 1], A[:, 1]), f(0))
        inter = ww * hh
        assert (inter[tame] <= np.minimum(aa, ab)[tame]).all()      # the monotonicity the bound rests on, in fp32
        with np.errstate(all="ignore"):
            dies = ~(inter / ((aa + ab) - inter) <= thr)
        assert not (tame & skipped & dies).any(), thr
        return float((tame & skipped).sum()) / max(int(tame.sum()), 1)

    n = 200_000
    for thr in (0.5, 0.45, 0.25, 0.1, 1.0, 0.99999, 2.0 ** -20, 4.0, 0.005):
        c, wh = rng.rand(n, 2).astype(f), rng.rand(n, 2).astype(f)
        A = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(f)
        c2, wh2 = rng.rand(n, 2).astype(f), rng.rand(n, 2).astype(f)
        B = np.concatenate([c2 - wh2 / 2, c2 + wh2 / 2], 1).astype(f)
        frac = check(thr, A, B)
        if thr == 0.5:
            assert frac > 0.6, frac
        # nested boxes: IoU == area ratio exactly (the bound is tight), ratios within a few ulp of 1 / thr and of the
        # bucket boundaries, at tiny / ordinary / huge scales (the last two exercise the clamped buckets)
        for scale in (1e-6, 1.0, 1e3):
            s = (np.exp(rng.uniform(-3, 3, n)) * scale).astype(f)
            r = (min(float(f(thr)), 1.0) * (1 + rng.randint(-64, 64, n) * 2.0 ** -23)).astype(f)
            z = np.zeros(n, f)
            big = np.stack([z, z, s, s], 1)
            small = np.stack([z, z, s, (s * r).astype(f)], 1)
            check(thr, big, small)
            check(thr, small, big)
    assert _area_skip(0.5) == 10 and _area_skip(1.0) == 2      # a little over one octave at 0.5; neighbours only at 1


def test_autotune_zero_copy_picks_the_fastest_mode_for_all_ranks():
    """yolo_v1_b200.host.autotune_zero_copy: every mode is timed after a barrier, the per-rank time goes through
    reduce_max (so all ranks see the same table and pick the same mode), the context is left in the winning mode."""
    import time
    from yolo_v1_b200 import host as yhost

    class FakeCtx:
        def __init__(self, cost):
            self.cost, self.mode, self.calls = cost, None, []

        def set_zero_copy(self, m):
            self.mode = m

        def loss(self, pred, target, batch_size, out_grad=None):
            self.calls.append(self.mode)
            time.sleep(self.cost[self.mode])

    events = []
    ctx = FakeCtx({0: 0.004, 1: 0.002, 2: 0.001, 4: 0.003})
    # this rank finds mode 2 fastest, but another rank (simulated by reduce_max) is slow in mode 2: the job-wide pick
    # must follow the maximum over the ranks
    slow_elsewhere = {2: 50.0}
    best, table = yhost.autotune_zero_copy(
        ctx, None, None, None, 8, modes=(2, 1, 4, 0), repeats=2,
        barrier=lambda: events.append(("barrier", ctx.mode)),
        reduce_max=lambda ms: max(ms, slow_elsewhere.get(ctx.mode, 0.0)))
    assert best == 1 and ctx.mode == 1 and set(table) == {0, 1, 2, 4} and table[2] == 50.0
    assert [m for _, m in events] == [2, 1, 4, 0]                      # one barrier per mode, before its timed loop
    assert ctx.calls.count(2) == 3 and ctx.calls.count(0) == 3         # one warm-up + `repeats` timed calls per mode
    with yhost.near_gpu(0) as n:                                        # no GPU / no sysfs answer: a no-op
        assert n.cpus is None or len(n.cpus) > 0


def test_staged_reference_archive_is_byte_identical_and_loads(tmp_path):
    """oracle/stage_reference.py packs the reference's two hot-path modules into oracle/_ref/reference_hot_path.zip for
    the GPU box; oracle/ref_loader.py must be able to run the reference from that archive alone, and the members must
    be the mounted files byte for byte.  (Build container only: needs the reference mount.)"""
    import hashlib
    import os
    import subprocess
    import sys
    import zipfile
    ref = "/root/reference"
    if not os.path.isfile(os.path.join(ref, "v1Loss.py")):
        pytest.skip("reference tree not mounted")
    root = os.path.dirname(os.path.dirname(os.path.abspath(y.__file__)))
    from oracle import stage_reference
    assert stage_reference.stage() and os.path.isfile(stage_reference.ARCHIVE)
    with zipfile.ZipFile(stage_reference.ARCHIVE) as z:
        assert sorted(z.namelist()) == ["utils/utils.py", "v1Loss.py"]
        for name in z.namelist():
            assert hashlib.sha256(z.read(name)).hexdigest() == hashlib.sha256(open(os.path.join(ref, name), "rb").read()).hexdigest()
    code = """
import sys, numpy as np, torch
sys.path.insert(0, %r)
from oracle import ref_loader, oracle as O
from yolo_v1_b200 import synth
assert ref_loader.REF_ROOT.endswith("reference_hot_path.zip"), ref_loader.REF_ROOT
Loss, U = ref_loader.load_reference()
pred, target = synth.make_loss_inputs(4, 7, seed=3, p_obj=0.2)
p = pred.clone().requires_grad_(True)
with ref_loader.quiet():
    l = Loss(4, 7, 2, 20, 5., .5, _device='cpu')(p, target); l.backward()
    k = U.nms(torch.tensor([[0., 0., 1., 1.], [0.1, 0., 1.1, 1.], [3., 3., 4., 4.]]), torch.tensor([.9, .8, .7]), 0.5)
t, g = O.loss(pred.numpy(), target.numpy(), batch_size=4)
assert abs(float(l) - float(t[4])) <= 2e-6 * abs(float(l)) and np.abs(g - p.grad.numpy()).max() <= 2e-6 * np.abs(g).max()
assert k.tolist() == [0, 2]
print("ok")
""" % root
    env = dict(os.environ, YOLO1_REFERENCE_ROOT=stage_reference.ARCHIVE, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
