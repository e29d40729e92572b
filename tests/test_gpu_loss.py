"""K1 parity: the fused CUDA loss (through the C ABI / the reference-shaped module) against
  (1) the golden vectors produced by the reference's own code (tests/golden/loss_cases.npz),
  (2) the CPU oracle (oracle/yolo1_oracle.c, itself pinned to those vectors) on seeded synthetic inputs,
  (3) size-independent properties at BASELINE config sizes.
Tolerance (north_star): loss and gradients within 1e-5 relative in fp32
  -- loss: |a-b| <= 1e-5 |b|;  gradient: max|a-b| <= 1e-5 max|b|.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from yolo_v1_b200 import synth

pytestmark = pytest.mark.gpu

TOL = 1e-5
VARIANTS = [0, 1, 2, 3, 5, 8, 13, -1, 31, 40, 41, 42]   # launch shapes of the streaming kernel; -1 = strided kernel;
# 31 = streaming default shape even for a small call; 40-42 = confidence-first sector-read kernel (loss_sparse.cu)


def _y():
    import yolo_v1_b200 as y
    return y


def _check(terms, grad, o_terms, o_grad, what, tol=TOL):
    terms = terms.detach().float().cpu().numpy()
    for t in range(5):
        assert abs(terms[t] - o_terms[t]) <= tol * max(abs(o_terms[t]), 1e-6), (what, "term", t, terms, o_terms)
    if grad is not None:
        g = grad.detach().float().cpu().numpy()
        assert g.shape == o_grad.shape
        if g.size == 0:
            return
        err = np.abs(g - o_grad).max()
        assert err <= tol * max(np.abs(o_grad).max(), 1e-12), (what, "grad", err, np.abs(o_grad).max())


def test_golden_cases_from_the_reference(golden_dir):
    y = _y()
    z = np.load(os.path.join(golden_dir, "loss_cases.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) >= 10
    for name in names:
        S, B, C, lc, ln, bs = z[name + "/hyper"]
        pred = torch.from_numpy(z[name + "/pred"]).cuda()
        target = torch.from_numpy(z[name + "/target"]).cuda()
        for variant in (0, 3, -1):
            loss, grad, terms = y.yolo_loss_fused(pred, target, batch_size=float(bs), S=int(S), B=int(B), C=int(C),
                                                  l_coord=float(lc), l_noobj=float(ln), variant=variant)
            ref_loss, ref_grad = float(z[name + "/loss"]), z[name + "/grad"]
            assert abs(float(loss) - ref_loss) <= TOL * max(abs(ref_loss), 1e-12), (name, variant)
            err = np.abs(grad.cpu().numpy() - ref_grad).max()
            assert err <= TOL * max(np.abs(ref_grad).max(), 1e-12), (name, variant, err)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("S,N,p_obj,kind", [(7, 67, None, "encoder"), (14, 33, None, "encoder"),
                                            (7, 40, 0.5, "mixed"), (14, 9, 0.5, "mixed")])
def test_against_oracle_all_launch_shapes(variant, S, N, p_obj, kind):
    y = _y()
    pred, target = synth.make_loss_inputs(N, S, p_obj=p_obj, seed=20241018 + S + N, variant=kind)
    o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=N)
    loss, grad, terms = y.yolo_loss_fused(pred.cuda(), target.cuda(), batch_size=N, variant=variant)
    _check(terms, grad, o_terms, o_grad, (variant, S, N, kind))
    assert float(loss) == float(terms[4])
    # forward only (grad = NULL) gives the same terms
    _, g2, t2 = y.yolo_loss_fused(pred.cuda(), target.cuda(), batch_size=N, variant=variant, want_grad=False)
    assert g2 is None and torch.equal(t2, terms)


@pytest.mark.parametrize("N", [0, 1, 2, 3, 5, 127, 128, 129, 1000])
def test_ragged_sizes_and_tail_path(N):
    """N*S*S not a multiple of the tile: the tail cells go through the direct path; N=0 gives zeros."""
    y = _y()
    pred, target = synth.make_loss_inputs(N, 7, seed=5 + N, p_obj=0.2)
    bs = max(N, 1)
    o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=bs)
    for variant in (0, 5, 3):
        _, grad, terms = y.yolo_loss_fused(pred.cuda(), target.cuda(), batch_size=bs, variant=variant)
        _check(terms, grad, o_terms, o_grad, ("ragged", N, variant))


def test_permuted_nchw_view_is_read_in_place():
    """The backbone hands the loss a permuted NCHW view (OriginResNet.py:189): same numbers, gradient
    written in the same physical layout, no .contiguous()."""
    y = _y()
    pred, target = synth.make_loss_inputs(50, 14, seed=77, p_obj=0.1)
    o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=50)
    planar = pred.cuda().permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    assert not planar.is_contiguous()
    _, grad, terms = y.yolo_loss_fused(planar, target.cuda(), batch_size=50)
    assert grad.stride() == planar.stride()
    _check(terms, grad, o_terms, o_grad, "planar")
    # an offset (unaligned) slice of a bigger buffer also works (falls to the strided kernel)
    big = torch.zeros(51 * 14 * 14 * 30 + 1, device="cuda")
    view = big[1:1 + 50 * 14 * 14 * 30].view(50, 14, 14, 30)
    view.copy_(pred)
    _, grad, terms = y.yolo_loss_fused(view, target.cuda(), batch_size=50)
    _check(terms, grad, o_terms, o_grad, "unaligned")


@pytest.mark.parametrize("S,N", [(7, 1), (7, 3), (7, 4), (7, 67), (14, 1), (14, 33), (3, 50), (16, 5), (17, 3)])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_planar_fast_path_all_shapes(S, N, dtype):
    """Channel-planar pred/grad (the backbone's view): whole-image tiles through the copy engine, image tails
    through the strided path, grids that do not fit a 256-cell tile (S=17) through the strided kernel."""
    y = _y()
    pred, target = synth.make_loss_inputs(N, S, seed=900 + S + N, p_obj=0.2, variant="mixed")
    if dtype == "bf16":
        pred = pred.to(torch.bfloat16)
    o_terms, o_grad = O.loss(pred.float().numpy(), target.numpy(), batch_size=N)
    planar = pred.cuda().permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    for variant in (0, 1, 20, 31, 50, 51):   # 31: streaming default; 50: confidence-first planar kernel; 51: dense planar
        _, grad, terms = y.yolo_loss_fused(planar, target.cuda(), batch_size=N, variant=variant)
        assert grad.stride() == planar.stride()
        if dtype == "f32":
            _check(terms, grad, o_terms, o_grad, ("planar", S, N, variant))
        else:
            _check(terms, None, o_terms, None, ("planar bf16", S, N, variant))
            g = grad.float().cpu().numpy()
            assert np.all(np.abs(g - o_grad) <= np.abs(o_grad) * 2.0 ** -8 + 1e-30)
        _, _, t2 = y.yolo_loss_fused(planar, target.cuda(), batch_size=N, variant=variant, want_grad=False)
        assert torch.equal(t2, terms)


@pytest.mark.parametrize("S,N", [(14, 4000), (7, 9000)])
def test_planar_confidence_first_kernel_many_tiles_per_cta(S, N):
    """csrc/loss_planar_sparse.cu keeps its gradient tiles in shared memory between tiles and only clears what an
    object cell wrote two tiles earlier: run it with several tiles per CTA and an object in every fourth cell, in every
    form (dense target, object lists, fused sigmoid head, bf16, forward only), against the oracle and against the dense
    planar kernels (variant 51)."""
    y = _y()
    pred, target = synth.make_loss_inputs(N, S, seed=77 + S, p_obj=0.25)
    o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=N)
    tc = target.cuda()
    planar = pred.cuda().permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    _, grad, terms = y.yolo_loss_fused(planar, tc, batch_size=N, variant=50)
    assert grad.stride() == planar.stride()
    _check(terms, grad, o_terms, o_grad, ("planar sparse", S, N))
    _, g51, t51 = y.yolo_loss_fused(planar, tc, batch_size=N, variant=51)
    assert torch.equal(g51, grad) and torch.allclose(t51, terms, rtol=1e-6)
    _, _, t0 = y.yolo_loss_fused(planar, tc, batch_size=N, variant=50, want_grad=False)
    assert torch.allclose(t0, terms, rtol=1e-6)
    # object lists (4-byte ownership map instead of the target rows)
    objmask = target[..., 0] == 1
    idx = objmask.nonzero()
    bx = target[idx[:, 0], idx[:, 1], idx[:, 2], 2:6]
    cxcy = (bx[:, :2] + torch.stack([idx[:, 2], idx[:, 1]], 1).float()) / S
    boxes = torch.cat([cxcy, bx[:, 2:]], 1).contiguous()
    labels = target[idx[:, 0], idx[:, 1], idx[:, 2], 10:].argmax(1).to(torch.int32)
    offs = torch.zeros(N + 1, dtype=torch.int64)
    offs[1:] = objmask.reshape(N, -1).sum(1).cumsum(0)
    enc = y.encode_targets(boxes.cuda(), labels.cuda(), offs.cuda(), S)
    oe_terms, oe_grad = O.loss(pred.numpy(), enc.cpu().numpy(), batch_size=N)
    _, gl, tl = y.yolo_loss_from_objects(planar, boxes.cuda(), labels.cuda(), offs.cuda(), batch_size=N, variant=50)
    _check(tl, gl, oe_terms, oe_grad, ("planar sparse lists", S, N))
    # fused sigmoid head, fp32 and bf16 logits
    z = torch.logit(pred.clamp(1e-4, 1 - 1e-4))
    zp = z.cuda().permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    # (the logits entry point has no variant argument: fp32 whole-image tiles of a 14x14 grid take the
    #  confidence-first kernel by default, 7x7 grids the dense planar kernel -- both against the oracle)
    p64 = torch.sigmoid(z.double())
    os_terms, os_grad = O.loss(p64.float().numpy(), target.numpy(), batch_size=N)
    want = os_grad.astype(np.float64) * (p64 * (1 - p64)).numpy()
    _, gs, ts_ = y.yolo_loss_fused(zp, tc, batch_size=N, from_logits=True)
    assert np.all(np.abs(ts_.cpu().numpy() - os_terms) <= TOL * np.abs(os_terms) + 1e-7)
    assert np.abs(gs.cpu().numpy() - want).max() <= TOL * np.abs(want).max()
    pb = planar.to(torch.bfloat16)
    _, gb, tb = y.yolo_loss_fused(pb, tc, batch_size=N, variant=50)
    _, gb2, tb2 = y.yolo_loss_fused(pb, tc, batch_size=N, variant=51)
    assert torch.equal(gb, gb2) and torch.allclose(tb, tb2, rtol=1e-6)


def test_batch_size_divisor_lambdas_and_paper_mode():
    y = _y()
    pred, target = synth.make_loss_inputs(16, 7, seed=3, p_obj=0.3, variant="mixed")
    for bs, lc, ln, mode in [(64, 5.0, 0.5, 0), (16, 2.5, 0.25, 0), (7, 5.0, 0.5, 1)]:
        o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), l_coord=lc, l_noobj=ln, batch_size=bs, coord_mode=mode)
        _, grad, terms = y.yolo_loss_fused(pred.cuda(), target.cuda(), batch_size=bs, l_coord=lc, l_noobj=ln,
                                           coord_mode="paper" if mode else "reference")
        _check(terms, grad, o_terms, o_grad, (bs, lc, ln, mode))


def test_general_B_and_C():
    """Any (B <= 8, C): contiguous tensors go through the runtime-channel-count bulk-copy kernel (tile sized to the
    channel count), other layouts through the strided kernel; both against the oracle."""
    y = _y()
    for B, C, N in [(1, 20, 12), (3, 5, 12), (2, 21, 12), (4, 1, 12), (2, 80, 300), (8, 88, 40), (1, 1, 1000)]:
        pred, target = synth.make_loss_inputs(N, 7, B=B, C=C, seed=B * 100 + C, p_obj=0.3, variant="mixed")
        o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), B=B, C=C, batch_size=N)
        for layout in ("nhwc", "strided"):
            pc = pred.cuda()
            if layout == "strided":
                pc = torch.zeros(N, 7, 7, 5 * B + C + 2, device="cuda")[..., 1:1 + 5 * B + C].copy_(pred)
            _, grad, terms = y.yolo_loss_fused(pc, target.cuda(), batch_size=N, B=B, C=C)
            _check(terms, grad, o_terms, o_grad, (B, C, layout))
            _, g0, t0 = y.yolo_loss_fused(pc, target.cuda(), batch_size=N, B=B, C=C, want_grad=False)
            assert g0 is None and torch.allclose(t0, terms, rtol=2e-6, atol=1e-7)
    # fused sigmoid head and bf16 with a non-default channel count
    B, C, N = 2, 80, 64
    _, target = synth.make_loss_inputs(N, 7, B=B, C=C, seed=5, p_obj=0.2)
    z = torch.randn(N, 7, 7, 5 * B + C, generator=torch.Generator().manual_seed(2))
    p64 = torch.sigmoid(z.double())
    o_terms, o_grad = O.loss(p64.float().numpy(), target.numpy(), B=B, C=C, batch_size=N)
    want = o_grad.astype(np.float64) * (p64 * (1 - p64)).numpy()
    _, grad, terms = y.yolo_loss_fused(z.cuda(), target.cuda(), batch_size=N, B=B, C=C, from_logits=True)
    assert np.all(np.abs(terms.cpu().numpy() - o_terms) <= TOL * np.abs(o_terms) + 1e-7), (terms, o_terms)
    err = np.abs(grad.cpu().numpy() - want).max() / np.abs(want).max()
    assert err <= TOL, err
    pb = torch.sigmoid(z).to(torch.bfloat16)
    o_terms, o_grad = O.loss(pb.float().numpy(), target.numpy(), B=B, C=C, batch_size=N)
    _, grad, terms = y.yolo_loss_fused(pb.cuda(), target.cuda(), batch_size=N, B=B, C=C)
    _check(terms, None, o_terms, None, "bf16 C=80")
    assert np.all(np.abs(grad.float().cpu().numpy() - o_grad) <= np.abs(o_grad) * 2.0 ** -8 + 1e-30)


def test_bf16_pred_and_grad():
    """bf16 I/O (config 5): math in fp32 on the bf16-rounded values; compare with the oracle fed the same
    rounded values; the gradient is rounded to bf16 on store (tolerance = bf16 half-ulp, 2^-8 relative)."""
    y = _y()
    pred, target = synth.make_loss_inputs(40, 7, seed=11, p_obj=0.2)
    pb = pred.to(torch.bfloat16)
    o_terms, o_grad = O.loss(pb.float().numpy(), target.numpy(), batch_size=40)
    for variant in (0, 3, -1):
        _, grad, terms = y.yolo_loss_fused(pb.cuda(), target.cuda(), batch_size=40, variant=variant)
        assert grad.dtype == torch.bfloat16
        _check(terms, None, o_terms, None, ("bf16", variant))
        g = grad.float().cpu().numpy()
        assert np.all(np.abs(g - o_grad) <= np.abs(o_grad) * 2.0 ** -8 + 1e-30), variant


def test_module_matches_reference_call_shape_and_autograd():
    """train.py:101,167,171: lossLayer = YOLOLossV1(bs,S,B,C,lc,ln); loss = lossLayer(pred,target); backward."""
    y = _y()
    pred, target = synth.make_loss_inputs(32, 7, seed=20241018)   # BASELINE config 1 shape
    o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=32)
    mod = y.YOLOLossV1(32, 7, 2, 20, 5., .5, _device='cuda:0')
    assert len(list(mod.parameters())) == 0 and len(mod.state_dict()) == 0
    p = pred.cuda().requires_grad_(True)
    loss = mod(p, target.cuda())
    assert loss.dim() == 0 and loss.grad_fn is not None
    loss.backward()
    _check(mod.last_terms, p.grad, o_terms, o_grad, "module")
    # grad_output != 1 (AMP loss scaling): gradient scales, through a sigmoid head like the backbone's
    z = torch.randn(32, 7, 7, 30, device="cuda", requires_grad=True)
    out = torch.sigmoid(z)
    (mod(out, target.cuda()) * 3.0).backward()
    zc = z.detach().cpu().numpy()
    sg = 1.0 / (1.0 + np.exp(-zc.astype(np.float64)))
    _, og = O.loss(sg.astype(np.float32), target.numpy(), batch_size=32)
    want = 3.0 * og * (sg * (1 - sg))
    err = np.abs(z.grad.cpu().numpy() - want).max() / np.abs(want).max()
    assert err <= TOL, err
    # no grad requested -> forward only
    with torch.no_grad():
        l2 = mod(pred.cuda(), target.cuda())
    assert abs(float(l2) - float(o_terms[4])) <= TOL * abs(float(o_terms[4]))


@pytest.mark.parametrize("N,S", [(1, 3), (1, 7), (32, 7), (12, 14), (16, 14), (36, 14), (150, 7), (128, 14)])
def test_small_call_cluster_kernel(N, S):
    """Small calls run as ONE 8-CTA cluster launch that settles the `[:2]` rule before it evaluates a cell and uses no
    workspace (csrc/loss_small.cu): BASELINE config 1 (32 x 7 x 7), train.py:38-41 (12 and 16 images of 14 x 14), an
    odd cell count (generic form), the resident form near its capacity (36 x 196 = 7 056 and 150 x 49 = 7 350 of
    7 360 cells for fp32 NHWC) and a call beyond it (128 x 196: streaming kernels), every layout / dtype / head form,
    K = 0..3 objects, and agreement with the streaming path (variant 31 keeps a small call on the streaming kernels)."""
    y = _y()
    cells = N * S * S
    for p_obj, kind in ((None, "encoder"), (0.3, "mixed")):
        pred, target = synth.make_loss_inputs(N, S, seed=1000 + N + S, p_obj=p_obj, variant=kind)
        o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=N)
        pc, tc = pred.cuda(), target.cuda()
        ws = torch.full((1 << 17,), 0xAB, dtype=torch.uint8, device="cuda")       # never read by the small kernels
        _, grad, terms = y.yolo_loss_fused(pc, tc, batch_size=N, workspace=ws)
        _check(terms, grad, o_terms, o_grad, ("small", N, S, kind))
        if cells <= 7360:
            assert bool((ws == 0xAB).all()), "the small-call kernels must not touch the workspace"
            _, g30, t30 = y.yolo_loss_fused(pc, tc, batch_size=N, variant=30)
            assert torch.equal(g30, grad) and torch.equal(t30, terms)
        else:
            with pytest.raises(RuntimeError):
                y.yolo_loss_fused(pc, tc, batch_size=N, variant=30)
        _, g31, t31 = y.yolo_loss_fused(pc, tc, batch_size=N, variant=31)      # the streaming kernel on the same call
        _check(t31, g31, o_terms, o_grad, ("streaming", N, S, kind))
        assert torch.allclose(g31, grad, rtol=1e-6, atol=1e-9)
        # the backbone's permuted NCHW view, forward only, bf16, paper mode
        planar = pc.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
        _, gp, tp = y.yolo_loss_fused(planar, tc, batch_size=N)
        assert gp.stride() == planar.stride()
        _check(tp, gp, o_terms, o_grad, ("small planar", N, S, kind))
        pb = pred.to(torch.bfloat16)
        ob_terms, ob_grad = O.loss(pb.float().numpy(), target.numpy(), batch_size=N)
        for view in (pb.cuda(), pb.cuda().permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)):
            _, gb, tb = y.yolo_loss_fused(view, tc, batch_size=N)
            _check(tb, None, ob_terms, None, ("small bf16", N, S, kind))
            assert np.all(np.abs(gb.float().cpu().numpy() - ob_grad) <= np.abs(ob_grad) * 2.0 ** -8 + 1e-30)
        _, g0, t0 = y.yolo_loss_fused(pc, tc, batch_size=N, want_grad=False)
        assert g0 is None and torch.equal(t0, terms)
        op_terms, op_grad = O.loss(pred.numpy(), target.numpy(), batch_size=N, coord_mode=1)
        _, gq, tq = y.yolo_loss_fused(pc, tc, batch_size=N, coord_mode="paper")
        _check(tq, gq, op_terms, op_grad, ("small paper", N, S, kind))
    # K = 0, 1, 2, 3 objects: the first two take the plain form, the third the square-root form
    pred, _ = synth.make_loss_inputs(N, S, seed=5)
    cells = [(0, 0, 0), (N - 1, S - 1, S - 1), (N // 2, S // 2, 0)]
    for K in range(4):
        target = torch.zeros(N, S, S, 30)
        for (n, i, j) in sorted(set(cells[:K])):
            target[n, i, j, :2] = 1.0
            target[n, i, j, 2:6] = torch.tensor([0.3, 0.6, 0.4, 0.5])
            target[n, i, j, 6:10] = torch.tensor([0.3, 0.6, 0.4, 0.5])
            target[n, i, j, 13] = 1.0
        o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=N)
        _, grad, terms = y.yolo_loss_fused(pred.cuda(), target.cuda(), batch_size=N)
        _check(terms, grad, o_terms, o_grad, ("small K", K, N, S))


def test_small_call_from_object_lists_and_logits():
    y = _y()
    from yolo_v1_b200 import encode as E
    N, S = 12, 14
    pred, _ = synth.make_loss_inputs(N, S, seed=9)
    g = torch.Generator().manual_seed(4)
    boxes = [torch.rand(3, 4, generator=g) * 0.8 + 0.1 for _ in range(N)]
    labels = [torch.randint(0, 20, (3,), generator=g) for _ in range(N)]
    bx, lb, offs = y.pack_objects(boxes, labels)
    target = y.encode_targets(bx.cuda(), lb.cuda(), offs.cuda(), S)
    o_terms, o_grad = O.loss(pred.numpy(), target.cpu().numpy(), batch_size=N)
    _, grad, terms = y.yolo_loss_from_objects(pred.cuda(), bx.cuda(), lb.cuda(), offs.cuda(), batch_size=N)
    _check(terms, grad, o_terms, o_grad, "small object lists")
    z = torch.randn(N, S, S, 30, generator=g)
    p64 = torch.sigmoid(z.double())
    o_terms, o_grad = O.loss(p64.float().numpy(), target.cpu().numpy(), batch_size=N)
    want = o_grad.astype(np.float64) * (p64 * (1 - p64)).numpy()
    _, grad, terms = y.yolo_loss_fused(z.cuda(), target, batch_size=N, from_logits=True)
    assert np.all(np.abs(terms.cpu().numpy() - o_terms) <= TOL * np.abs(o_terms) + 1e-7)
    assert np.abs(grad.cpu().numpy() - want).max() <= TOL * np.abs(want).max()


def test_graphed_loss_and_allreduce_replay_with_new_inputs():
    """the loss kernel + the NCCL all-reduce of its terms captured in ONE CUDA graph
    (yolo_v1_b200/graph.py); replays with new inputs match the oracle.  One GPU: a world-size-1 NCCL group, so the
    captured graph really holds the collective."""
    import socket
    import torch.distributed as dist
    y = _y()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    started = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
        started = True
    try:
        N, S = 12, 14                                   # train.py:38-41
        g = y.GraphedLoss(N, S, 2, 20)
        for it in range(4):
            pred, target = synth.make_loss_inputs(N, S, seed=300 + it, p_obj=0.05)
            o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=N)
            p = pred.cuda().requires_grad_(True)
            loss = g(p, target.cuda())
            assert loss.grad_fn is not None
            loss.backward()
            _check(g.global_terms, p.grad, o_terms, o_grad, ("graphed", it))
            assert abs(float(loss) - float(o_terms[4])) <= TOL * abs(float(o_terms[4]))
        assert g.graph is not None and g.serial == 4
        # the permuted NCHW view with bf16 logits (what the backbone hands over in config 5) needs its own capture
        g2 = y.GraphedLoss(N, S, 2, 20, from_logits=True)
        z = (torch.randn(N, 30, S, S, generator=torch.Generator().manual_seed(1)) * 1.5).to(torch.bfloat16)
        _, target = synth.make_loss_inputs(N, S, seed=77, p_obj=0.05)
        zv = z.cuda().permute(0, 2, 3, 1).requires_grad_(True)
        g2(zv, target.cuda()).backward()
        _, want_g, want_t = y.yolo_loss_fused(zv.detach(), target.cuda(), batch_size=N, from_logits=True)
        assert torch.equal(g2.global_terms, want_t) and torch.equal(zv.grad, want_g)
        with pytest.raises(ValueError):
            g2(zv.detach().contiguous(), target.cuda())
        g.release(), g2.release()
    finally:
        if started:
            dist.destroy_process_group()


def test_second_backward_raises_and_retain_graph_mode():
    """The reference's autograd graph (v1Loss.py:22-118) can be back-propagated again under retain_graph=True and
    accumulates; without it PyTorch raises.  The fused module hands its gradient buffer over once: a second
    backward() must raise (never a silent None), and `retain_graph=True` at construction gives the reference
    behaviour, including repeated autograd.grad with different grad_outputs."""
    y = _y()
    pred, target = synth.make_loss_inputs(8, 7, seed=3, p_obj=0.2)
    _, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=8)
    tc = target.cuda()
    mod = y.YOLOLossV1(8, 7, 2, 20)
    p = pred.cuda().requires_grad_(True)
    loss = mod(p, tc)
    loss.backward(retain_graph=True)
    g1 = p.grad.clone()
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()
    assert torch.equal(p.grad, g1)                      # the failed call left the accumulated gradient alone
    # scaled first backward (AMP): in place on the handed-over buffer, and the second call still raises
    p2 = pred.cuda().requires_grad_(True)
    l2 = mod(p2, tc)
    (g,) = torch.autograd.grad(l2, p2, grad_outputs=torch.tensor(4.0, device="cuda"), retain_graph=True)
    assert np.abs(g.cpu().numpy() - 4.0 * o_grad).max() <= TOL * 4.0 * np.abs(o_grad).max()
    with pytest.raises(RuntimeError, match="second time"):
        torch.autograd.grad(l2, p2)
    # retain mode: accumulation over two backward passes, and autograd.grad twice with different scales
    modr = y.YOLOLossV1(8, 7, 2, 20, retain_graph=True)
    p3 = pred.cuda().requires_grad_(True)
    l3 = modr(p3, tc)
    l3.backward(retain_graph=True)
    l3.backward(retain_graph=True)
    assert np.abs(p3.grad.cpu().numpy() - 2.0 * o_grad).max() <= TOL * 2.0 * np.abs(o_grad).max()
    (ga,) = torch.autograd.grad(l3, p3, grad_outputs=torch.tensor(3.0, device="cuda"), retain_graph=True)
    (gb,) = torch.autograd.grad(l3, p3, grad_outputs=torch.tensor(0.5, device="cuda"), retain_graph=True)
    assert np.abs(ga.cpu().numpy() - 3.0 * o_grad).max() <= TOL * 3.0 * np.abs(o_grad).max()
    assert np.abs(gb.cpu().numpy() - 0.5 * o_grad).max() <= TOL * 0.5 * np.abs(o_grad).max()   # no compounding
    # host tensors take the same rules
    ph = pred.clone().requires_grad_(True)
    lh = mod(ph, target)
    lh.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second time"):
        lh.backward()
    assert np.abs(ph.grad.numpy() - o_grad).max() <= TOL * np.abs(o_grad).max()


@pytest.mark.parametrize("layout", ["nhwc", "planar", "strided"])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_fused_sigmoid_head(layout, dtype):
    """from_logits: the head's sigmoid (OriginResNet.py:188) and its backward inside the loss kernel.  Same terms
    as the un-fused call on sigmoid(z); gradient = d loss / d p * p (1 - p)."""
    y = _y()
    N, S = 37, 7
    _, target = synth.make_loss_inputs(N, S, seed=5, p_obj=0.2, variant="mixed")
    z = torch.randn(N, S, S, 30, generator=torch.Generator().manual_seed(1)) * 1.5
    if dtype == "bf16":
        z = z.to(torch.bfloat16)
    zd = z.double()
    p64 = torch.sigmoid(zd)
    o_terms, o_grad = O.loss(p64.float().numpy(), target.numpy(), batch_size=N)
    want = o_grad.astype(np.float64) * (p64 * (1 - p64)).numpy()
    zc = z.cuda()
    if layout == "planar":
        zc = zc.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    elif layout == "strided":
        zc = torch.zeros(N, S, S, 32, dtype=z.dtype, device="cuda")[..., :30].copy_(z)
    _, grad, terms = y.yolo_loss_fused(zc, target.cuda(), batch_size=N, from_logits=True)
    assert grad.stride() == zc.stride() or layout == "strided"
    tol = TOL if dtype == "f32" else 2.0 ** -7     # north_star: 1e-5 relative in fp32; bf16: the store's half ulp
    t = terms.cpu().numpy()
    assert np.all(np.abs(t - o_terms) <= TOL * np.abs(o_terms) + 1e-7), (t, o_terms)
    g = grad.float().cpu().numpy()
    err = np.abs(g - want).max() / np.abs(want).max()
    print("fused head %s %s: grad rel err %.3e, terms rel err %.3e" % (layout, dtype, err, np.abs(t / o_terms - 1).max()))
    assert err <= tol, err
    # module form, through autograd, with the pre-sigmoid tensor as the leaf
    if dtype == "f32":
        mod = y.YOLOLossV1(N, S, 2, 20, from_logits=True)
        leaf = zc.clone().requires_grad_(True)
        mod(leaf, target.cuda()).backward()
        assert torch.equal(leaf.grad, grad)


def test_module_logging_hooks():
    y = _y()
    pred, target = synth.make_loss_inputs(4, 7, seed=1, p_obj=0.2)

    class Log:
        def __init__(self):
            self.lines = []

        def info(self, s):
            self.lines.append(s)

    class Vis:
        def __init__(self):
            self.calls = []

        def plot(self, name, v):
            self.calls.append((name, v))

    lg, vis = Log(), Vis()
    mod = y.YOLOLossV1(4, 7, 2, 20, _logger=lg, _vis=vis)
    mod(pred.cuda(), target.cuda())
    o_terms, _ = O.loss(pred.numpy(), target.numpy(), batch_size=4)
    import re
    assert len(lg.lines) == 1
    m = re.fullmatch(r'location loss : (\S+) contain loss : (\S+) not contain loss: (\S+) classify loss : (\S+)',
                     lg.lines[0])                                   # the format of v1Loss.py:108
    assert m and np.allclose([float(v) for v in m.groups()], o_terms[:4], atol=2e-5)
    assert np.allclose([c[1] for c in vis.calls], o_terms[:4], rtol=1e-5)
    assert [c[0] for c in vis.calls] == ['location loss', 'confidence loss', 'no object loss', 'classify loss']


def test_deterministic_and_stream_ordered():
    y = _y()
    pred, target = synth.make_loss_inputs(4096, 7, seed=9)
    p, t = pred.cuda(), target.cuda()
    _, g1, t1 = y.yolo_loss_fused(p, t, batch_size=4096)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        _, g2, t2 = y.yolo_loss_fused(p, t, batch_size=4096)
    s.synchronize()
    assert torch.equal(t1, t2) and torch.equal(g1, g2)


def test_host_buffer_path_equals_device_path_and_spans_chunks():
    """yolo1_loss_fwd_bwd_host: chunked H2D/kernel/D2H pipeline; the `[:2]` rule (v1Loss.py:101) must span
    chunk boundaries: objects placed so that the call's 1st, 2nd and 3rd objects sit in different chunks."""
    y = _y()
    N, S = 23, 7
    pred, target = synth.make_loss_inputs(N, S, seed=123, p_obj=0.0)
    g = torch.Generator().manual_seed(4)
    for (n, i, j) in [(1, 2, 3), (9, 0, 0), (10, 6, 6), (22, 3, 3)]:
        target[n, i, j, :2] = 1
        bx = torch.rand(4, generator=g) * 0.8 + 0.1
        target[n, i, j, 2:6] = bx
        target[n, i, j, 6:10] = bx
        target[n, i, j, 10 + (n % 20)] = 1
    o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=N)
    for chunk in (0, 4, 1, 23, 100):
        ctx = y.HostContext(S, chunk_images=chunk)
        terms, grad = ctx.loss(pred, target, batch_size=N)
        _check(terms, grad, o_terms, o_grad, ("host", chunk))
        # pinned + mapped buffers: the zero-copy kernel reads / writes host memory in place; staged path forced too
        pp, tp = pred.pin_memory(), target.pin_memory()
        gp = torch.full(pred.shape, float("nan")).pin_memory()
        for zc in (1, 2, 3, 0):
            ctx.set_zero_copy(zc)
            gp.fill_(float("nan"))
            t2, g2 = ctx.loss(pp, tp, batch_size=N, out_grad=gp)
            assert g2 is gp
            _check(t2, g2, o_terms, o_grad, ("host pinned", chunk, zc))
            t3, g3 = ctx.loss(pp, tp, batch_size=N, want_grad=False)
            assert g3 is None
            _check(t3, None, o_terms, None, ("host pinned fwd", chunk, zc))
        ctx.close()
    # the module accepts CPU tensors too (reference `_device='cpu'` call shape); arithmetic still on the GPU
    mod = y.YOLOLossV1(N, S, 2, 20, _device='cpu')
    p = pred.clone().requires_grad_(True)
    mod(p, target).backward()
    _check(mod.last_terms, p.grad, o_terms, o_grad, "host module")


def test_zero_copy_host_path_large_ragged_batch():
    """The host-mapped kernel on a batch with a ragged tail (N*S*S % 128 != 0) against the device path."""
    y = _y()
    N, S = 2049, 7
    pred, target = synth.make_loss_inputs(N, S, seed=77)
    _, gd, td = y.yolo_loss_fused(pred.cuda(), target.cuda(), batch_size=N)
    ctx = y.HostContext(S)
    pp, tp = pred.pin_memory(), target.pin_memory()
    gp = torch.empty(pred.shape).pin_memory()
    th, gh = ctx.loss(pp, tp, batch_size=N, out_grad=gp)
    assert torch.allclose(th, td.cpu(), rtol=2e-6, atol=0)
    assert torch.equal(gh, gd.cpu())
    ctx.close()


def test_full_size_properties_config3():
    """BASELINE config 3 (N=65536, S=14: 12.8 M cells, 1.54 GB per tensor).  The oracle checks a 512-image
    sub-batch; the rest is covered by properties: the gradient is local (a sub-batch run reproduces the
    slice bit for bit outside the first two object cells), and the raw sums are additive over a split."""
    y = _y()
    N, S = 65536, 14
    pred, target = synth.make_loss_inputs(N, S, seed=20241018 + 3000, device="cuda")
    _, grad, terms = y.yolo_loss_fused(pred, target, batch_size=N, coord_mode="paper")
    # additivity (paper mode has no call-order dependence): halves sum to the whole
    h = N // 2
    _, _, ta = y.yolo_loss_fused(pred[:h], target[:h], batch_size=N, coord_mode="paper")
    _, _, tb = y.yolo_loss_fused(pred[h:], target[h:], batch_size=N, coord_mode="paper")
    assert torch.allclose(ta + tb, terms, rtol=2e-6, atol=0)
    # locality + oracle on a sub-batch from the middle
    a, b = 30000, 30512
    _, gs, ts = y.yolo_loss_fused(pred[a:b], target[a:b], batch_size=N, coord_mode="paper")
    assert torch.equal(gs, grad[a:b])
    o_terms, o_grad = O.loss(pred[a:b].cpu().numpy(), target[a:b].cpu().numpy(), batch_size=N, coord_mode=1)
    _check(ts, gs, o_terms, o_grad, "config3 sub-batch")
    # reference mode: identical except the location term / the first two object cells of the call
    _, gr, tr = y.yolo_loss_fused(pred, target, batch_size=N)
    diff = (gr != grad).reshape(-1, 30).any(dim=1).nonzero().reshape(-1)
    objs = (target[..., 0].reshape(-1) == 1).nonzero().reshape(-1)
    assert diff.numel() == objs.numel()              # paper vs reference differ on every object cell (xy vs sqrt)
    o_small, og_small = O.loss(pred[:64].cpu().numpy(), target[:64].cpu().numpy(), batch_size=N)
    assert np.abs(gr[:64].cpu().numpy() - og_small).max() <= TOL * np.abs(og_small).max()
    assert torch.equal(tr[1:4], terms[1:4])


def test_cuda_graph_capture_and_replay():
    """The C-ABI calls are stream ordered, allocation free and sync free, so a step (loss fwd+bwd, decode+NMS) can be
    captured once into a CUDA graph and replayed -- what makes the launch-bound small-batch case (BASELINE config 1:
    N=32, S=7) cheap."""
    y = _y()
    N, S = 32, 7
    pred, target = synth.make_loss_inputs(N, S, seed=20241018, device="cuda")
    grad, terms = torch.empty_like(pred), torch.empty(5, device="cuda")
    ws = torch.empty(1 << 17, dtype=torch.uint8, device="cuda")
    M = S * S * 2
    outs = (torch.empty((N, M, 4), device="cuda"), torch.empty((N, M), dtype=torch.int32, device="cuda"),
            torch.empty((N, M), device="cuda"), torch.empty((N,), dtype=torch.int32, device="cuda"))

    def step():
        y.yolo_loss_fused(pred, target, batch_size=N, out_grad=grad, out_terms=terms, workspace=ws)
        y.decode_nms_batched(pred, 0.1, 0.5, out=outs)

    step()
    torch.cuda.synchronize()
    want = (grad.clone(), terms.clone(), outs[0].clone(), outs[3].clone())
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step()                       # warm-up on the capture stream
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            step()
    # new inputs, same buffers: the replay must see them
    p2, t2 = synth.make_loss_inputs(N, S, seed=7, device="cuda")
    pred.copy_(p2), target.copy_(t2)
    g.replay()
    torch.cuda.synchronize()
    o_terms, o_grad = O.loss(p2.cpu().numpy(), t2.cpu().numpy(), batch_size=N)
    _check(terms, grad, o_terms, o_grad, "graph replay")
    assert not torch.equal(grad, want[0])
    orc = O.decode_nms(p2.cpu().numpy(), thresh=0.1, nms_th=0.5)
    assert np.array_equal(outs[3].cpu().numpy(), orc["counts"])
    assert np.array_equal(outs[0].cpu().numpy().view(np.uint32), orc["boxes"].view(np.uint32))


def test_integration_md_ctypes_stub_runs_as_written():
    """The binding INTEGRATION.md shows a maintainer of the reference (a ctypes stub against the C ABI) is executed
    verbatim (only the library path is substituted) and must reproduce the oracle."""
    import re
    from yolo_v1_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    stub = [b for b in blocks if "class _Loss" in b]
    assert len(stub) == 1
    code = stub[0].replace('"libyolo1_b200.so"', repr(_lib.SO_PATH))
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    mod = ns["YOLOLossV1"]()
    mod.S, mod.B, mod.C, mod.batch_size, mod.lambda_coord, mod.lambda_noobj = 7, 2, 20, 32, 5., .5
    pred, target = synth.make_loss_inputs(32, 7, seed=20241018)
    p = pred.cuda().requires_grad_(True)
    loss = mod(p, target.cuda())
    loss.backward()
    o_terms, o_grad = O.loss(pred.numpy(), target.numpy(), batch_size=32)
    assert abs(float(loss) - float(o_terms[4])) <= TOL * abs(float(o_terms[4]))
    assert np.abs(p.grad.cpu().numpy() - o_grad).max() <= TOL * np.abs(o_grad).max()
