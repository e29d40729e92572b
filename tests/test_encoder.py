"""Target encoder: oracle vs the golden vectors the reference's own
`yoloDataset.encoder` produced (CPU), and the CUDA kernel vs both (GPU).  Bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O


def _cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "encoder_cases.npz"))
    for name in sorted({k.split("/")[0] for k in z.files}):
        S, B, C = [int(v) for v in z[name + "/params"]]
        yield name, S, B, C, z[name + "/boxes"], z[name + "/labels"], z[name + "/offsets"], z[name + "/target"]


def test_oracle_encoder_matches_reference_golden(golden_dir):
    n = 0
    for name, S, B, C, boxes, labels, offsets, target in _cases(golden_dir):
        got = O.encode(boxes, labels, offsets, S, B, C)
        assert np.array_equal(got.view(np.uint32), target.view(np.uint32)), name
        n += 1
    assert n >= 3
    with pytest.raises(IndexError):
        O.encode(np.array([[1.5, 0.5, 0.1, 0.1]], np.float32), [0], [0, 1])


def test_encoded_targets_round_trip_through_the_oracle_decoder(golden_dir):
    """decoder(encoder(boxes), gt=True) returns the boxes (the visual check of YOLODataLoader.py:233-257)."""
    rng = np.random.RandomState(0)
    boxes = np.array([[0.21, 0.33, 0.2, 0.1], [0.77, 0.61, 0.3, 0.25]], np.float32)
    t = O.encode(boxes, [4, 11], [0, 2])[0]
    b, c, s = O.decoder(t, gt=True)
    want = np.stack([boxes[:, 0] - boxes[:, 2] / 2, boxes[:, 1] - boxes[:, 3] / 2,
                     boxes[:, 0] + boxes[:, 2] / 2, boxes[:, 1] + boxes[:, 3] / 2], 1)
    got = {int(ci): bi for bi, ci in zip(b, c)}     # both slots decode to the same box; NMS(1.0) keeps them
    assert np.allclose(got[4], want[0], atol=1e-6) and np.allclose(got[11], want[1], atol=1e-6)


@pytest.mark.gpu
def test_cuda_encoder_bit_exact(golden_dir):
    import yolo_v1_b200 as y
    for name, S, B, C, boxes, labels, offsets, target in _cases(golden_dir):
        got = y.encode_targets(torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(),
                               torch.from_numpy(offsets).cuda(), S, B, C)
        assert np.array_equal(got.cpu().numpy().view(np.uint32), target.view(np.uint32)), name
    # single-image call shape of the reference method
    t = y.encoder(torch.tensor([[0.5, 0.5, 0.2, 0.3], [0.51, 0.52, 0.1, 0.1]]), torch.tensor([3, 5]))
    assert t.shape == (7, 7, 30) and float(t[3, 3, 15]) == 1.0 and float(t[3, 3, 13]) == 0.0
    # out-of-grid centre: IndexError like the reference; check=False skips it
    with pytest.raises(IndexError):
        y.encoder(torch.tensor([[1.5, 0.5, 0.1, 0.1]]), torch.tensor([0]))
    b, l, o = y.pack_objects([torch.tensor([[1.5, 0.5, 0.1, 0.1], [0.2, 0.2, 0.1, 0.1]])], [torch.tensor([0, 1])])
    t = y.encode_targets(b, l, o, check=False)
    assert int((t[..., 0] == 1).sum()) == 1


@pytest.mark.gpu
def test_cuda_encoder_large_random_batches_vs_oracle_and_loss_consumes_it():
    import yolo_v1_b200 as y
    rng = np.random.RandomState(3)
    for S, N in [(7, 4097), (14, 1025), (5, 33), (7, 30001), (14, 6001)]:      # the last two: several tiles per CTA
        counts = rng.randint(0, 8, size=N)
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        boxes = rng.rand(int(offsets[-1]), 4).astype(np.float32)
        boxes[:, 2:] = boxes[:, 2:] * 0.8 + 0.05
        labels = rng.randint(0, 20, size=int(offsets[-1])).astype(np.int32)
        want = O.encode(boxes, labels, offsets, S)
        got = y.encode_targets(torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(),
                               torch.from_numpy(offsets).cuda(), S)
        assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32)), (S, N)
        # the encoded target drives the loss exactly like a host-encoded one
        pred = torch.rand(N, S, S, 30, generator=torch.Generator().manual_seed(S)) * 0.98 + 0.01
        o_terms, o_grad = O.loss(pred.numpy(), want, batch_size=N)
        _, grad, terms = y.yolo_loss_fused(pred.cuda(), got, batch_size=N)
        assert np.allclose(terms.cpu().numpy(), o_terms, rtol=1e-5)
        assert np.abs(grad.cpu().numpy() - o_grad).max() <= 1e-5 * np.abs(o_grad).max()
    # empty batch / no objects at all
    t = y.encode_targets(torch.zeros(0, 4).cuda(), torch.zeros(0, dtype=torch.int32).cuda(),
                         torch.zeros(5, dtype=torch.int64).cuda(), 7)
    assert t.shape == (4, 7, 7, 30) and float(t.abs().sum()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["nhwc", "planar", "strided"])
@pytest.mark.parametrize("dtype,logits", [("f32", False), ("bf16", False), ("f32", True)])
def test_loss_from_object_lists_equals_loss_on_encoded_target(layout, dtype, logits):
    """yolo1_loss_fwd_bwd_objects == yolo1_loss_fwd_bwd(pred, encoder(lists)): same gradient bit for bit, same
    terms (up to the summation order of a different grid), and both match the oracle."""
    import yolo_v1_b200 as y
    rng = np.random.RandomState(11)
    for S, N in [(7, 131), (14, 37), (7, 2)]:
        counts = rng.randint(0, 6, size=N)
        counts[0] = 0                                     # first image empty: the call's first objects come later
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        boxes = rng.rand(int(offsets[-1]), 4).astype(np.float32)
        boxes[:, 2:] = boxes[:, 2:] * 0.8 + 0.05
        if len(boxes) > 3:
            boxes[1, :2] = boxes[0, :2]                   # collision: the last object in a cell wins
            boxes[2, :2] = [0.0, 1.0]                     # Python -1 indexing / edge
        labels = rng.randint(0, 20, size=int(offsets[-1])).astype(np.int32)
        dense = O.encode(boxes, labels, offsets, S)
        g = torch.Generator().manual_seed(S + N)
        pred = torch.randn(N, S, S, 30, generator=g) * 1.5 if logits else torch.rand(N, S, S, 30, generator=g) * 0.98 + 0.01
        if dtype == "bf16":
            pred = pred.to(torch.bfloat16)
        pc = pred.cuda()
        if layout == "planar":
            pc = pc.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
        elif layout == "strided":
            pc = torch.zeros(N, S, S, 32, dtype=pred.dtype, device="cuda")[..., :30].copy_(pred)
        b, l, o = torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(offsets).cuda()
        _, g1, t1 = y.yolo_loss_from_objects(pc, b, l, o, batch_size=N, from_logits=logits, check=True)
        _, g2, t2 = y.yolo_loss_fused(pc, torch.from_numpy(dense).cuda(), batch_size=N, from_logits=logits)
        assert torch.equal(g1, g2), (S, N, layout, dtype, logits)
        assert torch.allclose(t1, t2, rtol=2e-6, atol=1e-7)
        _, _, t3 = y.yolo_loss_from_objects(pc, b, l, o, batch_size=N, from_logits=logits, want_grad=False)
        assert torch.allclose(t3, t1, rtol=2e-6, atol=1e-7)
        if not logits and dtype == "f32":
            o_terms, o_grad = O.loss(pred.numpy(), dense, batch_size=N)
            assert np.allclose(t1.cpu().numpy(), o_terms, rtol=1e-5)
            assert np.abs(g1.cpu().numpy() - o_grad).max() <= 1e-5 * max(np.abs(o_grad).max(), 1e-12)
    # module form with autograd; out-of-grid centre -> IndexError with check=True
    mod = y.YOLOLossV1(4, 7, 2, 20)
    p = torch.rand(4, 7, 7, 30, device="cuda", requires_grad=True)
    bx, lb, of = y.pack_objects([torch.tensor([[0.3, 0.3, 0.2, 0.2]]), torch.zeros(0, 4), torch.tensor([[0.9, 0.1, 0.1, 0.1]]),
                                 torch.zeros(0, 4)], [torch.tensor([1]), torch.zeros(0), torch.tensor([2]), torch.zeros(0)])
    mod.forward_objects(p, bx, lb, of).backward()
    want = y.yolo_loss_fused(p.detach(), y.encode_targets(bx, lb, of), batch_size=4)[1]
    assert torch.equal(p.grad, want)
    with pytest.raises(IndexError):
        y.yolo_loss_from_objects(p.detach(), torch.tensor([[1.5, 0.5, 0.1, 0.1]]).cuda(), torch.tensor([0]).cuda(),
                                 torch.tensor([0, 1, 1, 1, 1]).cuda(), batch_size=4, check=True)


@pytest.mark.gpu
def test_loss_from_object_lists_empty_cases():
    import yolo_v1_b200 as y
    # no images at all
    e = torch.zeros(0, 4).cuda(), torch.zeros(0, dtype=torch.int32).cuda()
    _, g, t = y.yolo_loss_from_objects(torch.zeros(0, 7, 7, 30, device="cuda"), e[0], e[1],
                                       torch.zeros(1, dtype=torch.int64).cuda(), batch_size=1)
    assert g.numel() == 0 and float(t.abs().sum()) == 0
    # images without any object: only the no-object confidence term remains
    pred = torch.rand(5, 7, 7, 30, generator=torch.Generator().manual_seed(0)) * 0.98 + 0.01
    _, g, t = y.yolo_loss_from_objects(pred.cuda(), e[0], e[1], torch.zeros(6, dtype=torch.int64).cuda(), batch_size=5)
    o_terms, o_grad = O.loss(pred.numpy(), np.zeros((5, 7, 7, 30), np.float32), batch_size=5)
    assert np.allclose(t.cpu().numpy(), o_terms, rtol=1e-5) and o_terms[0] == 0 and o_terms[3] == 0
    assert np.abs(g.cpu().numpy() - o_grad).max() <= 1e-5 * np.abs(o_grad).max()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,logits", [("f32", False), ("bf16", False), ("f32", True)])
def test_object_list_owners_found_in_the_streaming_kernel(dtype, logits):
    """Calls beyond the small-call sizes on contiguous tensors skip the pre-pass: warp 0 of the streaming kernel finds
    the owner of every cell from the lists (variant 60 forces that form, 61 the pre-pass + map).  Both forms must agree
    bit for bit, with the dense-target call on encoder(lists), and with the oracle -- including images without
    objects, images with more objects than the 32 lanes of the pipeline, collisions, the Python -1 index, a ragged
    last tile, and an out-of-grid object (skipped, reported through the status word)."""
    import yolo_v1_b200 as y
    rng = np.random.RandomState(5)
    for S, N in [(14, 203), (7, 701), (3, 2000)]:
        counts = rng.randint(0, 6, size=N)
        counts[0] = 0
        counts[5] = 45                                    # more than one warp of objects in a tile's images
        counts[6] = 70
        counts[N - 1] = 3                                 # objects in the ragged tail
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        n_obj = int(offsets[-1])
        boxes = rng.rand(n_obj, 4).astype(np.float32)
        boxes[:, 2:] = boxes[:, 2:] * 0.8 + 0.05
        boxes[1, :2] = boxes[0, :2]                       # collision: the last object in a cell wins
        boxes[2, :2] = [0.0, 1.0]                         # Python -1 indexing / edge
        labels = rng.randint(0, 20, size=n_obj).astype(np.int32)
        dense = O.encode(boxes, labels, offsets, S)
        g = torch.Generator().manual_seed(S + N)
        pred = torch.randn(N, S, S, 30, generator=g) * 1.5 if logits else torch.rand(N, S, S, 30, generator=g) * 0.98 + 0.01
        if dtype == "bf16":
            pred = pred.to(torch.bfloat16)
        pc = pred.cuda()
        b, l, o = torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(offsets).cuda()
        _, g0, t0 = y.yolo_loss_from_objects(pc, b, l, o, batch_size=N, from_logits=logits, check=True)       # default
        _, g1, t1 = y.yolo_loss_from_objects(pc, b, l, o, batch_size=N, from_logits=logits, variant=60, check=True)
        _, g2, t2 = y.yolo_loss_from_objects(pc, b, l, o, batch_size=N, from_logits=logits, variant=61, check=True)
        _, g3, t3 = y.yolo_loss_fused(pc, torch.from_numpy(dense).cuda(), batch_size=N, from_logits=logits)
        assert torch.equal(g1, g2) and torch.equal(g0, g1) and torch.equal(g1, g3), (S, N, dtype, logits)
        assert torch.equal(t1, t2) and torch.equal(t0, t1), (t1, t2)
        assert torch.allclose(t1, t3, rtol=2e-6, atol=1e-7)
        _, _, t4 = y.yolo_loss_from_objects(pc, b, l, o, batch_size=N, from_logits=logits, variant=60, want_grad=False)
        assert torch.allclose(t4, t1, rtol=2e-6, atol=1e-7)
        if not logits and dtype == "f32":
            o_terms, o_grad = O.loss(pred.numpy(), dense, batch_size=N)
            assert np.allclose(t1.cpu().numpy(), o_terms, rtol=1e-5)
            assert np.abs(g1.cpu().numpy() - o_grad).max() <= 1e-5 * max(np.abs(o_grad).max(), 1e-12)
        # an object outside the grid: skipped by both forms, IndexError with check=True (the reference's behaviour)
        bad = boxes.copy()
        bad[n_obj // 2, 0] = 1.5
        bb = torch.from_numpy(bad).cuda()
        for v in (60, 61):
            with pytest.raises(IndexError):
                y.yolo_loss_from_objects(pc, bb, l, o, batch_size=N, from_logits=logits, variant=v, check=True)
        _, ga, ta = y.yolo_loss_from_objects(pc, bb, l, o, batch_size=N, from_logits=logits, variant=60, check=False)
        _, gb, tb = y.yolo_loss_from_objects(pc, bb, l, o, batch_size=N, from_logits=logits, variant=61, check=False)
        assert torch.equal(ga, gb) and torch.equal(ta, tb)
    # too small for the in-kernel form: forcing it is refused, the default takes the pre-pass
    with pytest.raises(Exception):
        y.yolo_loss_from_objects(pc[:4], b[:0], l[:0], torch.zeros(5, dtype=torch.int64).cuda(), batch_size=4, variant=60)
