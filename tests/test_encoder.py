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
    for S, N in [(7, 4097), (14, 1025), (5, 33)]:
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
