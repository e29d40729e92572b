"""Seeded synthetic inputs for the YOLO v1 hot path.

`target` tensors follow the format the reference's `encoder` writes
(/root/reference/utils/YOLODataLoader.py:200-230): on an object cell every confidence slot is 1, the
same (dx, dy, w, h) box is copied into every slot and the class is one-hot; every other cell is all
zero.  `pred` mimics sigmoid outputs (the backbone ends in a sigmoid,
/root/reference/backbones/OriginResNet.py:186-189).
"""
import torch


def default_p_obj(S):
    """About 3 objects per image, the VOC density (3/49 at S=7, 3/196 at S=14)."""
    return 3.0 / float(S * S)


def make_loss_inputs(N, S, B=2, C=20, p_obj=None, seed=0, device="cpu", variant="encoder",
                     dtype=torch.float32):
    """Returns (pred, target) of shape [N,S,S,5B+C], contiguous.

    variant 'encoder': GT exactly as `encoder` writes it.  variant 'mixed': slot-b GT boxes differ and
    conf_1.. are random (as `make_eval_tensor` does, /root/reference/utils/utils.py:83-88) -- pins
    "IoU uses GT slot 0, the object mask uses channel 0, the coordinate loss uses GT slot r".
    """
    D = 5 * B + C
    if p_obj is None:
        p_obj = default_p_obj(S)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    pred = torch.rand((N, S, S, D), generator=g, device=device, dtype=torch.float32) * 0.98 + 0.01
    obj = torch.rand((N, S, S), generator=g, device=device) < p_obj
    xy = torch.rand((N, S, S, 2), generator=g, device=device) * 0.98 + 0.01
    wh = torch.rand((N, S, S, 2), generator=g, device=device) * 0.88 + 0.02
    cls = torch.randint(0, C, (N, S, S), generator=g, device=device)
    target = torch.zeros((N, S, S, D), device=device, dtype=torch.float32)
    box = torch.cat([xy, wh], dim=-1)
    objf = obj.to(torch.float32).unsqueeze(-1)
    target[..., :B] = objf
    for b in range(B):
        target[..., B + 4 * b:B + 4 * b + 4] = box * objf
    target[..., 5 * B:] = torch.nn.functional.one_hot(cls, C).to(torch.float32) * objf
    if variant == "mixed":
        for b in range(1, B):
            other = torch.rand((N, S, S, 4), generator=g, device=device) * 0.88 + 0.02
            target[..., B + 4 * b:B + 4 * b + 4] = other * objf
            target[..., b] = torch.rand((N, S, S), generator=g, device=device)
    elif variant != "encoder":
        raise ValueError("variant must be 'encoder' or 'mixed'")
    if dtype != torch.float32:
        pred = pred.to(dtype)
    return pred, target


def make_decode_inputs(N, S, B=2, C=20, seed=0, device="cpu", dist="uniform"):
    """pred [N,S,S,5B+C] for decode/NMS: 'uniform' = U(0,1) (BASELINE config 2), 'sigmoid' =
    sigmoid(2*N(0,1) - 1) (sparser)."""
    D = 5 * B + C
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    if dist == "uniform":
        return torch.rand((N, S, S, D), generator=g, device=device, dtype=torch.float32)
    if dist == "sigmoid":
        z = torch.randn((N, S, S, D), generator=g, device=device, dtype=torch.float32)
        return torch.sigmoid(2.0 * z - 1.0)
    raise ValueError("dist must be 'uniform' or 'sigmoid'")


def score_tie_images(pred, B=2):
    """Indices of images whose candidate scores (conf * max class prob, fp32) contain duplicates.  The
    reference's sort order on ties is machine dependent (unstable torch sort), so bit-exact keep-list
    checks use tie-free images only."""
    N = pred.shape[0]
    conf = pred[..., :B].reshape(N, -1, B)
    mp = pred[..., 5 * B:].max(dim=-1).values.reshape(N, -1, 1)
    sc = (conf * mp).reshape(N, -1)
    s, _ = sc.sort(dim=1)
    dup = (s[:, 1:] == s[:, :-1]).any(dim=1)
    return dup.nonzero().reshape(-1)


def make_tie_free_decode_inputs(N, S, B=2, C=20, seed=0, device="cpu", dist="uniform"):
    """As make_decode_inputs, but images with duplicate scores are re-drawn (seed + 10**6, ...).
    Returns (pred, n_redrawn)."""
    pred = make_decode_inputs(N, S, B, C, seed, device, dist)
    redrawn = 0
    bump = 0
    while True:
        bad = score_tie_images(pred, B)
        if bad.numel() == 0:
            return pred, redrawn
        bump += 1
        redrawn += int(bad.numel())
        fresh = make_decode_inputs(int(bad.numel()), S, B, C, seed + bump * 10 ** 6, device, dist)
        pred[bad] = fresh
