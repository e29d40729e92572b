"""Target encoder on the GPU: the batched counterpart of the reference's
`yoloDataset.encoder` (/root/reference/utils/YOLODataLoader.py:200-230).  Only the ragged object lists cross
PCIe; the dense [N,S,S,5B+C] target the loss consumes is produced on the device (yolo_v1_b200/csrc/encode.cu).
"""
import ctypes

import torch

from . import _lib

__all__ = ["encode_targets", "encoder", "pack_objects"]


def pack_objects(boxes_list, labels_list, device="cuda"):
    """[boxes_i [k_i,4]], [labels_i [k_i]] per image -> (boxes [n,4] f32, labels [n] i32, offsets [N+1] i64) on
    `device` -- the CSR form `encode_targets` takes.  Host-side list handling only."""
    counts = [int(b.shape[0]) if torch.is_tensor(b) else len(b) for b in boxes_list]
    offsets = torch.zeros(len(counts) + 1, dtype=torch.int64)
    if counts:
        offsets[1:] = torch.tensor(counts, dtype=torch.int64).cumsum(0)
    nz = [torch.as_tensor(b, dtype=torch.float32).reshape(-1, 4) for b in boxes_list]
    lz = [torch.as_tensor(l).reshape(-1).to(torch.int32) for l in labels_list]
    boxes = torch.cat(nz) if nz else torch.zeros((0, 4))
    labels = torch.cat(lz) if lz else torch.zeros((0,), dtype=torch.int32)
    return boxes.to(device), labels.to(device), offsets.to(device)


def encode_targets(boxes, labels, offsets, S=7, B=2, C=20, out=None, check=True):
    """boxes [n,4] (cx,cy,w,h normalised to the image), labels [n], offsets [N+1] (CUDA tensors) ->
    target float32 [N,S,S,5B+C].  With check=True (one host sync) an out-of-grid centre or label raises IndexError
    as the reference does; check=False skips such objects silently and returns without synchronising."""
    if not offsets.is_cuda:
        raise RuntimeError("encode_targets needs CUDA tensors (use pack_objects to move the object lists)")
    dev = offsets.device
    N = int(offsets.shape[0]) - 1
    boxes = boxes.to(device=dev, dtype=torch.float32).contiguous().reshape(-1, 4)
    labels = labels.to(device=dev, dtype=torch.int32).contiguous().reshape(-1)
    offsets = offsets.to(torch.int64).contiguous()
    D = 5 * B + C
    target = out if out is not None else torch.empty((N, S, S, D), dtype=torch.float32, device=dev)
    if tuple(target.shape) != (N, S, S, D) or not target.is_contiguous() or target.dtype != torch.float32:
        raise ValueError("out must be a contiguous float32 [N,S,S,5B+C] tensor")
    status = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().yolo1_encode_targets(
            boxes.data_ptr() if boxes.numel() else None, labels.data_ptr() if labels.numel() else None,
            offsets.data_ptr(), N, S, B, C, target.data_ptr() if N else None, status.data_ptr(),
            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "yolo1_encode_targets")
    if check and N and int(status.item()) != 0:
        raise IndexError("encode_targets: a box centre or label lies outside the grid / class range")
    return target


def encoder(boxes, labels, S=7, B=2, C=20, device="cuda"):
    """One image, the reference method's arguments (utils/YOLODataLoader.py:200): boxes [k,4], labels [k] ->
    target [S,S,5B+C] (on `device`)."""
    b, l, o = pack_objects([boxes], [labels], device)
    return encode_targets(b, l, o, S, B, C)[0]
