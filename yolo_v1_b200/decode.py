"""Host-side mirror of the reference's decode / NMS interface (/root/reference/utils/utils.py:94-184) on
top of the C ABI (include/yolo1_b200.h); kernels in yolo_v1_b200/csrc/decode_nms.cu.

`decoder` and `nms` keep the reference signatures and return conventions (what `run_test_mAP`
utils/utils.py:405, `eval.py:94` and `YOLODataLoader.py:249` call); the batched entry points are additive.
Intentional difference: the reference `decoder` overwrites x,y of every candidate box in the caller's
`pred` (utils/utils.py:119,123) -- no caller reads `pred` afterwards; this implementation leaves it intact.
"""
import ctypes

import torch

from . import _lib

__all__ = ["decoder", "nms", "decode_nms_batched", "decode_batched", "nms_batched", "compute_iou_matrix",
           "convert_CxCyWH_to_X1Y1X2Y2"]


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _cuda_device(device=None):
    if device is not None and torch.device(device).type == "cuda":
        d = torch.device(device)
        return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())
    if not torch.cuda.is_available():
        raise RuntimeError("yolo_v1_b200 has no CPU implementation: a CUDA device is required")
    return torch.device("cuda", torch.cuda.current_device())


def _pred_dtype(t):
    if t.dtype == torch.float32:
        return _lib.DTYPE_F32
    if t.dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    raise TypeError("pred must be float32 or bfloat16, got %s" % t.dtype)


def _check_pred(pred, B):
    if pred.dim() == 3:
        pred = pred.unsqueeze(0)
    if pred.dim() != 4 or pred.shape[1] != pred.shape[2]:
        raise ValueError("pred must be [N,S,S,5B+C], got %s" % (tuple(pred.shape),))
    N, S, _, D = pred.shape
    C = D - 5 * B
    if C <= 0:
        raise ValueError("pred has %d channels, fewer than 5*B+1" % D)
    if not pred.is_cuda:
        raise RuntimeError("batched decode needs a CUDA tensor (host tensors: yolo_v1_b200.host.HostContext)")
    return pred, N, S, C


def decode_nms_batched(pred, thresh=0.3, nms_th=0.5, class_agnostic=True, B=2, return_keep=False, out=None):
    """decoder (utils/utils.py:94-147) for a whole batch in one launch, one CTA per image.

    pred [N,S,S,5B+C] CUDA float32/bfloat16, any strides.  Returns (boxes [N,M,4] xyxy normalised to the
    image, cls [N,M] int32, probs [N,M], counts [N] int32), M = S*S*B; image n's detections are rows
    [0, counts[n]) in descending score order; remaining rows are zero.  counts[n] == 0 is the case where
    the reference returns its all-zero sentinel.  class_agnostic=True is the reference behaviour (one NMS
    over all classes, utils/utils.py:146); False suppresses only within a class.
    With return_keep also returns (keep_idx [N,M] = kept candidate indices in emission order, as nms()
    returns them, cand_counts [N] = candidates before NMS).
    """
    pred, N, S, C = _check_pred(pred, B)
    M = S * S * B
    dev = pred.device
    if out is None:
        out = (torch.empty((N, M, 4), dtype=torch.float32, device=dev),
               torch.empty((N, M), dtype=torch.int32, device=dev),
               torch.empty((N, M), dtype=torch.float32, device=dev),
               torch.empty((N,), dtype=torch.int32, device=dev))
    boxes, cls, probs, counts = out
    keep = cand = None
    if return_keep:
        keep = torch.empty((N, M), dtype=torch.int32, device=dev)
        cand = torch.empty((N,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().yolo1_decode_nms(
            pred.data_ptr(), _lib.strides4(pred), _pred_dtype(pred), N, S, B, C, float(thresh), float(nms_th),
            0 if class_agnostic else 1, boxes.data_ptr(), probs.data_ptr(), cls.data_ptr(), counts.data_ptr(),
            keep.data_ptr() if keep is not None else None, cand.data_ptr() if cand is not None else None,
            _stream_ptr(dev))
        _lib.check(rc, "yolo1_decode_nms")
    if return_keep:
        return boxes, cls, probs, counts, keep, cand
    return boxes, cls, probs, counts


def decode_batched(pred, thresh=0.3, B=2):
    """Candidate stage of decoder only (utils/utils.py:108-132): (boxes [N,M,4], scores [N,M], cls [N,M] int32,
    counts [N] int32) in row-major (i, j, b) emission order."""
    pred, N, S, C = _check_pred(pred, B)
    M = S * S * B
    dev = pred.device
    boxes = torch.empty((N, M, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((N, M), dtype=torch.float32, device=dev)
    cls = torch.empty((N, M), dtype=torch.int32, device=dev)
    counts = torch.empty((N,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().yolo1_decode(pred.data_ptr(), _lib.strides4(pred), _pred_dtype(pred), N, S, B, C,
                                     float(thresh), boxes.data_ptr(), scores.data_ptr(), cls.data_ptr(),
                                     counts.data_ptr(), _stream_ptr(dev))
        _lib.check(rc, "yolo1_decode")
    return boxes, scores, cls, counts


def nms_batched(boxes, scores, counts, threshold=0.25, cls=None, per_class=False):
    """nms (utils/utils.py:150-184) for N independent box sets: boxes [N,M,4], scores [N,M], counts [N] int32
    (CUDA).  Returns (keep [N,M] int32 indices in descending score order, keep_counts [N] int32)."""
    if not boxes.is_cuda:
        raise RuntimeError("nms_batched needs CUDA tensors")
    N, M = scores.shape
    dev = boxes.device
    boxes = boxes.contiguous().float()
    scores = scores.contiguous().float()
    counts = counts.to(device=dev, dtype=torch.int32).contiguous()
    if cls is not None:
        cls = cls.to(device=dev, dtype=torch.int32).contiguous()
    keep = torch.empty((N, M), dtype=torch.int32, device=dev)
    keep_counts = torch.empty((N,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().yolo1_nms(boxes.data_ptr(), scores.data_ptr(), cls.data_ptr() if cls is not None else None,
                                  counts.data_ptr(), N, M, float(threshold), int(bool(per_class)),
                                  keep.data_ptr(), keep_counts.data_ptr(), _stream_ptr(dev))
        _lib.check(rc, "yolo1_nms")
    return keep, keep_counts


def nms(bboxes, scores, threshold=0.25):
    """Reference signature (utils/utils.py:150): bboxes [n,4] xyxy, scores [n] -> CPU LongTensor of kept
    indices in descending score order.  Class-agnostic greedy suppression; a box survives iff
    IoU <= threshold (fp32).  Score ties break towards the lower index."""
    n = int(scores.shape[0])
    if n == 0:
        return torch.zeros(0, dtype=torch.long)
    if n > 1024:
        raise ValueError("nms supports at most 1024 boxes per set (S*S*B of the detector), got %d" % n)
    dev = bboxes.device if bboxes.is_cuda else _cuda_device()
    b = bboxes.detach().to(dev, torch.float32).reshape(1, n, 4)
    s = scores.detach().to(dev, torch.float32).reshape(1, n)
    counts = torch.full((1,), n, dtype=torch.int32, device=dev)
    keep, kc = nms_batched(b, s, counts, threshold)
    both = torch.cat([kc, keep[0]]).cpu()     # one D2H copy, one sync
    return both[1:1 + int(both[0])].to(torch.long)


def decoder(pred, grid_num=7, B=2, device='cpu', thresh=0.3, nms_th=0.5, gt=False):
    """Reference signature (utils/utils.py:94): pred [1,S,S,5B+C] (or [S,S,5B+C]) ->
    (boxes [K,4] float32 xyxy, cls_indexs [K] int64, probs [K] float32) on `device`, descending score order.
    Nothing above `thresh` -> the reference's sentinel (zeros[1,4], zeros[1], zeros[1]) (:134-137).
    gt=True runs NMS with threshold 1.0 (:143-145)."""
    p = pred.detach()
    if p.dim() == 4:
        if p.shape[0] != 1:
            raise ValueError("decoder takes one image ([1,S,S,D]); use decode_nms_batched for batches")
        p = p[0]
    if p.shape[0] != grid_num or p.shape[1] != grid_num:
        raise ValueError("pred grid %s does not match grid_num=%d" % (tuple(p.shape[:2]), grid_num))
    dev = p.device if p.is_cuda else _cuda_device(device)
    if not p.is_cuda:
        p = p.to(dev)
    if p.dtype not in (torch.float32, torch.bfloat16):
        p = p.float()
    M = grid_num * grid_num * B
    boxes, cls, probs, counts = decode_nms_batched(p.unsqueeze(0), thresh, 1.0 if gt else nms_th, True, B)
    # one packed D2H copy: [count | cls | probs bits | boxes bits]
    packed = torch.cat([counts, cls[0], probs[0].view(torch.int32), boxes[0].reshape(-1).view(torch.int32)])
    packed = packed.to(device) if torch.device(device).type != "cpu" else packed.cpu()
    k = int(packed[0])
    if k == 0:
        return (torch.zeros((1, 4), device=device), torch.zeros(1, device=device), torch.zeros(1, device=device))
    out_cls = packed[1:1 + k].to(torch.long)
    out_probs = packed[1 + M:1 + M + k].view(torch.float32)
    out_boxes = packed[1 + 2 * M:1 + 2 * M + 4 * k].view(torch.float32).reshape(k, 4)
    return out_boxes, out_cls, out_probs


def compute_iou_matrix(bbox1, bbox2):
    """utils/utils.py:10-57: [N,4] x [M,4] xyxy -> IoU [N,M]; no +1, no epsilon.  Raises TypeError on
    non-tensor input (the reference prints and calls exit(), :30-32)."""
    if not (torch.is_tensor(bbox1) and torch.is_tensor(bbox2)):
        raise TypeError("compute_iou_matrix expects tensors")
    lt = torch.max(bbox1[:, None, :2], bbox2[None, :, :2])
    rb = torch.min(bbox1[:, None, 2:], bbox2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    a1 = (bbox1[:, 2] - bbox1[:, 0]) * (bbox1[:, 3] - bbox1[:, 1])
    a2 = (bbox2[:, 2] - bbox2[:, 0]) * (bbox2[:, 3] - bbox2[:, 1])
    return inter / (a1[:, None] + a2[None, :] - inter)


def convert_CxCyWH_to_X1Y1X2Y2(input_tensor, S, B=2, device='cpu'):
    """utils/utils.py:59-75: [n,4] cell-relative (x,y,w,h) -> (x/S - w/2, y/S - h/2, x/S + w/2, y/S + h/2)."""
    if input_tensor.shape[-1] != 4:
        raise AssertionError("last dimension must be 4")
    xy = input_tensor[..., :2] / float(S)
    half = 0.5 * input_tensor[..., 2:]
    return torch.cat([xy - half, xy + half], dim=-1)
