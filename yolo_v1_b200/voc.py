"""PASCAL-VOC evaluation around the decode/NMS kernels.

Restates the reference's evaluation driver and metric so that mAP computed from the CUDA detections equals
the reference's value exactly:
  * `run_test_mAP` post-processing (/root/reference/utils/utils.py:389-418): per image decode with
    thresh=0.005 / nms 0.45, clamp to [0,1], pixel = trunc(fl32(v * 448)), per-class detection lists in
    image order then score order, the all-zero "no detection" sentinel skipped;
  * `voc_eval` / `voc_ap` (utils/utils.py:215-319): greedy matching in descending confidence, IoU with the
    +1 pixel convention, a ground-truth box is consumed by its first match, area under the monotone
    precision envelope; a class without detections records -1 and ENDS the loop (`break`, :248-255); the
    -1 enters the mean.
The detections come from the GPU (one batched decode+NMS launch and one pixel-conversion launch per network
batch instead of one Python loop per image); the matching itself is sequential by nature and stays on the
host, vectorised per detection with NumPy.
"""
import ctypes
from collections import defaultdict

import numpy as np
import torch

from . import _lib
from .decode import decode_nms_batched

VOC_CLASSES = ('aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow',
               'diningtable', 'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train',
               'tvmonitor')

__all__ = ["VOC_CLASSES", "voc_ap", "voc_eval", "boxes_to_pixels", "detections_to_voc_preds", "run_test_mAP"]


def voc_ap(rec, prec, use_07_metric=False):
    """utils/utils.py:215-238.  Average precision from cumulative recall / precision arrays."""
    rec = np.asarray(rec, dtype=np.float64)
    prec = np.asarray(prec, dtype=np.float64)
    if use_07_metric:                                   # 11-point interpolation (:216-224)
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            sel = rec >= t
            ap = ap + (np.max(prec[sel]) if np.sum(sel) != 0 else 0) / 11.
        return ap
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]      # monotone envelope from the right (:231-232)
    steps = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[steps + 1] - mrec[steps]) * mpre[steps + 1])


def voc_eval(preds, target, VOC_CLASSES=VOC_CLASSES, threshold=0.5, use_07_metric=False, logger=None):
    """utils/utils.py:240-319.
    preds  {class: [[image_id, confidence, x1, y1, x2, y2], ...]}
    target {(image_id, class): [[x1, y1, x2, y2], ...]}   (consumed: matched boxes are removed, as upstream)
    Returns mAP (Python float)."""
    say = logger.info if logger else print
    aps = []
    for class_ in VOC_CLASSES:
        dets = preds[class_]
        if len(dets) == 0:                 # upstream records -1 and leaves the loop (:248-255)
            say('---class {} ap {}---'.format(class_, -1))
            aps.append(-1)
            break
        conf = np.array([float(d[1]) for d in dets])
        boxes = np.array([d[2:] for d in dets])
        order = np.argsort(-conf)          # same call as upstream (:261), so ties fall the same way
        npos = 0.
        for (_, cls_name), gts in target.items():
            if cls_name == class_:
                npos += len(gts)
        hit = np.zeros(len(dets))
        for rank, k in enumerate(order):
            key = (dets[k][0], class_)
            gts = target.get(key)
            if not gts:
                continue
            g = np.asarray(gts, dtype=np.float64).reshape(-1, 4)
            bb = boxes[k].astype(np.float64)
            iw = np.maximum(np.minimum(g[:, 2], bb[2]) - np.maximum(g[:, 0], bb[0]) + 1., 0.)
            ih = np.maximum(np.minimum(g[:, 3], bb[3]) - np.maximum(g[:, 1], bb[1]) + 1., 0.)
            inter = iw * ih
            union = (bb[2] - bb[0] + 1.) * (bb[3] - bb[1] + 1.) + (g[:, 2] - g[:, 0] + 1.) * (g[:, 3] - g[:, 1] + 1.) - inter
            with np.errstate(divide='ignore', invalid='ignore'):
                ok = np.nonzero(inter / union > threshold)[0]
            if ok.size:                    # the first ground-truth box in list order is consumed (:292-298)
                hit[rank] = 1
                del gts[int(ok[0])]
                if len(gts) == 0:
                    del target[key]
        tp = np.cumsum(hit)
        fp = np.cumsum(1 - hit)
        with np.errstate(divide='ignore', invalid='ignore'):
            rec = tp / float(npos)
        prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
        ap = voc_ap(rec, prec, use_07_metric)
        say('---class {} ap {}---'.format(class_, ap))
        aps.append(ap)
    mAP = np.mean(aps).item()
    say('---map {}---'.format(mAP))
    return mAP


def boxes_to_pixels(boxes, img_size=(448, 448)):
    """utils/utils.py:406-407 + :347-354 on the GPU: int32 [...,4] = trunc(fl32(clamp(box,0,1) * (w,h,w,h)))."""
    if not boxes.is_cuda:
        raise RuntimeError("boxes_to_pixels needs a CUDA tensor")
    b = boxes.contiguous().float()
    out = torch.empty(b.shape, dtype=torch.int32, device=b.device)
    n = b.numel() // 4
    with torch.cuda.device(b.device):
        rc = _lib.lib().yolo1_boxes_to_pixels(b.data_ptr(), n, float(img_size[0]), float(img_size[1]), out.data_ptr(),
                                              ctypes.c_void_p(torch.cuda.current_stream(b.device).cuda_stream))
        _lib.check(rc, "yolo1_boxes_to_pixels")
    return out


def detections_to_voc_preds(boxes, cls, probs, counts, image_ids, preds=None, img_size=(448, 448),
                            class_names=VOC_CLASSES):
    """Append a decoded batch to the per-class lists `voc_eval` consumes (utils/utils.py:408-411):
    image order, then descending score.  One D2H copy per batch."""
    preds = defaultdict(list) if preds is None else preds
    pix = boxes_to_pixels(boxes, img_size).cpu().numpy()
    cls_h, probs_h, counts_h = cls.cpu().numpy(), probs.cpu().numpy(), counts.cpu().numpy()
    for n, img_id in enumerate(image_ids):
        for j in range(int(counts_h[n])):      # count 0 == the reference's skipped sentinel (:408-409)
            p = pix[n, j]
            preds[class_names[int(cls_h[n, j])]].append(
                [img_id, float(probs_h[n, j]), int(p[0]), int(p[1]), int(p[2]), int(p[3])])
    return preds


def run_test_mAP(YOLONet, target, test_datasets, data_len, S=7, device='cuda:0', reversed=False, logger=None,
                 little_test=None, *, batch_size=64):
    """Signature of utils/utils.py:389.  `test_datasets` yields (image, target, file_name) per image; images are
    pushed through the network `batch_size` at a time, decoded in one launch per batch.  `reversed=True`
    (an NCHW prediction, eval.py:22-30) is handled by reading the permuted view in place."""
    preds = defaultdict(list)

    def flush(imgs, ids):
        if not imgs:
            return
        with torch.no_grad():
            pred = YOLONet(torch.stack(imgs).to(device))
        if reversed:
            pred = pred.permute(0, 2, 3, 1)
        if pred.dtype not in (torch.float32, torch.bfloat16):
            pred = pred.float()
        boxes, cls, probs, counts = decode_nms_batched(pred, 0.005, .45)       # thresholds of :405
        detections_to_voc_preds(boxes, cls, probs, counts, ids, preds)

    imgs, ids = [], []
    for i, (image, _now_target, fname) in enumerate(test_datasets):
        if little_test and i >= little_test:
            break
        imgs.append(image)
        ids.append(fname.split('/')[-1].split('.')[0])
        if len(imgs) == batch_size:
            flush(imgs, ids)
            imgs, ids = [], []
    flush(imgs, ids)
    (logger.info if logger else print)('---start evaluate---')
    return voc_eval(preds, target, VOC_CLASSES=VOC_CLASSES, threshold=0.5, use_07_metric=False, logger=logger)
