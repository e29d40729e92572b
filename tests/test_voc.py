"""VOC evaluation: the host restatement of voc_eval / voc_ap against the reference's
golden values (CPU), and -- on the GPU -- mAP from the CUDA detections equal to the reference's mAP as a Python
float (`==`), which is the north star's "VOC mAP bit-exact" gate."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from yolo_v1_b200 import voc


def _meta(golden_dir):
    return json.load(open(os.path.join(golden_dir, "golden_meta.json")))


def _gt(meta):
    return {(k[0], k[1]): [list(b) for b in k[2]] for k in meta["map_case"]["gt"]}


def test_voc_eval_reference_fixture(golden_dir):
    """utils/utils.py:321-324 `test_eval()` -- golden values produced by the reference itself."""
    preds = {'cat': [['image01', 0.9, 20, 20, 40, 40], ['image01', 0.8, 20, 20, 50, 50], ['image02', 0.8, 30, 30, 50, 50]],
             'dog': [['image01', 0.78, 60, 60, 90, 90]]}
    target = {('image01', 'cat'): [[20, 20, 41, 41]], ('image01', 'dog'): [[60, 60, 91, 91]],
              ('image02', 'cat'): [[30, 30, 51, 51]]}
    lines = []

    class Log:
        def info(self, s):
            lines.append(s)

    m = voc.voc_eval(preds, target, VOC_CLASSES=['cat', 'dog'], logger=Log())
    fx = _meta(golden_dir)["voc_eval_fixture"]
    assert m == fx["mAP"] == 0.9166666666666666
    assert lines == fx["log"]                      # same per-class lines, same formatting


def test_voc_eval_empty_class_breaks_the_loop():
    """A class without detections records -1 and ENDS the evaluation (utils/utils.py:248-255)."""
    preds = {'a': [['i', 0.9, 0, 0, 10, 10]], 'b': [], 'c': [['i', 0.9, 0, 0, 10, 10]]}
    target = {('i', 'a'): [[0, 0, 10, 10]], ('i', 'c'): [[0, 0, 10, 10]]}
    m = voc.voc_eval(preds, target, VOC_CLASSES=['a', 'b', 'c'], logger=type("L", (), {"info": lambda s, x: None})())
    assert m == (1.0 + -1) / 2


def test_voc_ap_07_metric_and_envelope():
    rec = np.array([0.2, 0.2, 0.4, 0.4, 0.6])
    prec = np.array([1.0, 0.5, 0.66, 0.5, 0.6])
    assert voc.voc_ap(rec, prec) == pytest.approx(0.2 * 1.0 + 0.2 * 0.66 + 0.2 * 0.6)
    # np.arange(0, 1.1, 0.1)[6] is 0.6000000000000001 > 0.6: the 0.6 threshold finds no recall point, as upstream
    assert voc.voc_ap(rec, prec, use_07_metric=True) == pytest.approx((3 * 1.0 + 2 * 0.66 + 1 * 0.6) / 11.)


def _preds_from_oracle(pred, S):
    """CPU stand-in for the GPU decode (tests may use the oracle): the same lists run_test_mAP builds."""
    from collections import defaultdict
    preds = defaultdict(list)
    for n in range(pred.shape[0]):
        b, c, s = O.decoder(pred[n], grid_num=S, thresh=0.005, nms_th=0.45)
        if len(s) == 1 and s[0] == 0:
            continue
        pix = (np.clip(b, np.float32(0), np.float32(1)) * np.float32(448)).astype(np.int64)
        for j in range(len(s)):
            preds[voc.VOC_CLASSES[int(c[j])]].append(["img%04d" % n, float(s[j])] + [int(v) for v in pix[j]])
    return preds


def test_map_case_host_logic_equals_reference(golden_dir):
    meta = _meta(golden_dir)
    pred = np.load(os.path.join(golden_dir, "map_case.npz"))["pred"]
    lines = []
    m = voc.voc_eval(_preds_from_oracle(pred, meta["map_case"]["S"]), _gt(meta),
                     logger=type("L", (), {"info": lambda s, x: lines.append(x)})())
    assert m == meta["map_case"]["mAP"]
    aps = [float(l.split(" ap ")[1].rstrip("-")) for l in lines if "---class" in l]
    assert aps == meta["map_case"]["aps"]


@pytest.mark.gpu
def test_map_from_cuda_detections_is_bit_exact(golden_dir):
    """run_test_mAP with an identity network over the golden predictions: decode+NMS, clamp, pixel conversion
    on the GPU; matching on the host; result == the reference's run_test_mAP value."""
    meta = _meta(golden_dir)
    pred = torch.from_numpy(np.load(os.path.join(golden_dir, "map_case.npz"))["pred"])
    dataset = [(pred[n], torch.zeros(1), "/x/img%04d.jpg" % n) for n in range(pred.shape[0])]
    for bs in (1, 7, 96):
        lines = []
        m = voc.run_test_mAP(lambda x: x, _gt(meta), dataset, len(dataset), S=7, device="cuda",
                             logger=type("L", (), {"info": lambda s, x: lines.append(x)})(), batch_size=bs)
        assert m == meta["map_case"]["mAP"], (bs, m)
        aps = [float(l.split(" ap ")[1].rstrip("-")) for l in lines if "---class" in l]
        assert aps == meta["map_case"]["aps"]
    # NCHW predictions (`reversed=True`, eval.py:22-30) are read through the permuted view
    ds2 = [(pred[n].permute(2, 0, 1).contiguous(), torch.zeros(1), "/x/img%04d.jpg" % n) for n in range(pred.shape[0])]
    m = voc.run_test_mAP(lambda x: x, _gt(meta), ds2, len(ds2), S=7, device="cuda", reversed=True,
                         logger=type("L", (), {"info": lambda s, x: None})())
    assert m == meta["map_case"]["mAP"]


@pytest.mark.gpu
def test_pixel_conversion_matches_reference_arithmetic():
    g = torch.Generator().manual_seed(3)
    b = torch.rand(1000, 4, generator=g) * 1.4 - 0.2            # values outside [0,1] get clamped
    got = voc.boxes_to_pixels(b.cuda()).cpu().numpy()
    want = (np.clip(b.numpy(), np.float32(0), np.float32(1)) * np.float32(448)).astype(np.int32)
    assert np.array_equal(got, want)
