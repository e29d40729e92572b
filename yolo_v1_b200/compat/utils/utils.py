"""Import shim for the reference's `from utils.utils import *` (train.py, eval.py, run_voc_mAP.py): the hot-path
functions come from the B200 implementation.  Only the names on the hot path and its helpers are provided;
the reference's drawing / logging / dataset helpers stay in the reference tree.  See INTEGRATION.md."""
from yolo_v1_b200.decode import (decoder, nms, compute_iou_matrix, convert_CxCyWH_to_X1Y1X2Y2,  # noqa: F401
                                 decode_nms_batched)

from yolo_v1_b200.voc import VOC_CLASSES, voc_ap, voc_eval, run_test_mAP  # noqa: F401
