"""yolo_v1_b200 -- B200-native (sm_100a) YOLO v1 hot path: fused loss forward+backward, decode, NMS.

Drop-in for the reference's call surface (haoran1062/YOLO_V1):
    from yolo_v1_b200 import YOLOLossV1, decoder, nms
The arithmetic lives in libyolo1_b200.so (C ABI: include/yolo1_b200.h, sources: yolo_v1_b200/csrc/); there is
no CPU fallback -- entry points raise if the library has not been built.
"""
from .loss import YOLOLossV1, yolo_loss_fused, yolo_loss_from_objects, scale_grad_   # noqa: F401
from .decode import (decoder, nms, decode_nms_batched, decode_batched, nms_batched,   # noqa: F401
                     compute_iou_matrix, convert_CxCyWH_to_X1Y1X2Y2)
from .voc import voc_eval, voc_ap, run_test_mAP, detections_to_voc_preds, boxes_to_pixels, VOC_CLASSES  # noqa: F401
from .encode import encode_targets, encoder, pack_objects             # noqa: F401
from .host import HostContext                                        # noqa: F401
from .dist import shard_range, all_reduce_terms, sharded_loss        # noqa: F401
from .graph import GraphedLoss                                       # noqa: F401

__version__ = "0.1.0"
