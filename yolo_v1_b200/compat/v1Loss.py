"""Import shim: put `yolo_v1_b200/compat` on sys.path and the reference's `from v1Loss import YOLOLossV1`
(train.py) binds to the B200 implementation.  See INTEGRATION.md."""
from yolo_v1_b200.loss import YOLOLossV1  # noqa: F401
