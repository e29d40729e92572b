"""Import shim for the reference's `from utils.utils import *` (train.py:14, eval.py:17, run_voc_mAP.py:3,
v1Loss.py, utils/YOLODataLoader.py): the hot-path functions come from the B200 implementation, everything else
stays the reference's own.  See INTEGRATION.md.

`yolo_v1_b200/compat/utils/` deliberately has NO `__init__.py`, and neither has the reference's `utils/`: with
`compat` ahead of the reference directory on sys.path, `utils` becomes ONE namespace package spanning both
directories (PEP 420).  `utils.utils` resolves here (first portion); `utils.YOLODataLoader` and `utils.visual`
(train.py:13,20) resolve in the reference tree.  This module then loads the reference's own `utils/utils.py` from the
other portion under a private name, re-exports every public name of it (`prep_test_data`, `create_logger`,
`cv_resize`, `convert_input_tensor_dim`, `bbox_un_norm`, ... and the modules it star-exports: `os`, `np`, `cv2`,
`torch`), and overrides the hot-path names with the CUDA implementations.  The reference module's OWN globals are
patched too, so its internal callers (`run_test_mAP` -> `decoder` -> `nms`, utils/utils.py:146,405) reach the kernels.
Without a reference tree on the path (tests, standalone use) only the hot-path names are provided.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys
import warnings as _warnings

_HERE = _os.path.dirname(_os.path.abspath(__file__))
_REF_NAME = "utils._reference_utils"


def _load_reference_utils():
    pkg = _sys.modules.get("utils")
    for d in list(getattr(pkg, "__path__", []) or []):
        cand = _os.path.join(d, "utils.py")
        if _os.path.abspath(d) == _HERE or not _os.path.isfile(cand):
            continue
        spec = _ilu.spec_from_file_location(_REF_NAME, cand)
        mod = _ilu.module_from_spec(spec)
        _sys.modules[_REF_NAME] = mod
        try:
            spec.loader.exec_module(mod)
        except ImportError as e:     # e.g. cv2 / tqdm missing: keep the hot path usable, say what is missing
            _sys.modules.pop(_REF_NAME, None)
            _warnings.warn("yolo_v1_b200.compat: the reference's %s could not be imported (%s); only the hot-path "
                           "names are available from utils.utils" % (cand, e))
            return None
        return mod
    return None


_ref = _load_reference_utils()
if _ref is not None:
    for _k, _v in vars(_ref).items():
        if not _k.startswith("_"):
            globals()[_k] = _v

from yolo_v1_b200.decode import (decoder, nms, compute_iou_matrix, convert_CxCyWH_to_X1Y1X2Y2,  # noqa: E402,F401
                                 decode_nms_batched)
from yolo_v1_b200.voc import VOC_CLASSES, voc_ap, voc_eval, run_test_mAP  # noqa: E402,F401

HOT_PATH_NAMES = ("decoder", "nms", "compute_iou_matrix", "convert_CxCyWH_to_X1Y1X2Y2", "voc_ap", "voc_eval",
                  "run_test_mAP")
if _ref is not None:
    for _k in HOT_PATH_NAMES:
        setattr(_ref, _k, globals()[_k])
