"""Load the REAL reference (haoran1062/YOLO_V1) from /root/reference for golden-vector generation.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py in the build container, where the
read-only reference mount exists, and by bench.py's CPU-reference legs on the GPU box, where only the two
files staged into oracle/_ref/ exist (oracle/stage_reference.py).  Nothing under yolo_v1_b200/ imports it.

The loss (`v1Loss.py:9-118`) is imported unmodified and built with `_device='cpu'`.
`utils/utils.py` needs a one-token compatibility shim to run on torch >= 0.5: at
`utils/utils.py:180` the expression `(ovr<=threshold).nonzero().squeeze()` yields a 0-dim tensor
when exactly one box survives a round, and `order[0]` (`:164`) then raises IndexError.
Replacing `.squeeze()` by `.reshape(-1)` restores the torch-0.4 behaviour the code was written for
(SURVEY.md section 8(c)).  Everything else is the reference's own code, executed as is.
"""
import contextlib
import io
import os
import sys
import types
import warnings

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_hot_path.zip")


def _ref_root():
    """The reference tree: $YOLO1_REFERENCE_ROOT, else the read-only mount of the build container, else the two
    hot-path files packed into oracle/_ref/reference_hot_path.zip at build() time by oracle/stage_reference.py
    (the only form that exists on the GPU box; Python imports from the archive directly)."""
    env = os.environ.get("YOLO1_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/v1Loss.py"):
        return "/root/reference"
    return _STAGED


REF_ROOT = _ref_root()
_SHIM_OLD = "(ovr<=threshold).nonzero().squeeze()"
_SHIM_NEW = "(ovr<=threshold).nonzero().reshape(-1)"


def _is_zip():
    return os.path.isfile(REF_ROOT) and REF_ROOT.endswith(".zip")


def available():
    return _is_zip() or os.path.isfile(os.path.join(REF_ROOT, "v1Loss.py"))


def _read_text(rel):
    if _is_zip():
        import zipfile
        with zipfile.ZipFile(REF_ROOT) as z:
            return z.read(rel).decode("utf-8")
    with open(os.path.join(REF_ROOT, rel), "r", encoding="utf-8") as f:
        return f.read()


def load_reference():
    """Returns (YOLOLossV1, utils_module) where utils_module carries the shimmed decoder/nms."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    path = os.path.join(REF_ROOT, "utils", "utils.py")
    src = _read_text("utils/utils.py")
    assert src.count(_SHIM_OLD) == 1, "reference nms source changed; shim no longer applies"
    src = src.replace(_SHIM_OLD, _SHIM_NEW)
    # the reference imports `utils.utils`; register the shimmed text under that name so that
    # v1Loss.py (`from utils.utils import *`) binds to the very same functions.
    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.join(REF_ROOT, "utils")]
    mod = types.ModuleType("utils.utils")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.utils", "v1Loss")}
    sys.modules["utils"] = pkg
    sys.modules["utils.utils"] = mod
    pkg.utils = mod
    sys.path.insert(0, REF_ROOT)
    try:
        sys.modules.pop("v1Loss", None)
        import v1Loss as ref_loss_mod  # noqa
    finally:
        sys.path.remove(REF_ROOT)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ref_loss_mod.YOLOLossV1, mod


@contextlib.contextmanager
def quiet():
    """The reference prints its four loss components on every call (`v1Loss.py:110`)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with contextlib.redirect_stdout(io.StringIO()):
            yield
