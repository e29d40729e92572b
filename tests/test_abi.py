"""CPU-only: the C-ABI library loads and exports every symbol include/yolo1_b200.h declares, the ctypes
binding lists exactly those symbols, and argument validation (which happens before any CUDA call) returns
the documented codes.  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "yolo1_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"YOLO1_API\s+[\w\s\*]+?\b(yolo1_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from yolo_v1_b200 import _lib
    if not os.path.isfile(_lib.SO_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("yolo1_loss_fwd_bwd", "yolo1_scale_grad", "yolo1_decode", "yolo1_nms", "yolo1_decode_nms",
                 "yolo1_loss_fwd_bwd_host", "yolo1_decode_nms_host", "yolo1_loss_workspace_bytes"):
        assert must in names
    assert len(names) >= 15


def test_library_exports_every_declared_symbol(lib):
    L = lib.lib()
    for name in _declared():
        assert hasattr(L, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.SO_PATH], text=True)
    exported = sorted(set(re.findall(r" T (yolo1_\w+)", out)))
    assert exported == _declared()          # nothing undeclared leaks out either
    assert sorted(lib.SIGNATURES) == _declared()


def test_library_is_self_contained_sm100a(lib):
    """No torch / pybind dependency in the library; device code is sm_100a."""
    out = subprocess.check_output(["ldd", lib.SO_PATH], text=True)
    assert "torch" not in out and "python" not in out
    elf = subprocess.run(["cuobjdump", "-lelf", lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf


def test_argument_validation_without_cuda(lib):
    L = lib.lib()
    assert L.yolo1_abi_version() == 1
    assert L.yolo1_loss_workspace_bytes(32, 7, 2, 20) >= 4096
    st = (ctypes.c_int64 * 4)(1470, 210, 30, 1)
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    # null pred
    assert L.yolo1_loss_fwd_bwd(None, st, 0, p, st, None, None, p, 1, 7, 2, 20, 5.0, 0.5, 1.0, 0, p, 1 << 20, None) == -1
    # bad dtype / coord mode
    assert L.yolo1_loss_fwd_bwd(p, st, 7, p, st, None, None, p, 1, 7, 2, 20, 5.0, 0.5, 1.0, 0, p, 1 << 20, None) == -1
    assert L.yolo1_loss_fwd_bwd(p, st, 0, p, st, None, None, p, 1, 7, 2, 20, 5.0, 0.5, 1.0, 9, p, 1 << 20, None) == -1
    # B beyond what the kernels are built for
    assert L.yolo1_loss_fwd_bwd(p, st, 0, p, st, None, None, p, 1, 7, 9, 20, 5.0, 0.5, 1.0, 0, p, 1 << 20, None) == -2
    # workspace too small
    assert L.yolo1_loss_fwd_bwd(p, st, 0, p, st, None, None, p, 1, 7, 2, 20, 5.0, 0.5, 1.0, 0, p, 16, None) == -1
    # misaligned pointer
    assert L.yolo1_loss_fwd_bwd(p + 2, st, 0, p, st, None, None, p, 1, 7, 2, 20, 5.0, 0.5, 1.0, 0, p, 1 << 20, None) == -3
    # decode: too many candidates per image (S*S*B > 1024), null outputs
    assert L.yolo1_decode_nms(p, st, 0, 1, 32, 2, 20, 0.1, 0.5, 0, p, p, p, p, None, None, None) == -2
    assert L.yolo1_decode_nms(p, st, 0, 1, 7, 2, 20, 0.1, 0.5, 0, None, p, p, p, None, None, None) == -1
    assert L.yolo1_nms(p, p, None, p, 1, 2048, 0.5, 0, p, p, None) == -2
    assert L.yolo1_nms(p, p, None, p, 1, 98, 0.5, 1, p, p, None) == -1      # per_class without cls
    assert L.yolo1_scale_grad(None, 0, 4, p, None) == -1
    assert L.yolo1_host_ctx_create(None, 0, 7, 2, 20, 0) == -1
    assert b"invalid argument" in L.yolo1_error_string(-1)
    assert b"aligned" in L.yolo1_error_string(-3)


def test_python_surface_fails_loudly_without_gpu_or_library(lib, monkeypatch):
    import torch
    import yolo_v1_b200 as y
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            y.yolo_loss_fused(torch.zeros(1, 7, 7, 30), torch.zeros(1, 7, 7, 30), batch_size=1)
        with pytest.raises(RuntimeError):
            y.nms(torch.zeros(2, 4), torch.zeros(2))
    # a missing shared object is an error, never a fallback
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "SO_PATH", "/nonexistent/libyolo1_b200.so")
    with pytest.raises(lib.Yolo1LibraryError):
        lib.lib()


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under yolo_v1_b200/ (Python or CUDA) may import, load or name it,
    and there is no CPU fallback to route through."""
    import pathlib
    import re
    root = pathlib.Path(__file__).resolve().parents[1] / "yolo_v1_b200"
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/\\.](_ref|_build|oracle|ref_loader|yolo1_oracle)|libyolo1_oracle")
    hits = []
    for f in list(root.rglob("*.py")) + list(root.rglob("*.cu")) + list(root.rglob("*.cuh")) + list(root.rglob("Makefile")):
        for i, line in enumerate(f.read_text(errors="replace").splitlines(), 1):
            if pat.search(line):
                hits.append("%s:%d: %s" % (f.relative_to(root), i, line.strip()))
    assert not hits, "\n".join(hits)
