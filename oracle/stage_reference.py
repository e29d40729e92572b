"""Stage the reference's own hot-path modules into oracle/_ref/ so that they travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/yolo1_oracle.c header): nothing under yolo_v1_b200/ may import it.

The reference (haoran1062/YOLO_V1) is pure Python: there is nothing to compile.  `__graft_entry__.build()` calls
`stage()` in the build container, where the read-only mount /root/reference exists; the two files the hot path lives
in -- `v1Loss.py` (YOLOLossV1.forward, :22-118) and `utils/utils.py` (decoder :94-147, nms :150-184) -- are packed
BYTE FOR BYTE into one archive, oracle/_ref/reference_hot_path.zip (git-ignored: the reference's sources never enter
this repository's history; not gpurun-ignored: the archive ships to the GPU box like the built .so files).  `bench.py --impl reference` and the
cpu_baseline leg then time the UNMODIFIED reference on the GPU box's host cores (kind: "reference") next to the C
port, through oracle/ref_loader.py (which applies the one-token `.squeeze()` -> `.reshape(-1)` shim in memory, SURVEY
section 8(c); the staged file itself is unmodified, as MANIFEST.json's sha256 shows).
"""
import hashlib
import json
import os
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("YOLO1_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("v1Loss.py", os.path.join("utils", "utils.py"))


ARCHIVE = os.path.join(DST, "reference_hot_path.zip")


def _sha_bytes(b):
    return hashlib.sha256(b).hexdigest()


def stage(verbose=False):
    """Pack the files when the reference mount is present; returns True when the archive exists afterwards."""
    if os.path.isfile(os.path.join(REF_SRC, FILES[0])):
        os.makedirs(DST, exist_ok=True)
        blobs = {rel.replace(os.sep, "/"): open(os.path.join(REF_SRC, rel), "rb").read() for rel in FILES}
        manifest = {rel: {"sha256": _sha_bytes(b), "bytes": len(b)} for rel, b in blobs.items()}
        current = None
        if os.path.isfile(ARCHIVE):
            try:
                with zipfile.ZipFile(ARCHIVE) as z:
                    current = {n: _sha_bytes(z.read(n)) for n in z.namelist()}
            except zipfile.BadZipFile:
                current = None
        if current != {rel: m["sha256"] for rel, m in manifest.items()}:
            with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
                for rel, b in blobs.items():
                    z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), b)   # reproducible bytes
        with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
            json.dump({"source": "haoran1062/YOLO_V1 (read-only mount %s)" % REF_SRC, "files": manifest,
                       "note": "byte-for-byte members of reference_hot_path.zip; git-ignored; only "
                               "oracle/ref_loader.py reads them"}, f, indent=1)
        if verbose:
            print("staged %d reference files into %s" % (len(FILES), ARCHIVE))
    return staged()


def staged():
    return os.path.isfile(ARCHIVE)


if __name__ == "__main__":
    print("staged" if stage(verbose=True) else "reference mount absent and oracle/_ref incomplete")
