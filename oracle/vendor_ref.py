"""Recipe that makes the UNMODIFIED reference modules of the hot path available
where /root/reference does not exist (the GPU box).

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by karios_b200/).

    python -m oracle.vendor_ref            # run by __graft_entry__.build()

The reference is pure Python; "building" it means placing byte-identical copies
of the few files of the path under oracle/_ref/karios/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box like the built .so files, and never
enters the history).  Nothing is edited: MANIFEST.json records the sha256 of
every source and of its copy, and tests/test_oracle.py::test_vendored_reference
re-checks the copies against /root/reference whenever that tree is present.
The third-party imports the files make at module level but the path never
calls (skimage.io, osgeo) are stubbed at import time by oracle/refimport.py
(SURVEY.md Appendix B), not here.

Used by `bench.py --impl reference` and by the `cpu_baseline` leg
(`cpu_baseline.kind = "reference"`): KLT(conf).match(...) and
ZNCCService().compute_zncc(...) of karios/matcher/klt.py:198-349 and
karios/matcher/zncc_service.py:162-184, driven as karios/api/core.py:845-891
drives them.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("KARIOS_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")

# the files of the path and what they import from the package
FILES = [
    "karios/matcher/klt.py",
    "karios/matcher/zncc_service.py",
    "karios/matcher/mutual_info_service.py",
    "karios/matcher/large_offset.py",
    "karios/core/configuration.py",
    "karios/core/errors.py",
    "karios/core/image.py",
    "LICENSE",
    "NOTICE",
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available() -> bool:
    """True when the vendored tree is complete."""
    return os.path.exists(os.path.join(DEST, "MANIFEST.json")) and all(
        os.path.exists(os.path.join(DEST, f)) for f in FILES)


def vendor(force: bool = False) -> str | None:
    """Copy the files (only when the reference tree is present).  -> DEST or None."""
    if not os.path.isdir(os.path.join(REF_SRC, "karios", "matcher")):
        return DEST if available() else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(DEST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = {"sha256": _sha(dst), "bytes": os.path.getsize(dst)}
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "telespazio-tim/karios (Apache-2.0), copied unmodified by oracle/vendor_ref.py",
                   "files": manifest}, f, indent=1, sort_keys=True)
    return DEST


def verify() -> list[str]:
    """Files whose copy differs from the manifest (tampering check).  [] = intact."""
    bad = []
    try:
        man = json.load(open(os.path.join(DEST, "MANIFEST.json")))["files"]
    except Exception:  # noqa: BLE001
        return ["MANIFEST.json"]
    for rel, meta in man.items():
        p = os.path.join(DEST, rel)
        if not os.path.exists(p) or _sha(p) != meta["sha256"]:
            bad.append(rel)
    return bad


if __name__ == "__main__":
    out = vendor(force="--force" in sys.argv)
    print(out if out else "reference tree absent and no vendored copy")
