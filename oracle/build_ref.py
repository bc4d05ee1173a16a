#!/usr/bin/env python
"""Recipe for oracle/_ref/: stage the reference's OWN implementation of the hot path so that it can be
timed and called on the GPU box, where /root/reference does not exist.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference is pure Python: "building" it means placing its two source files, byte for byte, where the
snapshot that travels to the GPU box carries them — the analogue of compiling a C reference into oracle/_ref/.
oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored.
A MANIFEST with the SHA-256 of every staged file is written next to them; ref_loader.load() re-checks it.

    python oracle/build_ref.py          # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = ["src/model.py", "src/retrieval.py"]          # the only files on the hot path


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def build() -> bool:
    """Returns True when oracle/_ref/ holds the reference files (staged now or already there)."""
    if not os.path.isdir(REF_ROOT):
        return os.path.exists(os.path.join(OUT, "MANIFEST.json"))
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT, os.path.basename(rel))
        shutil.copyfile(src, dst)                      # unmodified, byte for byte
        manifest[os.path.basename(rel)] = {"source": rel, "sha256": sha256(dst)}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref staged:", build())
