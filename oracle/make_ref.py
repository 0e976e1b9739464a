"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that places the UNMODIFIED reference modules of the hot path under
``oracle/_ref/`` (git-ignored, NOT gpurun-ignored), so that they travel to the GPU box with the snapshot.

``/root/reference`` exists only in the build container.  The reference path is Python: there is nothing to compile, the
"build" is a verbatim copy of the files below, with their relative paths kept, plus a manifest with their SHA-256.  The
files are never committed (``oracle/_ref/`` is in .gitignore) and nothing under ``lidar_vision_vqa_b200/`` imports them:
``oracle/ref_loader.py`` loads them for the checker (tests) and for the baseline legs of ``bench.py`` (the reference's own
PillarVFE / PointPillarScatter timed on the host cores and, eager, on the B200).

    python -m oracle.make_ref            # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("LVVQA_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")

FILES = [
    "src/lidar-encoder/pcdet/models/backbones_3d/vfe/vfe_template.py",
    "src/lidar-encoder/pcdet/models/backbones_3d/vfe/pillar_vfe.py",
    "src/lidar-encoder/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py",
    "src/lidar-encoder/pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py",
    "src/lidar-encoder/pcdet/models/backbones_2d/base_bev_backbone.py",
    "src/encoder-decoder/training/models/vat_lidar.py",
    "src/encoder-decoder/training/models/vat_blocks.py",
]


def make_ref(verbose: bool = False) -> bool:
    """Copies the files; returns False (and leaves an existing copy alone) when the reference tree is absent."""
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, FILES[1])):
        return False
    manifest = {}
    for rel in FILES:
        src = os.path.join(REFERENCE_ROOT, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
        if verbose:
            print("copied", rel)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REFERENCE_ROOT, "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = make_ref(verbose=True)
    print("oracle/_ref", "written" if ok else "NOT written: reference tree absent")
    sys.exit(0)
