#!/usr/bin/env python
"""Stages the UNMODIFIED reference model packages into baseline/_ref/ so that they travel to the GPU box.

    python tools/stage_reference.py            (build container only: needs /root/reference)

baseline/_ref/ is git-ignored (never enters history) but NOT gpurun-ignored, exactly like oracle/_ref/: the snapshot that goes to
the GPU box carries it. Only what `from time_interval_machine.models.tim import TIM` needs is copied, byte for byte, per variant
(both variants use the same package name, so each lives under its own root and is imported in its own process):
    <variant>/time_interval_machine/__init__.py, models/** (tim.py, build.py, helpers/*), utils/** (utils/__init__.py imports all of it)
Users: tests/test_real_reference_gpu.py (patch_model on the real nn.Module, on a B200, against the same module run un-patched in
fp32) and bench.py's gpu_eager_baseline leg (the reference's own eager-PyTorch forward on the same GPU). Nothing under tim_b200/
imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def stage() -> str:
    if not os.path.isdir(SRC):
        return ""
    manifest = {}
    for variant in ("recognition", "detection"):
        pkg_src = os.path.join(SRC, variant, "time_interval_machine")
        pkg_dst = os.path.join(DST, variant, "time_interval_machine")
        if os.path.isdir(pkg_dst):
            shutil.rmtree(pkg_dst)
        os.makedirs(pkg_dst, exist_ok=True)
        shutil.copy2(os.path.join(pkg_src, "__init__.py"), os.path.join(pkg_dst, "__init__.py"))
        for sub in ("models", "utils"):       # utils/__init__.py imports every utils module, so the whole directory travels
            shutil.copytree(os.path.join(pkg_src, sub), os.path.join(pkg_dst, sub), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for dp, _, fs in os.walk(pkg_dst):
            for f in fs:
                p = os.path.join(dp, f)
                manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()[:16]
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    d = stage()
    print(d or "reference tree absent, nothing staged")
    sys.exit(0)
