#!/usr/bin/env python
"""Populate baseline/_ref/ with the UNMODIFIED upstream Python packages the parity tests and the
reference arm of bench.py execute (git-ignored, not gpurun-ignored: it travels to the GPU box with
the snapshot exactly like the built libspq_b200.so; nothing under it is product source).

    python tools/install_ref.py [--src /root/reference] [--force]

The upstream repo is pure Python without packaging metadata (no setup.py / pyproject), so the base
contract's `pip install --target baseline/_ref /root/reference` has nothing to build; this is the
equivalent: a byte-for-byte copy of the three packages that hold the hot path and its callers

    part1_switchable_precision/   quantization, quantization_methods, lora, switchable_batchnorm,
                                  models_sp + the callers train_sp / distillation_manager / deploy
    part2_cyclic_precision_training/  the CPT variant and its scheduler / calibration manager
    part5_squad/ (+ tests/)       byte-identical copies of the hot modules and the only real pytest
                                  suite of the reference (API conformance)

with a MANIFEST.json of sha256 digests so a test can show the files it ran are the upstream ones.
`__graft_entry__.build()` runs this whenever /root/reference is present (the build container); on
the GPU box the prebuilt copy is used as shipped.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("part1_switchable_precision", "part2_cyclic_precision_training", "part5_squad")
SKIP_DIRS = {"__pycache__", "test"}          # part2/test: stale print scripts


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def install(src="/root/reference", force=False, quiet=False):
    """Returns the destination directory, or None when the upstream tree is not available."""
    if not os.path.isdir(src):
        return DEST if os.path.exists(os.path.join(DEST, "MANIFEST.json")) else None
    manifest = {}
    files = []
    for pkg in PACKAGES:
        for dirpath, dirnames, filenames in os.walk(os.path.join(src, pkg)):
            dirnames[:] = [d for d in dirnames if d not in SKIP_DIRS]
            for f in filenames:
                if f.endswith((".py", ".json")):
                    full = os.path.join(dirpath, f)
                    files.append((full, os.path.relpath(full, src)))
    for full, rel in files:
        manifest[rel] = _sha(full)
    mpath = os.path.join(DEST, "MANIFEST.json")
    if not force and os.path.exists(mpath):
        try:
            with open(mpath) as fh:
                if json.load(fh).get("files") == manifest:
                    return DEST
        except Exception:
            pass
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    for full, rel in files:
        out = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(full, out)
    with open(mpath, "w") as fh:
        json.dump({"source": src, "packages": list(PACKAGES), "files": manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print(f"installed {len(files)} upstream files into {DEST}")
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    d = install(a.src, a.force)
    if d is None:
        print("upstream tree not found and no previous install", file=sys.stderr)
        sys.exit(1)
    print(d)
