"""Stage the unmodified reference where it can travel to the GPU box.

TEST INFRASTRUCTURE, build container only.  The reference (ipl-uw/ZeDO-Release) is pure Python; the
GPU box has no ``/root/reference``.  This script copies the reference's Python packages ``lib/``,
``run/`` and ``configs/`` byte for byte from the read-only tree into the git-ignored directory
``oracle/_ref/`` (listed in ``.gitignore``, NOT in ``.gpurunignore``, so ``gpurun`` ships it) and writes
``oracle/_ref/MANIFEST.json`` (sha256 per file) so a run can state exactly which sources it executed.
Nothing under ``oracle/_ref/`` is ever committed, imported by the product, or edited.

    python oracle/fetch_ref.py            # (re)stage
    python oracle/fetch_ref.py --check    # verify the staged copy against the manifest

Consumers: ``oracle/ref_runner.py`` (the reference's own PyTorch path as the parity oracle on the same
device and as the CPU / eager-GPU baseline of ``bench.py --impl reference``) and the drop-in tests
that execute ``run/opt_main.py`` unmodified.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("ZEDO_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
PACKAGES = ("lib", "run", "configs")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(src: str = REF_SRC, dst: str = REF_DST) -> dict:
    if not os.path.isdir(src):
        raise SystemExit(f"reference tree {src} not found: fetch_ref.py only runs in the build container")
    manifest = {}
    for pkg in PACKAGES:
        for dirpath, _, files in os.walk(os.path.join(src, pkg)):
            for fn in sorted(files):
                if not fn.endswith(".py"):
                    continue
                s = os.path.join(dirpath, fn)
                rel = os.path.relpath(s, src)
                d = os.path.join(dst, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                manifest[rel] = _sha(d)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": "ipl-uw/ZeDO-Release (unmodified copy of lib/, run/, configs/)", "files": manifest}, f,
                  indent=1, sort_keys=True)
    return manifest


def check(dst: str = REF_DST) -> bool:
    mpath = os.path.join(dst, "MANIFEST.json")
    if not os.path.exists(mpath):
        return False
    with open(mpath) as f:
        manifest = json.load(f)["files"]
    return all(os.path.exists(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == h
               for rel, h in manifest.items())


def available(dst: str = REF_DST) -> bool:
    return os.path.exists(os.path.join(dst, "MANIFEST.json")) and os.path.isdir(os.path.join(dst, "lib"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    if a.check:
        ok = check()
        print("oracle/_ref matches its manifest" if ok else "oracle/_ref missing or modified")
        sys.exit(0 if ok else 1)
    m = stage()
    print(f"staged {len(m)} reference files into {REF_DST}")
