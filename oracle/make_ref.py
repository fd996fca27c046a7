"""TEST / BENCH INFRASTRUCTURE ONLY -- recipe for ``oracle/_ref``: a VERBATIM copy of the reference's eight source files
and six configs (``/root/reference/{src/*.py,config/*.json}``) so that the unmodified reference can be timed on the GPU
box's host cores (``bench.py --impl reference``, ``cpu_baseline.kind = "reference"``).

``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so it
travels with the snapshot like the built ``libagym.so``.  Nothing is edited: the two compatibility shims the reference
needs on this image (plotting stubs, ``ReduceLROnPlateau(verbose=)``) live in ``oracle/ref_harness.py``, applied at
import time.  ``python -m oracle.make_ref`` prints a sha256 manifest; ``__graft_entry__.build()`` calls it whenever
``/root/reference`` is present.
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def make(src_root="/root/reference", quiet=False):
    if not os.path.isfile(os.path.join(src_root, "src", "Auction.py")):
        return None
    manifest = []
    for sub, pat in (("src", "*.py"), ("config", "*.json")):
        os.makedirs(os.path.join(DEST, sub), exist_ok=True)
        for f in sorted(glob.glob(os.path.join(src_root, sub, pat))):
            dst = os.path.join(DEST, sub, os.path.basename(f))
            shutil.copyfile(f, dst)
            with open(dst, "rb") as fh:
                manifest.append((hashlib.sha256(fh.read()).hexdigest()[:16], os.path.join(sub, os.path.basename(f))))
    with open(os.path.join(DEST, "MANIFEST.txt"), "w") as fh:
        fh.write("verbatim copy of /root/reference (soopark0221/auction-gym), sha256[:16] per file\n")
        fh.writelines(f"{h}  {n}\n" for h, n in manifest)
    if not quiet:
        for h, n in manifest:
            print(h, n)
    return DEST


def ref_src():
    """Path of the reference sources to import: oracle/_ref/src when the copy exists, else /root/reference/src, else None."""
    for p in (os.path.join(DEST, "src"), "/root/reference/src"):
        if os.path.isfile(os.path.join(p, "Auction.py")):
            return p
    return None


if __name__ == "__main__":
    out = make(*(sys.argv[1:2]))
    print("wrote", out) if out else print("no reference tree found; nothing copied")
