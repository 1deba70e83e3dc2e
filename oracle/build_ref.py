"""Stage the reference's own Python package for the benchmark's reference arm (TEST INFRASTRUCTURE).

    python oracle/build_ref.py          (run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python: there is nothing to compile.  Its hot path (graphphysics/models/{layers,processors,
simulator}.py, graphphysics/utils/{loss,scheduler,nodetype}.py and the two modules those import,
models/transolver.py and utils/vectorial_operators.py) is copied UNMODIFIED from /root/reference into
oracle/_ref/graphphysics/ -- a git-ignored directory (never part of the repository's history) that travels to
the GPU box like a built .so, because /root/reference does not exist there.  `bench.py --impl reference` and the
`cpu_baseline` leg import it through oracle/ref_shim.py (stand-ins for torch_geometric / dgl, which cannot be
installed) and time the reference's own modules on the box's host cores.  A MANIFEST with the sha256 of every
staged file is written next to them; oracle/ref_trainer.py refuses to run if a staged file differs from it."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ["graphphysics/models/layers.py", "graphphysics/models/processors.py", "graphphysics/models/simulator.py",
         "graphphysics/models/transolver.py", "graphphysics/utils/loss.py", "graphphysics/utils/scheduler.py",
         "graphphysics/utils/nodetype.py", "graphphysics/utils/vectorial_operators.py"]
PKG_INITS = ["graphphysics/__init__.py", "graphphysics/models/__init__.py", "graphphysics/utils/__init__.py"]


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"build_ref: {SRC} not present; keeping the staged copy in {DST}" if os.path.isdir(DST)
              else f"build_ref: {SRC} not present and nothing staged")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for rel in FILES + PKG_INITS:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        else:                                   # namespace-style package in the reference: an empty __init__ is enough
            open(dst, "w").close()
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    print(f"build_ref: staged {len(FILES)} reference modules into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
