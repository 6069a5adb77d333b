#!/usr/bin/env python
"""TEST INFRASTRUCTURE: recipe for oracle/_ref -- the reference's own CPU implementation of the hot path, so that
`bench.py --impl reference` and the `cpu_baseline` leg can time THE REFERENCE (kind = "reference") on the GPU box,
where /root/reference does not exist.

    python oracle/build_ref.py        # build container only; __graft_entry__.build() runs it when /root/reference exists

The reference is pure Python with no setup.py / pyproject.toml (`pip install /root/reference` has nothing to
install), so "building" it means taking the modules of the path, byte for byte, from where they lie:

    /root/reference/utils/{homophily_metrics,util_funcs,homophily_plot,datasets}.py  ->  oracle/_ref/utils/

oracle/_ref/ is listed in .gitignore (the reference sources never enter the history) but not in .gpurunignore, so it
travels to the GPU box like the built .so.  Nothing is patched: the stubs for the absent third-party imports live in
oracle/ref_shim.py, which is ours.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["utils/homophily_metrics.py", "utils/util_funcs.py", "utils/homophily_plot.py", "utils/datasets.py"]


def build():
    if not os.path.isdir(os.path.join(SRC, "utils")):
        return None     # GPU box / no reference: use what travelled with the repo
    os.makedirs(os.path.join(DST, "utils"), exist_ok=True)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    return DST


if __name__ == "__main__":
    print(build())
