"""TEST / BASELINE INFRASTRUCTURE - not part of the product.

The reference is pure Python, so "building" it for the GPU box means staging its UNCHANGED files where the box can
import them: /root/reference exists only in the build container.  This script copies the files of the hot path and of
its callers into oracle/_ref/reference/ (git-ignored: never part of the history; not gpurun-ignored: it travels with the
snapshot like a built .so).  Nothing under oracle/_ref is edited.  Users: tests/ (the unchanged reference models on the
compat/ drop-in on a B200, the f3 / A7 pins), bench.py --impl reference and its cpu_baseline leg.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MVB_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
FILES = ["nn/conv.py", "nn/pool.py", "models/cheb_VAE.py", "models/cheb_cls.py", "logpdf.py", "utils.py", "main.py",
         "data.py", "transform.py", "config_parser.py", "mesh_operations.py", "model.py", "inference.py", "crecon.py",
         "files/default.cfg", "files/crecon.cfg", "template/template5k.obj"]


def make(verbose: bool = True) -> str:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"make_ref: {SRC} not present - keeping {DST if os.path.isdir(DST) else 'nothing'}")
        return DST if os.path.isdir(DST) else ""
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    if verbose:
        print(f"make_ref: staged {len(FILES)} unchanged reference files under {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if make() or not os.path.isdir(SRC) else 1)
