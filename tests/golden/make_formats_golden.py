"""Golden outputs of the UNCHANGED reference's file-format helpers (run in the build container only):
  * config_parser.read_config (config_parser.py:49-92) on tests/golden/default_like.cfg (the key/value set of
    files/default.cfg) -> formats_config.json
  * data.save_obj (data.py:20-26) on a small seeded mesh -> formats_save_obj.obj
      python tests/golden/make_formats_golden.py
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MVB_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))


def main():
    import config_parser                                       # the reference's own file
    cfg = config_parser.read_config(os.path.join(HERE, "default_like.cfg"))
    with open(os.path.join(HERE, "formats_config.json"), "w") as fp:
        json.dump(cfg, fp, indent=1, sort_keys=True)
    # data.py imports torchvision / open3d / psbody at module level; only save_obj is needed: exec that one function
    src = open(os.path.join(REF, "data.py")).read()
    start = src.index("def save_obj")
    end = src.index("def OnUnitCube")
    ns = {}
    exec(compile(src[start:end], "data.py:save_obj", "exec"), ns)
    rng = np.random.default_rng(5)
    v = rng.normal(size=(7, 3)) * 100.0
    f = rng.integers(0, 7, size=(5, 3))
    ns["save_obj"](os.path.join(HERE, "formats_save_obj.obj"), v, f)
    np.savez(os.path.join(HERE, "formats_save_obj_input.npz"), v=v, f=f)
    print(cfg)


if __name__ == "__main__":
    main()
