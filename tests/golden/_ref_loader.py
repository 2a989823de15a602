"""Import the UNMODIFIED reference hot-path modules from /root/reference (container only).

The reference cannot be imported as shipped under Python 3.12 / numpy 2.3:
  * it imports matplotlib / pylab, which are not installed        -> stub modules
  * it uses the removed alias ``np.int``                           -> restored for the session
  * utils/cube_to_equi.py:47,49 uses ``.cuda(async=True)``        -> SyntaxError on py>=3.7;
    the text is patched in memory (never written to disk) to run on a CPU-only box.

Nothing here is copied into the repo: the sources are read where they lie and exec'd.
This module is only used by ``make_golden.py`` (fixture generation) and by the optional
``tests/test_oracle_vs_reference.py`` which skips when /root/reference is absent (GPU box).
"""
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("CP360_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "cube_pad.py"))


def _install_stubs():
    if not hasattr(np, "int"):
        np.int = int  # cube_pad.py:13,64 ; cube_to_equi.py:49
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    if "pylab" not in sys.modules:
        pl = types.ModuleType("pylab")
        for k in dir(np):
            if not k.startswith("_"):
                setattr(pl, k, getattr(np, k))
        pl.__all__ = [k for k in dir(np) if not k.startswith("_")]
        sys.modules["pylab"] = pl
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def load():
    """Returns (cube_pad_module, equi_to_cube_module, cube_to_equi_module_cpu)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    _install_stubs()
    cube_pad = importlib.import_module("model.cube_pad")
    e2c = importlib.import_module("utils.equi_to_cube")
    # cube_to_equi: in-memory text patch so that it parses and runs without CUDA.
    path = os.path.join(REF_ROOT, "utils", "cube_to_equi.py")
    src = open(path).read()
    src = src.replace(".cuda(async=True)", "")
    src = src.replace(", requires_grad=True).cuda()", ")")
    mod = types.ModuleType("utils.cube_to_equi_cpu")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return cube_pad, e2c, mod
