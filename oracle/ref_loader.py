"""Run the UNMODIFIED reference (TEST INFRASTRUCTURE / BASELINE ONLY — see oracle/__init__.py).

The reference is pure Python. It is read where it lies: ``/root/reference`` in the build container,
or ``oracle/_ref/`` — a byte-for-byte staging copy of the hot-path files made by ``stage()`` (called
from ``__graft_entry__.build()``; git-ignored, but it travels to the GPU box with the repo snapshot
like the built ``.so``). ``MANIFEST.json`` in the staging directory holds the sha256 of every file as
it was in /root/reference; ``verify()`` re-checks them, so what runs on the GPU box is the reference's
own source, not an edited copy. Nothing from the reference is committed to this repository.

The reference cannot be imported as shipped under Python 3.12 / numpy 2.3:
  * it imports matplotlib / pylab, which are not installed        -> stub modules (no behaviour)
  * it uses the removed alias ``np.int``                           -> restored for the session
  * ``.cuda(async=True)`` (utils/cube_to_equi.py:47,49, class_activation_model.py:58) is a
    SyntaxError on py>=3.7 -> the text is patched IN MEMORY (never on disk) to
    ``.cuda(non_blocking=True)`` (device="cuda") or dropped (device="cpu", together with the
    trailing ``.cuda()`` of :56, so the CPU path runs on a box without a GPU).

Used by tests/, by bench.py's ``--impl reference`` / ``cpu_baseline`` / ``gpu_aten_baseline`` legs
and by tests/golden/make_golden*.py. The product package never imports this module.
"""
import hashlib
import importlib
import json
import os
import shutil
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "oracle", "_ref")
SOURCE = os.environ.get("CP360_REFERENCE", "/root/reference")

# the hot path and the call sites that own it (a-7)
FILES = ["model/cube_pad.py", "model/resnet_cubic.py", "model/clstm.py",
         "utils/__init__.py", "utils/equi_to_cube.py", "utils/cube_to_equi.py", "utils/sph_utils.py",
         "utils/utils.py", "static_model/class_activation_model.py", "config.yaml", "LICENSE"]


def _has(root):
    return bool(root) and os.path.isfile(os.path.join(root, "model", "cube_pad.py"))


def ref_root():
    """Directory the reference is read from, or None: the live checkout first, else the staged copy."""
    if _has(SOURCE):
        return SOURCE
    if _has(STAGED):
        return STAGED
    return None


REF_ROOT = ref_root() or SOURCE


def available() -> bool:
    return ref_root() is not None


def kind() -> str:
    r = ref_root()
    return "live" if r == SOURCE and _has(SOURCE) else ("staged" if r else "absent")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(src=None, dst=None):
    """Copy the hot-path files of the reference, byte for byte, into oracle/_ref/ (git-ignored) and
    write MANIFEST.json. Returns the manifest, or None when the reference is not present."""
    src = src or SOURCE
    dst = dst or STAGED
    if not _has(src):
        return None
    manifest = {"source": "hsientzucheng/CP-360-Weakly-Supervised-Saliency (unmodified copy; test/baseline use only)",
                "files": {}}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest["files"][rel] = _sha(s)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


def verify(root=None):
    """True iff every staged file still has the sha256 recorded at staging time."""
    root = root or STAGED
    try:
        with open(os.path.join(root, "MANIFEST.json")) as f:
            files = json.load(f)["files"]
    except (OSError, ValueError, KeyError):
        return False
    return bool(files) and all(os.path.isfile(os.path.join(root, rel)) and _sha(os.path.join(root, rel)) == h
                               for rel, h in files.items())


def _install_stubs(root=None):
    root = root or ref_root()
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
    if root and root not in sys.path:
        sys.path.insert(0, root)


def _exec_patched(rel, modname, device):
    root = ref_root()
    path = os.path.join(root, rel)
    with open(path) as f:
        src = f.read()
    if device == "cuda":
        src = src.replace(".cuda(async=True)", ".cuda(non_blocking=True)")
    else:
        src = src.replace(".cuda(async=True)", "")
        src = src.replace(", requires_grad=True).cuda()", ")")
    mod = types.ModuleType(modname)
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def load(device="cpu"):
    """Returns (cube_pad_module, equi_to_cube_module, cube_to_equi_module).

    device="cpu": cube_to_equi is patched to run without CUDA (fixtures, CPU baseline);
    device="cuda": only ``async`` -> ``non_blocking`` (the reference's own GPU path)."""
    if not available():
        raise RuntimeError("reference not present (neither %s nor %s)" % (SOURCE, STAGED))
    _install_stubs()
    cube_pad = importlib.import_module("model.cube_pad")
    e2c = importlib.import_module("utils.equi_to_cube")
    c2e = _exec_patched(os.path.join("utils", "cube_to_equi.py"), "utils.cube_to_equi_" + device, device)
    return cube_pad, e2c, c2e


def load_models():
    """Returns (cube_pad, resnet_cubic, clstm) modules of the reference (model/*.py, unpatched)."""
    cube_pad = load()[0]
    return cube_pad, importlib.import_module("model.resnet_cubic"), importlib.import_module("model.clstm")


def force_cpu_cubepad(cube_pad):
    """The reference hard-wires ``use_gpu=True`` at its call sites (resnet_cubic.py:71,116-117, clstm.py:38;
    cube_pad.py:75-78 then builds CUDA index tensors). On a box without a GPU make False the default."""
    if not getattr(cube_pad.CubePad, "_cp360_cpu_default", False):
        orig = cube_pad.CubePad.__init__

        def cpu_init(self, lrtd_pad, use_gpu=False):
            orig(self, lrtd_pad, use_gpu=False)
        cube_pad.CubePad.__init__ = cpu_init
        cube_pad.CubePad._cp360_cpu_default = True
