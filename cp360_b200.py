"""Import alias: the package directory is named ``cp-360-weakly-supervised-saliency_b200`` (not a
valid Python identifier), so ``import cp360_b200`` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cp-360-weakly-supervised-saliency_b200")
_spec = importlib.util.spec_from_file_location(
    "cp360_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cp360_b200"] = _mod
_spec.loader.exec_module(_mod)

if __name__ == "__main__":
    # the reference's own self-test (`python model/cube_pad.py`, cube_pad.py:257-262; BASELINE.json configs[0])
    import numpy as np
    import torch
    aa = torch.FloatTensor(np.zeros([12, 64, 256, 256])).cuda()
    cp = _mod.CubePad(2)
    print(cp(aa).size())
