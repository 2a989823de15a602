"""Import the UNMODIFIED reference hot-path modules (thin alias of oracle/ref_loader.py, which documents
the import stubs and the in-memory ``async`` patch). The reference is read from /root/reference in the
build container or from the byte-identical staging copy oracle/_ref/ on the GPU box; nothing of it is
committed. Used by ``make_golden*.py`` (fixture generation) and by the reference-vs-oracle tests, which
skip when neither location exists."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_loader import (REF_ROOT, _install_stubs, available, force_cpu_cubepad, kind, load,  # noqa: E402,F401
                               load_models, ref_root, stage, verify)
