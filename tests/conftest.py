import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    _build_missing_artefacts()


def _build_missing_artefacts():
    """The native pieces are built in-tree (python __graft_entry__.py) and are git-ignored; in a fresh checkout build the ones that
    are MISSING before collection, so that nothing is skipped or fails for lack of a build step.  Existing files are never
    rebuilt here (the GPU box gets them with the snapshot)."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "vectorvisualization_b200", "libvv_b200.so")
    if not os.path.exists(lib) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        from vectorvisualization_b200 import build as b
        b.build()
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libvv_ref.so")
    if not os.path.exists(ref_lib) and os.path.isdir("/root/reference/VectorVisualization"):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "build_ref.py")], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    """oracle/libvv_oracle.so (CPU restatement) -- the checker, never the product"""
    from oracle import vvo
    vvo.lib()
    return vvo


@pytest.fixture(scope="session")
def vv():
    import vectorvisualization_b200 as vv
    vv.load_library()
    return vv
