import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no libctk.so yet (built artefacts are not in the history): build it when
    # nvcc is available, as __graft_entry__.build() does; on a GPU box the shipped library is used
    lib = os.path.join(ROOT, "clustertracking_b200", "libctk.so")
    if not os.path.exists(lib):
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            from clustertracking_b200 import build
            build.build(verbose=False)


def pytest_collection_modifyitems(config, items):
    """Tests marked ``gpu`` are skipped (not failed) where there is no CUDA device."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
