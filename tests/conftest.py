import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """Plain `pytest` on a host without CUDA: gpu-marked tests are skipped instead of failing (the product has no CPU
    fallback, so they cannot run there)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200): run with `-m gpu` on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, built in-tree on first use."""
    sys.path.insert(0, os.path.join(ROOT, "smart-nar_fast_tts_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("fs2_build", os.path.join(ROOT, "smart-nar_fast_tts_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from smart_nar_fast_tts_b200 import load_library
    return load_library()
