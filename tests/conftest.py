import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "reference: needs the reference checkout at /root/reference")
    config.addinivalue_line("markers", "reference_copy: needs the reference checkout or its travelling copy oracle/_ref "
                                       "(oracle/make_ref.py) — the only way a test can see the reference on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(os.environ.get("BB_REFERENCE", "/root/reference"))
    have_copy = have_ref or os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "src", "agents"))
    gpu = None
    for item in items:
        if "reference_copy" in item.keywords and not have_copy:
            item.add_marker(pytest.mark.skip(reason="neither /root/reference nor oracle/_ref present"))
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="reference checkout not mounted"))
        if "gpu" in item.keywords:
            if gpu is None:
                gpu = _has_gpu()
            if not gpu:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))


def reference_paths():
    """sys.path entries that make the reference importable: the checkout if mounted, else oracle/_ref."""
    ref = os.environ.get("BB_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(ref, "src")):
        return [os.path.join(GOLDEN, "_gym_stub"), os.path.join(ref, "src")]
    return [os.path.join(ROOT, "oracle", "_ref", "_gym_stub"), os.path.join(ROOT, "oracle", "_ref", "src")]


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
