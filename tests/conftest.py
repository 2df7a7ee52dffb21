import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
BUILT = os.path.join(ROOT, "assets", "_built")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """GPU-marked tests need a CUDA device and the built library: skip (not fail) them elsewhere, so that a plain
    `pytest tests` on a CPU box runs the CPU suite.  On a GPU box nothing is skipped: a missing libvapb200.so fails
    loudly there (the product has no fallback path)."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); run with -m gpu on the GPU box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def built_asset(name):
    p = os.path.join(BUILT, name)
    if not os.path.exists(p):
        pytest.skip(f"{p} missing: run tools/prepare_assets.py in the build container")
    return p


@pytest.fixture(scope="session")
def vap_weights():
    from vap_realtime_b200 import weights
    return weights.load(built_asset("vap_jp_20hz_2500msec.vapw"))


@pytest.fixture(scope="session")
def bc_weights():
    from vap_realtime_b200 import weights
    return weights.load(built_asset("vap_bc_erica_20hz_5000msec.vapw"))


@pytest.fixture(scope="session")
def fixture_audio():
    d = np.load(os.path.join(GOLDEN, "ref_vap_ctx2500.npz"))
    return d["audio"].astype(np.float32) / 32768.0, d["out"]


def chunk(audio, n, hz=20):
    shift = 16000 // hz
    return audio[..., shift * n: shift * n + shift + 320]
