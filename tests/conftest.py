import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "stages.pt"), map_location="cpu")


@pytest.fixture(scope="session")
def weights():
    """Synthetic checkpoint (seed 0), infer-path tensors only, fp32 CPU."""
    from detail_tts_b200 import synth
    return synth.synth_state_dict(0, keys=synth.infer_path_key)


@pytest.fixture(scope="session")
def dlib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from detail_tts_b200 import _lib
    return _lib.lib()
