import os
import sys
from pathlib import Path

import pytest

# the test-suite builds seeded random-weight models; no attempt to fetch the pretrained 'gpt2' checkpoint (there is no
# network here or on the GPU box).  tests/test_pretrained_cpu.py switches it back on around a fake `from_pretrained`.
os.environ.setdefault("CAPDEC_GPT2_PRETRAINED", "0")

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
