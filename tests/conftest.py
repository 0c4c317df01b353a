"""pytest config: `gpu` marker; CPU-only runs use `-m "not gpu"`."""
import os
import sys
from pathlib import Path

# Thread-emulated ranks (tests/test_sync_bn_gpu.py) run two spin-waiting exchange kernels on ONE GPU: their streams must
# not share a hardware work queue (false serialisation = deadlock until the exchange times out). The default is 8 queues;
# real multi-process runs (one GPU per rank) are not affected. Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / name) as z:
        return {k: z[k] for k in z.files}


def state_dict_from(g: dict, prefix: str = "sd/") -> dict:
    return {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith(prefix)}


@pytest.fixture(scope="session")
def golden():
    return load_golden
