"""pytest plugin used by tests/test_dropin_reference.py: installs the keisei_b200 shim into the importable reference
BEFORE the reference's test modules are collected (they bind `KataGoPPOAlgorithm` & co. by name at import time)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def pytest_configure(config):
    import keisei_b200.dropin as dropin
    dropin.install_into_reference()
