"""The C-ABI library loads on a CPU-only box and exports every symbol include/keisei_b200.h declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_are_exported():
    from keisei_b200 import _lib, model_ops  # noqa: F401  (registers the model signatures)
    lib = _lib.load()
    header = (ROOT / "include" / "keisei_b200.h").read_text()
    names = set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in keisei_b200.h but not exported"
    assert lib.kb_abi_version() == 1 and lib.kb_compiled_sm() == 100
    for n in _lib._SIGS:
        assert n in names, f"{n} bound in Python but missing from the header"


def test_argument_errors_do_not_need_a_gpu():
    from keisei_b200 import _lib
    lib = _lib.load()
    rc = lib.kb_gae_scan(None, None, None, 5, None, None, None, None, 4, 4, 0.99, 0.95, 0, None)
    assert rc != 0 and b"term_kind" in lib.kb_last_error()


def test_product_package_never_imports_the_oracle():
    for py in (ROOT / "keisei_b200").rglob("*.py"):
        txt = py.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, py
