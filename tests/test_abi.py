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


def test_peer_sync_host_side_checks_do_not_need_a_gpu():
    """csrc/peer_sync.cu: buffer sizing and argument validation are host code."""
    from keisei_b200 import _lib
    from keisei_b200.distributed import KbPeerCtx
    lib = _lib.load()
    # data [slots][world][slot_doubles] doubles + flags [slots][world] u64
    assert lib.kb_peer_buffer_bytes(8, 4, 512) == 4 * 8 * 512 * 8 + 4 * 8 * 8
    assert lib.kb_peer_buffer_bytes(17, 4, 512) < 0 and lib.kb_peer_buffer_bytes(2, 1, 512) < 0
    assert ctypes.sizeof(KbPeerCtx) == 16 * 8 + 4 * 4 + 8 + 8 + 8 + 8   # + status word pointer, timeout_ms
    ctx = KbPeerCtx()
    ctx.rank, ctx.world, ctx.n_slots, ctx.slot_doubles = 3, 2, 4, 512          # rank outside the world
    dummy = ctypes.c_void_p(16)
    assert lib.kb_peer_allreduce_f64(dummy, 8, ctypes.byref(ctx), 0, None) != 0 and b"bad context" in lib.kb_last_error()
    ctx.rank = 0
    assert lib.kb_peer_allreduce_f64(dummy, 4096, ctypes.byref(ctx), 0, None) != 0 and b"exceed the slot size" in lib.kb_last_error()
    assert lib.kb_peer_allreduce_f64(dummy, 8, ctypes.byref(ctx), 0, None) != 0 and b"not mapped" in lib.kb_last_error()
    ctx.world = 1                                                                # a single rank is a no-op, no launch
    assert lib.kb_peer_allreduce_hook(ctypes.byref(ctx), dummy, 8, None) == 0 and ctx.seq == 1
