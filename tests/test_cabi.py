"""The C-ABI shared library loads on a CPU-only box and exports every symbol ``include/slode_b200.h`` declares
(no compute calls here: those need a GPU)."""
import ctypes
import os
import re

import pytest

from structured_latent_odes_b200 import _build, _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "slode_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slode_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def handle():
    _build.build()
    return ctypes.CDLL(_cabi.LIB_PATH)


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for want in ("slode_query", "slode_last_error", "slode_mlp_supported", "slode_mlp_fixed_fwd", "slode_mlp_fixed_bwd"):
        assert want in syms


def test_library_exports_every_declared_symbol(handle):
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} is declared in slode_b200.h but not exported"


def test_ctypes_signatures_cover_the_header():
    assert sorted(_cabi.SIGNATURES) == declared_symbols()


def test_query_without_a_gpu(handle):
    L = _cabi.lib()
    assert L.slode_query(_cabi.Q_SM_ARCH) == 100
    assert L.slode_query(_cabi.Q_VERSION) >= 1
    shapes = _cabi.supported_shapes()
    assert (25, 5) in shapes and (25, 8) in shapes  # CVS / challenge and proc configs
    assert L.slode_mlp_supported(25, 5) == 1 and L.slode_mlp_supported(7, 3) == 0
    assert L.slode_query(12345) == -1


def test_argument_validation_happens_before_any_cuda_call(handle):
    L = _cabi.lib()
    # unsupported (H,S): no generic fallback
    rc = L.slode_mlp_fixed_fwd(_cabi.METHOD_RK4, 4, 3, 7, 3, *([None] * 8), None, 0, 0, None, 0, None)
    assert rc == 2 and b"not compiled in" in L.slode_last_error()
    with pytest.raises(NotImplementedError):
        _cabi.check(rc, "fwd")
    rc = L.slode_mlp_fixed_fwd(_cabi.METHOD_RK4, -1, 3, 25, 5, *([None] * 8), None, 0, 0, None, 0, None)
    assert rc == 1
    rc = L.slode_mlp_fixed_fwd(_cabi.METHOD_RK4, 4, 3, 25, 5, *([None] * 8), None, 0, 0, None, 0, None)
    assert rc == 1 and b"null pointer" in L.slode_last_error()
    with pytest.raises(_cabi.SlodeError):
        _cabi.check(rc, "fwd")
    rc = L.slode_mlp_fixed_bwd(_cabi.METHOD_RK4, 7, 4, 3, 25, 5, *([None] * 7), None, 0, 0, None, 0, 0, None, None, None, None, 0, None)
    assert rc == 1 and b"unknown mode" in L.slode_last_error()
    # the workspace query validates like the entry points (no CUDA call for rejected arguments)
    assert L.slode_fixed_workspace_bytes(1, _cabi.METHOD_RK4, 0, 4, 3, 15, 7, 3, 2, 0) == -1
    assert L.slode_fixed_workspace_bytes(1, 9, 0, 4, 3, 15, 25, 5, 2, 0) == -1
    assert L.slode_fixed_workspace_bytes(1, _cabi.METHOD_RK4, 0, 0, 3, 15, 25, 5, 2, 0) == 0
    assert L.slode_dopri5_supported(25, 5) == 1 and L.slode_dopri5_supported(512, 5) == 0
    assert L.slode_mlp_supported(512, 5) == 1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.SlodeError, match="no fallback"):
        _cabi.lib()


def test_library_was_built_from_the_sources_in_the_tree(handle):
    """slode_query(SLODE_Q_SOURCE_HASH) is the digest of csrc/ + the public header at build time: a prebuilt .so that
    no longer matches the sources is detected (and __graft_entry__.build() rebuilds the unit that carries it)."""
    from structured_latent_odes_b200 import _build, _cabi
    assert handle.slode_query(_cabi.Q_SOURCE_HASH) == _build.source_hash()
