"""CPU checks of the drop-in boundary: the library builds/loads and exports every
symbol include/sejonggo_b200.h declares; the host refuses to run without a GPU."""
import ctypes as C
import pytest
import torch

from sejonggo_b200 import _abi, _build


def test_library_exports_every_header_symbol():
    so = _build.build()
    lib = C.CDLL(so)
    syms = _abi.header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s
    assert _abi.load().sgo_abi_version() == 1


def test_every_symbol_has_a_prototype():
    for s in _abi.header_symbols():
        assert s in _abi._PROTOS or s in ("sgo_last_error", "sgo_launch_count"), s


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from sejonggo_b200.engine import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(size=9, n_games=1)
    cfg = _abi.SgoConfig(device=0, size=9, n_games=1, trees_per_game=1, max_leaves=1, arena_blocks=2, komi=5.5)
    h = C.c_void_p()
    assert _abi.load().sgo_create(C.byref(cfg), C.byref(h)) != 0


def test_product_library_has_no_ablation_switch():
    """The conv kernel's timing ablations (SGO_CONV_DEBUG) exist only in the --ablate build: the shipped library must not
    read an environment variable that makes the timed kernel skip work."""
    from sejonggo_b200 import _build
    so = _build.build()
    blob = open(so, "rb").read()
    assert b"SGO_CONV_DEBUG" not in blob          # (getenv itself is imported by the static CUDA runtime)
