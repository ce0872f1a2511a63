"""The C-ABI library loads and exports every symbol include/amps_gpu.h declares (no compute without a GPU)."""
import ctypes
import os
import re

from amps_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "amps_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(amps_gpu_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/amps_gpu.h but not exported by libamps_gpu.so"
    assert set(syms) == set(_capi.PROTOTYPES), set(syms) ^ set(_capi.PROTOTYPES)


def test_struct_layouts_match_the_header():
    # sizes the C compiler gives the structs (kept in sync by hand; a mismatch would corrupt every call)
    assert ctypes.sizeof(_capi.Config) == 3 * 4 + 3 * 4 + 6 * 4 + 8 + 4 * 8 * 8 + 4 * 8 + 2 * 4 + 2 * 8 + 8 + 8
    assert ctypes.sizeof(_capi.ExitRecord) == 4 * 4 + 6 * 8
    assert ctypes.sizeof(_capi.MoveStats) == 7 * 8
    assert ctypes.sizeof(_capi.AosLayout) == 8 + 7 * 4 + 4


def test_init_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        return
    lib = _capi.load_library()
    cfg = _capi.Config()
    for d in range(3):
        cfg.block_cells[d] = 8
        cfg.ghost_cells[d] = 1
    cfg.n_species, cfg.capacity = 1, 16
    h = ctypes.c_void_p()
    rc = lib.amps_gpu_init(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == _capi.ERR_NO_DEVICE and not h  # no CPU fallback
