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


def test_struct_layouts_match_the_header(tmp_path):
    # sizes and a few offsets as the C compiler sees them; a mismatch would corrupt every call
    import subprocess

    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "amps_gpu.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(amps_gpu_config), sizeof(amps_gpu_mesh),"
        " sizeof(amps_gpu_exit_record), sizeof(amps_gpu_move_stats), sizeof(amps_gpu_aos_layout),"
        " offsetof(amps_gpu_config, capacity), offsetof(amps_gpu_config, speed_of_light), offsetof(amps_gpu_config, carry_magnetic_moment));"
        "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(t) for t in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_capi.Config), ctypes.sizeof(_capi.Mesh), ctypes.sizeof(_capi.ExitRecord), ctypes.sizeof(_capi.MoveStats),
            ctypes.sizeof(_capi.AosLayout), _capi.Config.capacity.offset, _capi.Config.speed_of_light.offset,
            _capi.Config.carry_magnetic_moment.offset]
    assert got == want, (got, want)


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


def test_a_missing_library_fails_loudly_also_through_the_variant_override(tmp_path):
    # AMPS_GPU_LIB only points the loader at another build of the same library; a path that does not exist must raise
    # (there is no CPU fallback behind it)
    import subprocess
    import sys

    code = ("import os; os.environ['AMPS_GPU_LIB'] = r'%s'\n"
            "from amps_b200 import _capi\n"
            "try:\n    _capi.load_library()\nexcept RuntimeError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\n" % (tmp_path / "nope.so"))
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT, text=True)
    assert "RAISED True" in out, out
