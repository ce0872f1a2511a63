"""The structure cache of the coupler's AMR stencil (csrc/cplr_stencil.cuh: neighbour nodes per block, the cells behind the logical
coarse centres per dual cell) must not change a bit: the test-particle movers on the 3-level dipole mesh with the cache (default) and
without it (AMPS_GPU_CPLR_CACHE=0, every stencil built per particle as the reference does) leave identical particles, counters and
exit records.  The same cases are compared with the CPU oracle in test_relativistic_boris.py / test_guiding_center.py."""
import os

import numpy as np
import pytest

from amps_b200 import _capi
from tests import tp_util as tp


@pytest.mark.gpu
@pytest.mark.parametrize("mover", [_capi.MOVER_RELATIVISTIC_BORIS, _capi.MOVER_BORIS])
@pytest.mark.parametrize("ghost", [(1, 1, 1), (2, 2, 2)])
def test_cached_stencils_are_bit_identical(mover, ghost):
    kw = dict(backward=mover == _capi.MOVER_RELATIVISTIC_BORIS, interp=_capi.CPLR_LINEAR, boundary=_capi.BOUNDARY_USER_FUNCTION, sphere=True,
              amr_levels=2, n_blocks=4, ghost_cells=ghost)
    dt = 0.1
    if mover == _capi.MOVER_BORIS:
        dt *= 0.02
        kw["rigidity_gv"] = (0.001, 0.05)
    m, cfg, parts, bg = tp.make_tp_case(n_particles=60000, dt=dt, seed=11, **kw)
    runs = {}
    for sw in ("0", "1"):
        os.environ["AMPS_GPU_CPLR_CACHE"] = sw
        try:
            runs[sw] = tp.run_gpu_tp(m, cfg, parts, bg, mover=mover)
        finally:
            os.environ.pop("AMPS_GPU_CPLR_CACHE", None)
    a, b = runs["0"], runs["1"]
    assert a["rc"] == 0 and b["rc"] == 0
    assert a["stats"] == b["stats"] and a["stats"]["n_error"] == 0
    assert a["n_records"] == b["n_records"] and a["records"] == b["records"]
    oa, ob = np.argsort(a["moved"]["ptrs"]), np.argsort(b["moved"]["ptrs"])
    assert (a["moved"]["ptrs"][oa] == b["moved"]["ptrs"][ob]).all()
    assert (a["moved"]["cells"][oa] == b["moved"]["cells"][ob]).all()
    assert (a["moved"]["x"][:, oa] == b["moved"]["x"][:, ob]).all() and (a["moved"]["v"][:, oa] == b["moved"]["v"][:, ob]).all()
    lev = m.leaf_level()
    assert len(set(int(v) for v in lev)) == 3  # the mesh is the refined one
