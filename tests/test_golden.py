"""Golden fixture tests/golden/ecsim_box_16x8x8.npz (made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from tests import parity_util as pu
from tests.golden.make_golden import CASE

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ecsim_box_16x8x8.npz"))


def _inputs():
    m, cfg, parts, fields = pu.make_case(**CASE)
    # the fixture, not the generator, is the input of record
    parts = (G["x"], G["v"], G["w"], G["species"], G["cells"])
    fields = (G["E_half"], G["B_prev"], G["B_cur"])
    return m, cfg, parts, fields


def test_generator_reproduces_fixture_inputs():
    m, cfg, parts, fields = pu.make_case(**CASE)
    assert (parts[0] == G["x"]).all() and (parts[1] == G["v"]).all() and (parts[4] == G["cells"]).all()


def test_oracle_matches_golden():
    m, cfg, parts, fields = _inputs()
    ora = pu.run_oracle(m, cfg, parts, fields)
    assert (ora["final_cell"] == G["final_cell"]).all()
    assert (ora["particles"]["x"] == G["x_out"]).all() and (ora["particles"]["v"] == G["v_out"]).all()
    assert [ora["stats"][k] for k in sorted(ora["stats"]) if k != "n_sub_steps"] == list(G["stats"])  # fixture predates n_sub_steps
    assert (ora["J"] == G["J"]).all() and (ora["M"] == G["M"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [1, 0])
def test_gpu_matches_golden(exact):
    m, cfg, parts, fields = _inputs()
    cfg.exact_arithmetic = exact
    gpu = pu.run_gpu(m, cfg, parts, fields)
    n = parts[0].shape[1]
    mv = gpu["moved"]
    gx, gv, gc = np.empty((3, n)), np.empty((3, n)), np.empty(n, dtype=np.int64)
    gx[:, mv["ptrs"]], gv[:, mv["ptrs"]], gc[mv["ptrs"]] = mv["x"], mv["v"], mv["cells"]
    assert (gc == G["final_cell"]).all()                      # bit-exact cell assignment
    if exact:   # every operation rounded like the CPU build
        assert (gx == G["x_out"]).all() and (gv == G["v_out"]).all()
    else:       # contracted arithmetic away from cell faces: a few ulp (contract: 1e-10)
        alive = G["final_cell"] >= 0
        assert pu.rel_elementwise(gx[:, alive], G["x_out"][:, alive]) <= 1e-12
        assert pu.rel_elementwise(gv[:, alive], G["v_out"][:, alive]) <= 1e-10
    assert [gpu["stats"][k] for k in sorted(gpu["stats"]) if k != "n_sub_steps"] == list(G["stats"])
    assert pu.rel_scaled(gpu["J"], G["J"]) <= pu.REL_TOL and pu.rel_scaled(gpu["M"], G["M"]) <= pu.REL_TOL
    assert abs(gpu["energy"] - float(G["energy"])) <= pu.REL_TOL * float(G["energy"])
