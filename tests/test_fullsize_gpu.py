"""Parity at the benchmark's own size: BASELINE configs[1] -- 64^3 cells, 64 ppc per species, e/p,
3.36e7 particles -- one ECSIM particle phase through the C ABI against the CPU oracle (all host threads; the OpenMP
variant of the oracle differs from its serial one by summation order only, tests/test_oracle_cpu.py).

Bars: (block, cell) of every particle and the crossing counters bit-exact; x', v' BIT-IDENTICAL with cfg.exact_arithmetic (the
reference's operation order) and <= 1e-10 of the particle's |x|, |v| with the production (contracted) mover -- among 1e8 velocity
components some pass within 1e-7 of zero, where an element-wise quotient only measures that absolute 1e-16;
J, M <= 1e-10 of the array maximum AND <= 1e-10 relative element-wise on every entry that is not a cancellation residue
(|entry| >= 1e-3 of the array maximum); energy, cfl <= 1e-10."""
import os

import numpy as np
import pytest

from tests import parity_util as pu


def significant_rel(a, b, floor=1e-3):
    s = np.abs(b).max()
    sel = np.abs(b) >= floor * s
    return float((np.abs(a[sel] - b[sel]) / np.abs(b[sel])).max()), int(sel.sum())


@pytest.mark.gpu
@pytest.mark.timeout(1500)
def test_u64_full_size_step_matches_the_oracle():
    m, cfg, parts, fields = pu.make_case(n_cells=(64, 64, 64), ppc=64, seed=100, E_amp=0.01)
    assert parts[0].shape[1] == 64 ** 3 * 128
    threads = os.cpu_count() or 1
    ora = pu.run_oracle(m, cfg, parts, fields, n_threads=threads)
    gpu = pu.run_gpu(m, cfg, parts, fields)
    res = pu.compare(m, parts, ora, gpu)
    relJ, nJ = significant_rel(gpu["J"], ora["J"])
    relM, nM = significant_rel(gpu["M"], ora["M"])
    print({k: res[k] for k in ("n", "cell_mismatch", "max_rel_x", "max_rel_v", "max_relnorm_x", "max_relnorm_v", "max_rel_J", "max_rel_M", "rel_energy", "rel_cfl", "n_redo")},
          "elementwise J", relJ, nJ, "M", relM, nM)
    assert res["cell_mismatch"] == 0 and res["stats_equal"], res
    assert res["max_relnorm_x"] <= 1e-10 and res["max_relnorm_v"] <= 1e-10, res
    assert res["sorted_ok"] and res["table_ok"] and res["perm_ok"], res
    assert res["max_rel_J"] <= 1e-10 and res["max_rel_M"] <= 1e-10, res
    assert relJ <= 1e-10 and relM <= 1e-10 and nM > 1e6, (relJ, relM, nJ, nM)
    assert res["rel_energy"] <= 1e-10 and res["rel_cfl"] <= 1e-10, res
    # the exact mover: every particle bit for bit
    cfg.exact_arithmetic = 1
    gpu = pu.run_gpu(m, cfg, parts, fields)
    res = pu.compare(m, parts, ora, gpu)
    print("exact mover:", {k: res[k] for k in ("cell_mismatch", "bit_mismatch_xv", "max_rel_M")})
    assert res["cell_mismatch"] == 0 and res["stats_equal"] and res["bit_mismatch_xv"] == 0, res
    assert res["max_rel_J"] <= 1e-10 and res["max_rel_M"] <= 1e-10, res
