"""f4 (first half): ECSIM::ComputeNetCharge (pic_field_solver_ecsim.cpp:4690-4828).  CPU: total charge is conserved by the
restatement; GPU: rho on the unique centre nodes within 1e-10 of the array maximum."""
import numpy as np
import pytest

from amps_b200 import api
from oracle.oracle_py import Oracle
from tests import parity_util as pu


def _oracle_rho(m, cfg, parts, conv=1.0):
    x, v, w, sp, cells = parts
    o = Oracle(cfg, m)
    o.add_particles(x, v, w, sp, cells)
    rho = o.net_charge(conv)
    o.close()
    return rho


def test_total_charge_is_conserved_cpu():
    m, cfg, parts, fields = pu.make_case(n_cells=(16, 16, 16), ppc=4, seed=41)
    rho = _oracle_rho(m, cfg, parts)
    x, v, w, sp, cells = parts
    q = np.array([cfg.charge[s] for s in range(cfg.n_species)])[sp] * np.array([cfg.species_weight[s] for s in range(cfg.n_species)])[sp] * w
    assert abs(rho.sum() * 1.0 - q.sum()) <= 1e-9 * np.abs(q).sum()    # unit cells: sum rho * V = sum q (periodic: nothing is dropped)
    only_e = (parts[0][:, sp == 0], v[:, sp == 0], w[sp == 0], sp[sp == 0], cells[sp == 0])
    assert (_oracle_rho(m, cfg, only_e) <= 0).all()                     # electrons alone: negative everywhere


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(n_cells=(16, 16, 16), ppc=8, seed=43), dict(n_cells=(16, 16, 16), ppc=6, seed=45, periodic=False),
                                dict(n_cells=(32, 16, 8), ppc=5, seed=47, block_cells=(16, 8, 4)), dict(n_cells=(16, 16, 16), ppc=4, seed=49, ghost_cells=(2, 2, 2)),
                                dict(n_cells=(16, 16, 16), ppc=6, seed=51, four_species=True)])
def test_gpu_net_charge_matches_oracle(kw):
    m, cfg, parts, fields = pu.make_case(**kw)
    ref = _oracle_rho(m, cfg, parts, 0.7)
    x, v, w, sp, cells = parts
    g = api.Context(cfg, m)
    g.particles_upload(x, v, w, sp, cells)
    rho = g.ComputeNetCharge(0.7)
    g.close()
    assert pu.rel_scaled(rho, ref) <= pu.REL_TOL
    assert np.abs(ref).max() > 0
