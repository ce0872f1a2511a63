"""Row f1 on the GPU through the C ABI: amps_gpu_field_step (UpdateRhs -> GMRES -> ProcessFinalSolution -> UpdateB -> UpdateE on the
device, J and M never leaving it) against the numpy oracle (oracle/ecsim_field.py, itself pinned on the reference's own
ECSIM::TimeStep in tests/test_reference_field_solve.py) and against the reference run here.

Tolerances: both sides iterate to |r| <= 1e-12 |r0|, the fields are compared to 1e-9 of their maximum (the solve), the step that
follows to the bit (the particle phase must see exactly the fields the solve left in the tiles)."""
import numpy as np
import pytest

from amps_b200 import api
from oracle import ecsim_field
from oracle.oracle_py import Oracle
from oracle.ref_pic import ref_pic
from tests import parity_util as pu


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.gpu
def test_field_step_matches_the_oracle_and_feeds_the_particle_step():
    m, cfg, parts, fields = pu.make_case(n_cells=(16, 16, 16), ppc=8, seed=5, E_amp=0.01)
    cfg.exact_arithmetic = 1
    E_half0, B_prev0, B_cur0 = fields
    rng = np.random.default_rng(3)
    E_n = 0.01 * rng.standard_normal((m.n_corners, 3))
    g = api.Context(cfg, m)
    g.fields_upload(E_half0, B_prev0, B_cur0)
    g.particles_upload(*parts)
    g.UpdateJMassMatrix()
    J, M = g.JM_download()
    g.field_solver_init()
    g.E_upload(E_n)
    its, res = g.field_step(theta=0.5, tol=1e-12, max_iter=300, restart=40)
    got = g.fields_download()
    s = ecsim_field.EcsimField(m, (1.0, 1.0, 1.0), cfg.ecsim_light_speed, cfg.ecsim_dt_total, theta=0.5)
    Eh, En, Bn, its_o = s.step(E_n, B_cur0, J, M, tol=1e-12, max_iter=300)
    print("iterations gpu", its, "oracle", its_o, "residual", res)
    assert res <= 1e-12 and 5 < its <= 300
    assert rel(got["E_half"], Eh) <= 1e-9 and rel(got["E"], En) <= 1e-9 and rel(got["B"], Bn) <= 1e-9
    # node-local updates given the device's own E^{n+theta}: to rounding
    assert rel(got["B"], s.update_B(B_cur0, got["E_half"])) <= 1e-14 and rel(got["E"], s.update_E(E_n, got["E_half"])) <= 1e-14
    # a restart in the middle of the solve gives the same answer
    g.fields_upload(E_half0, B_prev0, B_cur0)
    g.E_upload(E_n)
    its2, res2 = g.field_step(theta=0.5, tol=1e-12, max_iter=300, restart=7)
    got2 = g.fields_download()
    assert res2 <= 1e-12 and rel(got2["E_half"], Eh) <= 1e-9

    # warm start: the same state solved again from the previous increment converges at once and lands on the same fields
    g.fields_upload(E_half0, B_prev0, B_cur0)
    g.E_upload(E_n)
    its3, res3 = g.field_step(theta=0.5, tol=1e-12, max_iter=300, restart=7, warm_start=True)
    got3 = g.fields_download()
    assert its3 <= 2 and rel(got3["E_half"], Eh) <= 1e-9 and rel(got3["B"], Bn) <= 1e-9, (its3, res3)
    got2 = got3

    # the particle step that follows reads the staged E^{n+theta}, B^n (mover) and B^{n+1} (deposit): compare with the oracle
    # given exactly those fields
    before = g.particles_download()
    st = g.MoveParticles()
    moved = g.particles_download()
    g.sort()
    g.UpdateJMassMatrix()
    J2, M2 = g.JM_download()
    g.close()
    o = Oracle(cfg, m, "parity")
    o.set_fields(got2["E_half"], B_cur0, got2["B"])
    o.add_particles(before["x"], before["v"], before["w"], before["species"], before["cells"])
    rc, st_o, ret, fc = o.move(0, 1)
    pp = o.particles()
    Jo, Mo, _, _ = o.deposit(1)
    o.close()
    assert (moved["cells"] == fc).all() and (moved["x"] == pp["x"]).all() and (moved["v"] == pp["v"]).all()
    assert rel(J2, Jo) <= 1e-10 and rel(M2, Mo) <= 1e-10


@pytest.mark.gpu
@pytest.mark.skipif(not ref_pic.available(), reason="oracle/_ref/libref_pic.so not built")
def test_field_step_matches_the_reference_compiled_here():
    from tests import ref_ecsim_case as rc

    c = rc.case()
    m, cfg, ref = c["mesh"], c["cfg"], c["ref"]
    f = ref["field"]
    x, v, w, sp, cells = c["parts"]
    cfg.exact_arithmetic = 1
    g = api.Context(cfg, m)
    g.fields_upload(*c["fields"])
    g.particles_upload(x, v, w, sp, cells)
    g.MoveParticles()
    g.sort()
    g.UpdateJMassMatrix()  # J, M of the moved particles: what the reference's field step used
    g.field_solver_init()
    g.E_upload(f["E"])
    its, res = g.field_step(theta=f["theta"], tol=1e-12, max_iter=400, restart=60)
    got = g.fields_download()
    g.close()
    print("iterations gpu", its, "reference", f["iterations"])
    assert res <= 1e-12
    assert rel(got["E_half"], f["E_half"]) <= 1e-9 and rel(got["E"], f["E_new"]) <= 1e-9 and rel(got["B"], f["B_new"]) <= 1e-9
