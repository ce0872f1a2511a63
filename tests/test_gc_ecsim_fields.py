"""The guiding-centre movers on the ECSIM arrays (cfg.gc_fields_ecsim): the reference built with the ECSIM field solver reads
ECSIM::GetMagneticField / GetElectricField / GetMagneticFieldGradient (pic_field_solver_ecsim.cpp:7440-7547) in InitiateMagneticMoment,
GuidingCenterMotion and the final-field update of Mover_FirstOrder (pic_mover_guiding_center.cpp:103, :179-184, :727) instead of the
coupler's tables.

CPU: in fields that are linear in x both interpolations and the half-cell differences are exact, so the oracle's ECSIM branch must
reproduce its coupler branch (fed with the analytic gradient).  GPU: the kernels against the oracle's ECSIM branch."""
import numpy as np
import pytest

from amps_b200 import _capi, api
from oracle.oracle_py import Oracle
from tests import tp_util as tp

GC1, GC2 = _capi.MOVER_GC_FIRST_ORDER, _capi.MOVER_GC_SECOND_ORDER


def ecsim_case(linear=False, n_particles=4096, dt=0.01, seed=21, **kw):
    m, cfg, parts, (Ec, Bc) = tp.make_tp_case(n_particles=n_particles, seed=seed, dt=dt, boundary=_capi.BOUNDARY_DELETE, rigidity_gv=(0.001, 0.02), **kw)
    cfg.carry_magnetic_moment = 1
    cfg.ideal_mhd = 0
    if linear:
        G = np.array([[1.0, -2.0, 0.5], [0.3, 0.7, -1.1], [-0.6, 0.4, -1.7]]) * 2.0e-13  # dB_i/dx_j, trace 0
        B0 = np.array([2.0e-6, -1.0e-6, 3.0e-6])
        fB = lambda x: B0[None, :] + x @ G.T
        fE = lambda x: np.broadcast_to(np.array([1.0e-4, -2.0e-4, 0.5e-4]), x.shape).copy() + 1.0e-11 * x[:, [1, 2, 0]]
        gradB = np.broadcast_to(G.reshape(9), (m.n_centers, 9)).copy()
    else:
        def fB(x):
            r = np.sqrt((x ** 2).sum(1))
            return tp.dipole(np.where(r[:, None] < 0.5 * tp.RE, x + 0.5 * tp.RE, x))

        def fE(x):
            return -np.cross(np.broadcast_to(np.array([-4.0e5, 0.0, 0.0]), x.shape), fB(x))
        gradB = None
    return m, cfg, parts, dict(E_corner=fE(m.corner_x), B_center=fB(m.center_x), E_center=fE(m.center_x), gradB=gradB)


def run_oracle(m, cfg, parts, f, mover, ecsim, pre_init):
    cfg.gc_fields_ecsim = 1 if ecsim else 0
    o = Oracle(cfg, m)
    if ecsim:
        o.set_fields(f["E_corner"], f["B_center"], f["B_center"])
        o.set_E_current(f["E_corner"])
    else:
        o.set_background(f["E_center"], f["B_center"])
        o.set_background_gradB(f["gradB"])
    o.add_particles(*parts)
    if pre_init:
        o.magnetic_moment_init(mover)
    rc, st, ret, fc = o.move(mover, 1)
    pp = o.particles()
    mu, flag = o.magnetic_moment()
    o.close()
    return {"rc": rc, "stats": st, "final_cell": fc, "particles": pp, "mu": mu}


@pytest.mark.parametrize("mover", [GC1, GC2])
def test_oracle_ecsim_branch_equals_the_coupler_branch_in_linear_fields(mover):
    m, cfg, parts, f = ecsim_case(linear=True, sphere=False)
    a = run_oracle(m, cfg, parts, f, mover, True, mover == GC2)
    b = run_oracle(m, cfg, parts, f, mover, False, mover == GC2)
    both = (a["final_cell"] >= 0) & (b["final_cell"] >= 0)
    assert both.sum() > 0.9 * len(both)
    assert (a["final_cell"][both] == b["final_cell"][both]).mean() > 0.999  # a face-grazing point may round to the other cell
    xa, xb = a["particles"]["x"][:, both], b["particles"]["x"][:, both]
    va, vb = a["particles"]["v"][:, both], b["particles"]["v"][:, both]
    assert np.abs(xa - xb).max() <= 1e-9 * np.abs(xb).max()
    assert np.abs(va - vb).max() <= 1e-8 * np.abs(vb).max()
    assert np.abs(a["mu"][both] - b["mu"][both]).max() <= 1e-10 * np.abs(b["mu"][both]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("mover", [GC1, GC2])
@pytest.mark.parametrize("sphere", [True, False])
def test_gpu_gc_movers_on_ecsim_fields_match_the_oracle(mover, sphere):
    m, cfg, parts, f = ecsim_case(linear=False, n_particles=8192, sphere=sphere)
    pre = mover == GC2
    ora = run_oracle(m, cfg, parts, f, mover, True, pre)
    assert ora["rc"] in (0, _capi.ERR_PARTICLE)
    cfg.gc_fields_ecsim = 1
    g = api.Context(cfg, m)
    g.fields_upload(f["E_corner"], f["B_center"], f["B_center"])
    g.E_upload(f["E_corner"])
    g.particles_upload(*parts)
    if pre:
        g.InitiateMagneticMoment(mover)
    st = g.MoveParticles(mover, raise_on_particle_error=False)
    mv = g.particles_download()
    mu_dev = g.magnetic_moment_download()
    g.close()
    n = parts[0].shape[1]
    gx, gv, gc, gmu = np.full((3, n), np.nan), np.full((3, n), np.nan), np.full(n, -1, dtype=np.int64), np.full(n, np.nan)
    gx[:, mv["ptrs"]], gv[:, mv["ptrs"]], gc[mv["ptrs"]], gmu[mv["ptrs"]] = mv["x"], mv["v"], mv["cells"], mu_dev
    oc = ora["final_cell"].astype(np.int64)
    alive = oc >= 0
    assert alive.sum() > 0.5 * n
    assert (gc == oc).all()
    assert st == ora["stats"]
    rel = lambda a, b: float((np.abs(a - b) / np.maximum(np.abs(b), 1e-300)).max())
    assert rel(gx[:, alive], ora["particles"]["x"][:, alive]) < 1e-12
    assert rel(gv[:, alive], ora["particles"]["v"][:, alive]) < 1e-9
    assert rel(gmu[alive], ora["mu"][alive]) < 1e-12
