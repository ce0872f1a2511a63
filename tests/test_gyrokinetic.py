"""f2: PIC::GYROKINETIC::Mover_FirstOrder / Mover_SecondOrder (src/pic/gyro/gyro_mover.cpp:383-720) with EvalRHS (:245-336) and
CommitReducedStateAndVelocity (:338-381) on coupler fields.  The reduced state is (x, v_parallel, mu).
CPU: the oracle restatement against drift physics.  GPU: (block,cell) assignment and statistics bit-exact, x, v, v_parallel
within 1e-12."""
import numpy as np
import pytest

from amps_b200 import _capi, api
from tests import tp_util as tp

GK1, GK2 = _capi.MOVER_GYROKINETIC_FIRST_ORDER, _capi.MOVER_GYROKINETIC_SECOND_ORDER


def make_case(**kw):
    m, cfg, parts, bg, gradB = tp.make_gc_case(**kw)
    cfg.carry_v_parallel = 1
    x, v = parts[0], parts[1]
    # the reduced state a host would set at injection (PB::SetVParallel / SetMagneticMoment): from v and the analytic field
    uB = kw.get("uniform_B")
    E, B = tp._bg_analytic(x.T.copy(), uniform_B=uB, E_uniform=kw.get("E_uniform"), convection=kw.get("convection", True))
    Bn = np.linalg.norm(B, axis=1)
    b = B / Bn[:, None]
    vpar = (v.T * b).sum(1)
    mu = 0.5 * tp.MP * ((v ** 2).sum(0) - vpar ** 2) / Bn
    return m, cfg, parts, bg, gradB, mu, vpar


def run_oracle(m, cfg, parts, bg, gradB, mu, vpar, mover, n_threads=1):
    o = tp.Oracle(cfg, m)
    o.set_background(*bg)
    o.set_background_gradB(gradB)
    o.add_particles(*parts)
    o.set_reduced_state(mu, vpar)
    rc, st, ret, fc = o.move(mover, n_threads)
    pp = o.particles()
    out = {"rc": rc, "stats": st, "final_cell": fc, "particles": pp, "vpar": o.v_parallel(), "mu": o.magnetic_moment()[0], "lists": o.check_lists()}
    o.close()
    return out


def run_gpu(m, cfg, parts, bg, gradB, mu, vpar, mover):
    g = api.Context(cfg, m)
    g.background_upload(*bg)
    g.background_upload_gradB(gradB)
    g.particles_upload(*parts)
    g.magnetic_moment_upload(mu)
    g.v_parallel_upload(vpar)
    st = g.MoveParticles(mover, raise_on_particle_error=False)
    moved = g.particles_download()
    vp = g.v_parallel_download()
    g.sort()
    srt = g.particles_download()
    vp_sorted, mu_sorted = g.v_parallel_download(), g.magnetic_moment_download()
    n_after = g.particle_count()
    g.close()
    return {"stats": st, "moved": moved, "vpar": vp, "sorted": srt, "vpar_sorted": vp_sorted, "mu_sorted": mu_sorted, "n_after": n_after}


@pytest.mark.parametrize("mover", [GK1, GK2])
def test_uniform_field_ExB_drift(mover):
    # uniform B z, E y, grad B = 0: dx/dt = E x B / B^2 + v_par b, v_par constant, v = b v_par + v_drift
    Bu, Eu = (0.0, 0.0, 2.0e-5), (0.0, 1.0e-3, 0.0)
    m, cfg, parts, bg, gradB, mu, vpar = make_case(n_particles=1024, uniform_B=Bu, E_uniform=Eu, sphere=False, dt=0.01)
    r = run_oracle(m, cfg, parts, bg, gradB, mu, vpar, mover)
    assert r["rc"] == 0 and r["lists"] == 0 and r["stats"]["n_error"] == 0
    alive = r["final_cell"] >= 0
    assert alive.sum() > 900
    x0 = parts[0][:, alive]
    x1, v1 = r["particles"]["x"][:, alive], r["particles"]["v"][:, alive]
    dx = (x1 - x0) / cfg.time_step[0]
    assert np.allclose(dx[0], 1.0e-3 / 2.0e-5, rtol=1e-9) and np.abs(dx[1]).max() < 1e-3
    assert np.allclose(dx[2], vpar[alive], rtol=1e-9)
    assert (r["vpar"][alive] == vpar[alive]).all() and (r["mu"][alive] == mu[alive]).all()
    assert np.allclose(v1[0], 50.0, rtol=1e-9) and np.allclose(v1[2], vpar[alive], rtol=1e-9)


@pytest.mark.parametrize("mover", [GK1, GK2])
def test_dipole_energy_and_threads(mover):
    # E = 0: m v_par^2/2 + mu B is conserved by the guiding-centre equations
    m, cfg, parts, bg, gradB, mu, vpar = make_case(n_particles=4096, dt=0.005, convection=False, sphere=False)
    a = run_oracle(m, cfg, parts, bg, gradB, mu, vpar, mover)
    b = run_oracle(m, cfg, parts, bg, gradB, mu, vpar, mover, n_threads=4)
    assert a["rc"] == 0 and a["lists"] == 0
    assert (a["final_cell"] == b["final_cell"]).all() and a["stats"] == b["stats"] and (a["vpar"] == b["vpar"]).all()
    alive = a["final_cell"] >= 0
    assert alive.sum() > 3000
    x0, x1 = parts[0][:, alive].T, a["particles"]["x"][:, alive].T
    e0 = 0.5 * tp.MP * vpar[alive] ** 2 + mu[alive] * np.linalg.norm(tp.dipole(x0), axis=1)
    e1 = 0.5 * tp.MP * a["vpar"][alive] ** 2 + mu[alive] * np.linalg.norm(tp.dipole(x1), axis=1)
    tol = 2e-2 if mover == GK1 else 1e-2  # (limited by the tabulated field, not by the integrator)
    assert np.median(np.abs(e1 / e0 - 1.0)) < tol
    assert np.linalg.norm(x1 - x0, axis=1).max() > 0.0


CASES = {
    "dipole_linear": dict(),
    "dipole_constant": dict(interp=_capi.CPLR_CONSTANT),
    "dipole_sphere": dict(dt=0.05),
    "uniform_ExB": dict(uniform_B=(1.0e-6, -2.0e-6, 2.0e-5), E_uniform=(2.0e-4, 1.0e-3, 0.0), sphere=False, dt=0.02),
    "amr_dipole": dict(amr_levels=2, n_blocks=4, dt=0.05),
}


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    s = np.maximum(np.abs(b), 1e-300)
    return float((np.abs(a - b) / s).max()) if a.size else 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("mover", [GK1, GK2])
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_parity(name, mover):
    m, cfg, parts, bg, gradB, mu, vpar = make_case(n_particles=8192, seed=17, **CASES[name])
    ora = run_oracle(m, cfg, parts, bg, gradB, mu, vpar, mover)
    gpu = run_gpu(m, cfg, parts, bg, gradB, mu, vpar, mover)
    assert ora["rc"] in (0, _capi.ERR_PARTICLE)
    n = parts[0].shape[1]
    mv = gpu["moved"]
    gx, gv, gc, gp = np.empty((3, n)), np.empty((3, n)), np.empty(n, dtype=np.int64), np.empty(n)
    gx[:, mv["ptrs"]], gv[:, mv["ptrs"]], gc[mv["ptrs"]], gp[mv["ptrs"]] = mv["x"], mv["v"], mv["cells"], gpu["vpar"]
    oc = ora["final_cell"].astype(np.int64)
    alive = oc >= 0
    assert (gc == oc).all()                       # bit-exact block/cell assignment and deletions
    assert gpu["stats"] == ora["stats"]
    assert alive.sum() > n // 2
    tol = 1e-12
    assert rel(gx[:, alive], ora["particles"]["x"][:, alive]) < tol
    assert rel(gp[alive], ora["vpar"][alive]) < tol
    vs = np.abs(ora["particles"]["v"][:, alive]).max(axis=0)  # components of v that cancel are compared against |v|
    assert (np.abs(gv[:, alive] - ora["particles"]["v"][:, alive]) <= 1e-10 * vs).all()
    assert gpu["n_after"] == int(alive.sum())
    # the reduced state travels with the sort
    s = gpu["sorted"]
    assert rel(gpu["vpar_sorted"], ora["vpar"][s["ptrs"]]) < tol and (gpu["mu_sorted"] == mu[s["ptrs"]]).all()
    nbit = int((gx[:, alive] != ora["particles"]["x"][:, alive]).sum())
    print(name, mover, ora["stats"], "x words that differ by rounding:", nbit)
