"""PIC::Mover::Relativistic::Boris (a7) + domain exit / internal sphere (a15).

CPU: the restated mover against the REFERENCE's own stand-alone BorisStep compiled from srcEarth/gridless
(oracle/_ref/libref_gridless.so) and against the analytic gyration.  GPU: bit parity with the oracle through the C ABI."""
import numpy as np
import pytest

from amps_b200 import _capi
from tests import tp_util as tp


def _single(x0, v0, B, dt, n_steps):
    """n_steps calls of the oracle mover for one proton in a uniform field"""
    from oracle.oracle_py import Oracle

    m, cfg, parts, bg = tp.make_tp_case(n_particles=1, half_width_re=8.0, n_blocks=4, dt=dt, uniform_B=B, sphere=False,
                                        boundary=_capi.BOUNDARY_DELETE)
    x = np.array(x0, dtype=np.float64).reshape(3, 1)
    v = np.array(v0, dtype=np.float64).reshape(3, 1)
    # cell of x
    parts = (x, v, np.ones(1), np.zeros(1, dtype=np.uint8), parts[4])
    N = np.array(m.block_cells)
    leaf = m.find_leaf_ix([int((x[d, 0] - m.c.x_global_min[d]) / m.c.dx_max_refinement[d]) for d in range(3)])
    lo = m.leaf_xmin()[leaf]
    dxc = (m.leaf_xmax()[leaf] - lo) / N
    c = np.floor((x[:, 0] - lo) / dxc).astype(int)
    cells = np.array([leaf * int(N.prod()) + c[0] + N[0] * (c[1] + N[1] * c[2])], dtype=np.int32)
    o = Oracle(cfg, m)
    o.set_background(*bg)
    o.add_particles(x, v, np.ones(1), np.zeros(1, dtype=np.uint8), cells)
    traj = []
    for _ in range(n_steps):
        rc, st, ret, fc = o.move(_capi.MOVER_RELATIVISTIC_BORIS, 1)
        assert rc == 0 and ret[0] == _capi.PARTICLE_MOTION_FINISHED
        pp = o.particles()
        traj.append((pp["x"][:, 0].copy(), pp["v"][:, 0].copy()))
    o.close()
    return traj


def test_momentum_rotation_matches_reference_gridless_boris():
    ref = tp.load_ref_gridless()
    if ref is None:
        pytest.skip("oracle/_ref/libref_gridless.so not built (needs /root/reference)")
    B = np.array([0.0, 0.0, 2.0e-5])
    v0 = np.array([0.6 * tp.CLIGHT, 0.1 * tp.CLIGHT, 0.2 * tp.CLIGHT])
    gamma = 1.0 / np.sqrt(1.0 - (v0 ** 2).sum() / tp.CLIGHT ** 2)
    fg = tp.QP * np.linalg.norm(B) / (2 * np.pi * tp.MP * gamma)
    dt = 0.02 / fg  # well below the sub-cycling limit 1/f_g: one Boris rotation per call
    traj = _single([1.0e7, 2.0e6, -3.0e6], v0, B, dt, 25)
    x = np.array([1.0e7, 2.0e6, -3.0e6])
    p = gamma * tp.MP * v0
    for k, (xo, vo) in enumerate(traj):
        ref.ref_boris_uniform(x.ctypes.data, p.ctypes.data, tp.QP, tp.MP, dt, B.ctypes.data, 1)
        g2 = np.sqrt(1.0 + (p ** 2).sum() / (tp.MP * tp.CLIGHT) ** 2)
        vref = p / (g2 * tp.MP)
        assert np.abs(vo - vref).max() <= 1e-12 * np.linalg.norm(vref), k   # same rotation, different code
        assert abs(np.linalg.norm(vo) - np.linalg.norm(v0)) <= 1e-13 * np.linalg.norm(v0)  # |p| conserved
    # positions: kick-drift vs drift-kick-drift differ by O(dt * dv) per step, both follow the same circle
    assert np.linalg.norm(traj[-1][0] - x) < 0.05 * np.linalg.norm(v0) * dt * len(traj)


def test_gyroradius_and_subcycling():
    B = np.array([0.0, 0.0, 3.0e-5])
    v0 = np.array([0.5 * tp.CLIGHT, 0.0, 0.0])
    gamma = 1.0 / np.sqrt(1.0 - (v0 ** 2).sum() / tp.CLIGHT ** 2)
    fg = tp.QP * np.linalg.norm(B) / (2 * np.pi * tp.MP * gamma)
    rg = gamma * tp.MP * np.linalg.norm(v0) / (tp.QP * np.linalg.norm(B))
    # dt = 3.5 gyro periods: the mover sub-cycles with 1/f_g ( = one full period per sub-step: the Boris rotation
    # angle is 2 atan(pi) per sub-step, so the speed is conserved and the particle stays within 2 r_g of the start )
    traj = _single([0.0, 0.0, 1.0e6], v0, B, 3.5 / fg, 4)
    for xo, vo in traj:
        assert abs(np.linalg.norm(vo) - np.linalg.norm(v0)) <= 1e-12 * np.linalg.norm(v0)
        assert np.linalg.norm(xo[:2]) <= 2.0 * rg * (1 + np.pi)  # bounded gyration, no secular drift
        assert abs(xo[2] - 1.0e6) < 1e-6


CASES = {
    "forward_linear_user": dict(backward=False, interp=_capi.CPLR_LINEAR, boundary=_capi.BOUNDARY_USER_FUNCTION, sphere=True),
    "backward_linear_user": dict(backward=True, interp=_capi.CPLR_LINEAR, boundary=_capi.BOUNDARY_USER_FUNCTION, sphere=True),
    "backward_constant_delete": dict(backward=True, interp=_capi.CPLR_CONSTANT, boundary=_capi.BOUNDARY_DELETE, sphere=True),
    "forward_linear_nosphere": dict(backward=False, interp=_capi.CPLR_LINEAR, boundary=_capi.BOUNDARY_USER_FUNCTION, sphere=False, dt=0.5),
    # config 1 geometry: cells grow with the distance from the planet (3 levels); the coupler stencil takes the AMR
    # branch of CellCentered::Linear::InitStencil (coarse-lattice stencil, 2x2x2 fine averages, blending)
    "amr_backward_linear_user": dict(backward=True, interp=_capi.CPLR_LINEAR, boundary=_capi.BOUNDARY_USER_FUNCTION, sphere=True, amr_levels=2,
                                     n_blocks=4, dt=0.1),
    "amr_forward_linear_ghost2": dict(backward=False, interp=_capi.CPLR_LINEAR, boundary=_capi.BOUNDARY_DELETE, sphere=True, amr_levels=2,
                                      n_blocks=4, dt=0.1, ghost_cells=(2, 2, 2)),
}


def test_boris_oracle_energy_in_pure_magnetic_field():
    """non-relativistic Boris: |v| is conserved exactly by the rotation when E = 0 and gravity is off"""
    m, cfg, parts, bg = tp.make_tp_case(n_particles=512, dt=0.01, sphere=False, boundary=_capi.BOUNDARY_DELETE, rigidity_gv=(0.001, 0.01))
    E, B = bg
    ora = tp.run_oracle_tp(m, cfg, parts, (np.zeros_like(E), B), mover=_capi.MOVER_BORIS)
    alive = ora["final_cell"] >= 0
    v0 = np.linalg.norm(parts[1], axis=0)[alive]
    v1 = np.linalg.norm(ora["particles"]["v"], axis=0)[alive]
    assert alive.sum() > 400 and np.abs(v1 - v0).max() <= 1e-13 * v0.max()


def test_oracle_threads_agree_and_exits_are_recorded():
    m, cfg, parts, bg = tp.make_tp_case(n_particles=2048, dt=0.3, **{k: v for k, v in CASES["backward_linear_user"].items()})
    a = tp.run_oracle_tp(m, cfg, parts, bg)
    b = tp.run_oracle_tp(m, cfg, parts, bg, n_threads=4)
    assert a["rc"] == 0 and a["lists"] == 0
    assert (a["final_cell"] == b["final_cell"]).all() and a["stats"] == b["stats"] and a["records"] == b["records"]
    st = a["stats"]
    assert st["n_left_domain"] == a["n_records"] > 0          # every exit (face or sphere) produced a record
    faces = [r[2] for r in a["records"]]
    assert _capi.EXIT_SPHERE in faces and any(f < 6 for f in faces)
    assert st["n_moved"] == 2048 and st["n_error"] == 0


def test_markidis2010_oracle_matches_the_closed_form():
    """PIC::Mover::Markidis2010 (pic_mover_boris.cpp:557-835) in uniform fields: eqs 22-23 evaluated with numpy"""
    Bu = np.array([1.0e-6, -2.0e-6, 2.0e-5])
    m, cfg, parts, bg = tp.make_tp_case(n_particles=1024, dt=0.01, sphere=False, boundary=_capi.BOUNDARY_DELETE, uniform_B=Bu,
                                        rigidity_gv=(0.001, 0.01))
    E = np.broadcast_to(np.array([2.0e-4, 1.0e-3, -3.0e-4]), bg[1].shape).copy()
    ora = tp.run_oracle_tp(m, cfg, parts, (E, bg[1]), mover=_capi.MOVER_MARKIDIS2010)
    assert ora["rc"] == 0 and ora["lists"] == 0
    alive = ora["final_cell"] >= 0
    assert alive.sum() > 900
    x0, v0 = parts[0][:, alive].T, parts[1][:, alive].T
    b1 = tp.QP * 0.01 / tp.MP
    b2 = 0.5 * b1
    vp = v0 + b1 * E[0]
    den = 1.0 / (1.0 + b2 * b2 * (Bu @ Bu))
    vf = den * (vp + b2 * np.cross(vp, Bu) + (b2 * b2 * (vp @ Bu))[:, None] * Bu)
    assert np.allclose(ora["particles"]["v"][:, alive].T, vf, rtol=1e-12)
    assert np.allclose(ora["particles"]["x"][:, alive].T, x0 + 0.01 * vf, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("mover", [_capi.MOVER_RELATIVISTIC_BORIS, _capi.MOVER_BORIS, _capi.MOVER_MARKIDIS2010])
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_parity(name, mover):
    kw = dict(CASES[name])
    dt = kw.pop("dt", 0.3)
    if mover == _capi.MOVER_MARKIDIS2010:
        dt *= 0.02
        kw["rigidity_gv"] = (0.001, 0.05)
        kw["backward"] = False          # the scheme has no backward-time mode
    if mover == _capi.MOVER_BORIS:
        dt *= 0.02   # single step without sub-cycling: keep the rotation angle moderate
        kw["rigidity_gv"] = (0.001, 0.05)  # non-relativistic protons
    m, cfg, parts, bg = tp.make_tp_case(n_particles=8192, dt=dt, seed=7, **kw)
    if mover == _capi.MOVER_BORIS:
        cfg.gravity_gm = 3.986004418e14
    ora = tp.run_oracle_tp(m, cfg, parts, bg, mover=mover)
    gpu = tp.run_gpu_tp(m, cfg, parts, bg, mover=mover)
    assert ora["rc"] == 0 and gpu["rc"] == 0
    n = parts[0].shape[1]
    mv = gpu["moved"]
    gx, gv, gc = np.empty((3, n)), np.empty((3, n)), np.empty(n, dtype=np.int64)
    gx[:, mv["ptrs"]], gv[:, mv["ptrs"]], gc[mv["ptrs"]] = mv["x"], mv["v"], mv["cells"]
    oc = ora["final_cell"].astype(np.int64)
    alive = oc >= 0
    assert (gc == oc).all()                                     # bit-exact block/cell assignment, same deletions
    assert (gx[:, alive] == ora["particles"]["x"][:, alive]).all()
    assert (gv[:, alive] == ora["particles"]["v"][:, alive]).all()
    assert gpu["stats"] == ora["stats"]
    assert gpu["n_records"] == ora["n_records"] and gpu["records"] == ora["records"]   # same faces, bit-equal x, v
    assert gpu["n_after"] == int(alive.sum())
    print(name, mover, ora["stats"], "records", ora["n_records"])
