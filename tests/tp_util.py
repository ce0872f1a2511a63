"""Test-particle fixtures: protons in an Earth-like dipole (+ a weak convection E) tabulated on the centre nodes of an
open box, like srcEarth/main_lib.cpp:686-813 tabulates T96 and srcMoverTest/main_lib.cpp:127-225 its dipole."""
import ctypes as C
import os

import numpy as np

from amps_b200 import _capi, api, mesh as meshmod
from oracle.oracle_py import Oracle

from amps_b200.workload import B0, CLIGHT, MP, QP, RE, dipole, gc_gradB, gca_var15  # noqa: F401
from amps_b200.workload import background_analytic as _bg_analytic


def make_tp_case(n_particles=4096, half_width_re=8.0, n_blocks=8, block_cells=(4, 4, 4), seed=1, dt=0.05, backward=False,
                 interp=_capi.CPLR_LINEAR, boundary=_capi.BOUNDARY_USER_FUNCTION, sphere=True, uniform_B=None, rigidity_gv=(0.5, 20.0),
                 amr_levels=0, ghost_cells=(1, 1, 1)):
    L = half_width_re * RE
    refine = None
    if amr_levels > 0:
        # cell size grows with the distance from the planet, like dx = 0.25 R_E (r/R_E) of input/earth-cutoff-rigidity.input
        def refine(level, lo, hi):
            near = np.clip(np.zeros(3), lo, hi)
            r = float(np.linalg.norm(near)) / RE
            return r < half_width_re * (0.45, 0.12, 0.03)[level]  # nested so that neighbours differ by one level at most
    m = meshmod.build_mesh((-L, -L, -L), (L, L, L), (n_blocks,) * 3, block_cells, ghost_cells, periodic=False, refine=refine, max_level=amr_levels)
    xc = m.center_x
    if uniform_B is None:
        r = np.sqrt((xc ** 2).sum(1))
        B = dipole(np.where(r[:, None] < 0.5 * RE, xc + 0.5 * RE, xc))
        vbg = np.array([-4.0e5, 0.0, 0.0])
        E = -np.cross(np.broadcast_to(vbg, B.shape), B)
    else:
        B = np.broadcast_to(np.asarray(uniform_B, dtype=np.float64), xc.shape).copy()
        E = np.zeros_like(B)
    rng = np.random.default_rng(seed)
    # launch points: shell 1.2..6 RE, isotropic directions, log-uniform rigidity
    u = rng.standard_normal((n_particles, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    rad = RE * rng.uniform(1.2, min(6.0, half_width_re - 0.5), n_particles)
    x = (u * rad[:, None]).T.copy()
    d = rng.standard_normal((n_particles, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    R = np.exp(rng.uniform(np.log(rigidity_gv[0]), np.log(rigidity_gv[1]), n_particles)) * 1e9  # volts
    p = R * QP / CLIGHT
    gamma = np.sqrt(1.0 + (p / (MP * CLIGHT)) ** 2)
    speed = p / (gamma * MP)
    v = (d * speed[:, None]).T.copy()
    sp = np.zeros(n_particles, dtype=np.uint8)
    # cells from positions: leaf by the findTreeNode lattice, then the cell inside the leaf
    N = np.array(block_cells)
    gmin = np.array([m.c.x_global_min[d] for d in range(3)])
    dxr = np.array([m.c.dx_max_refinement[d] for d in range(3)])
    lat = np.floor((x.T - gmin) / dxr).astype(np.int64)
    leaf = np.array([m.find_leaf_ix([int(b[0]), int(b[1]), int(b[2])]) for b in lat])
    lo, hi = m.leaf_xmin()[leaf], m.leaf_xmax()[leaf]
    cidx = np.floor((x.T - lo) / ((hi - lo) / N)).astype(np.int64)
    cidx = np.clip(cidx, 0, N - 1)
    cells = (leaf * int(N.prod()) + cidx[:, 0] + N[0] * (cidx[:, 1] + N[1] * cidx[:, 2])).astype(np.int32)
    cfg = api.make_config(block_cells, ghost_cells, (QP,), (MP,), (1.0,), dt, periodic=False, capacity=n_particles + 16, boundary_mode=boundary)
    cfg.time_step_mode = _capi.DT_SPECIES_GLOBAL
    cfg.coupler_interpolation = interp
    cfg.backward_time_integration = 1 if backward else 0
    cfg.speed_of_light = CLIGHT
    cfg.internal_sphere_radius = RE if sphere else 0.0
    cfg.exit_record_capacity = n_particles
    return m, cfg, (x, v, np.ones(n_particles), sp, cells), (E, B)


def run_oracle_tp(m, cfg, parts, bg, mover=_capi.MOVER_RELATIVISTIC_BORIS, n_threads=1):
    o = Oracle(cfg, m)
    o.set_background(*bg)
    o.add_particles(*parts)
    rc, st, ret, fc = o.move(mover, n_threads)
    pp = o.particles()
    nrec, recs = o.exit_records()
    lists = o.check_lists()
    o.close()
    return {"rc": rc, "stats": st, "ret": ret, "final_cell": fc, "particles": pp, "records": sorted(recs), "n_records": nrec, "lists": lists}


def run_gpu_tp(m, cfg, parts, bg, mover=_capi.MOVER_RELATIVISTIC_BORIS):
    g = api.Context(cfg, m)
    g.background_upload(*bg)
    g.particles_upload(*parts)
    try:
        st = g.MoveParticles(mover)
        rc = 0
    except api.AmpsGpuError:
        st, rc = None, 5
    moved = g.particles_download()
    nrec, recs = g.exit_records()
    g.sort()
    n_after = g.particle_count()
    g.close()
    return {"rc": rc, "stats": st, "moved": moved, "records": sorted(recs), "n_records": nrec, "n_after": n_after}


def load_ref_gridless():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_gridless.so")
    if not os.path.exists(p):
        return None
    lib = C.CDLL(p)
    lib.ref_boris_uniform.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int]
    lib.ref_boris_dipole.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    return lib


# ---- relativistic guiding-centre fixtures (config 5, srcMoverTest) -----------------------------------------
def make_gca_case(n_particles=4096, seed=2, dt=0.02, interp=_capi.CPLR_LINEAR, sphere=True, uniform_B=None, E_uniform=None,
                  rigidity_gv=(0.001, 0.05), convection=True, **kw):
    m, cfg, parts, _ = make_tp_case(n_particles=n_particles, seed=seed, dt=dt, interp=interp, boundary=_capi.BOUNDARY_DELETE, sphere=sphere,
                                    uniform_B=uniform_B, rigidity_gv=rigidity_gv, **kw)
    E, B = _bg_analytic(m.center_x, uniform_B=uniform_B, E_uniform=E_uniform, convection=convection)
    var15 = gca_var15(m.center_x, 1.0e3, uniform_B=uniform_B, E_uniform=E_uniform, convection=convection)
    cfg.carry_magnetic_moment = 1
    return m, cfg, parts, (E, B), var15


def run_oracle_gca(m, cfg, parts, bg, var15, n_threads=1):
    o = Oracle(cfg, m)
    o.set_background(*bg)
    o.set_background_gca(var15)
    o.add_particles(*parts)
    mu = o.magnetic_moment_init()
    rc, st, ret, fc = o.move(_capi.MOVER_RELATIVISTIC_GCA, n_threads)
    pp = o.particles()
    nrec, recs = o.exit_records()
    lists = o.check_lists()
    o.close()
    return {"rc": rc, "stats": st, "ret": ret, "final_cell": fc, "particles": pp, "records": sorted(recs), "n_records": nrec, "lists": lists, "mu": mu}


def run_gpu_gca(m, cfg, parts, bg, var15):
    g = api.Context(cfg, m)
    g.background_upload(*bg)
    g.background_upload_gca(var15)
    g.particles_upload(*parts)
    g.InitiateMagneticMoment()
    mu0 = g.magnetic_moment_download()
    ptr0 = g.particles_download()["ptrs"]
    st = g.MoveParticles(_capi.MOVER_RELATIVISTIC_GCA)
    moved = g.particles_download()
    nrec, recs = g.exit_records()
    g.sort()
    n_after = g.particle_count()
    srt = g.particles_download()
    mu1 = g.magnetic_moment_download()
    g.close()
    mu_by_ptr = np.empty(len(ptr0))
    mu_by_ptr[ptr0] = mu0
    return {"stats": st, "moved": moved, "records": sorted(recs), "n_records": nrec, "n_after": n_after, "mu": mu_by_ptr, "sorted": srt,
            "mu_sorted": mu1}


# ---- non-relativistic guiding centre (pic_mover_guiding_center.cpp) ----------------------------------------
def make_gc_case(n_particles=4096, seed=3, dt=0.01, interp=_capi.CPLR_LINEAR, sphere=True, uniform_B=None, E_uniform=None,
                 rigidity_gv=(0.001, 0.02), convection=True, ideal_mhd=1, **kw):
    m, cfg, parts, _ = make_tp_case(n_particles=n_particles, seed=seed, dt=dt, interp=interp, boundary=_capi.BOUNDARY_DELETE, sphere=sphere,
                                    uniform_B=uniform_B, rigidity_gv=rigidity_gv, **kw)
    E, B = _bg_analytic(m.center_x, uniform_B=uniform_B, E_uniform=E_uniform, convection=convection)
    gradB = gc_gradB(m.center_x, 1.0e3, uniform_B=uniform_B, E_uniform=E_uniform, convection=convection)
    cfg.carry_magnetic_moment = 1
    cfg.ideal_mhd = ideal_mhd
    return m, cfg, parts, (E, B), gradB


def run_oracle_gc(m, cfg, parts, bg, gradB, mover, pre_init, n_threads=1):
    o = Oracle(cfg, m)
    o.set_background(*bg)
    o.set_background_gradB(gradB)
    o.add_particles(*parts)
    mu0 = o.magnetic_moment_init(mover) if pre_init else None
    v0 = o.particles()["v"] if pre_init else None
    rc, st, ret, fc = o.move(mover, n_threads)
    pp = o.particles()
    mu1, flag = o.magnetic_moment()
    nrec, recs = o.exit_records()
    lists = o.check_lists()
    o.close()
    return {"rc": rc, "stats": st, "ret": ret, "final_cell": fc, "particles": pp, "records": sorted(recs), "n_records": nrec, "lists": lists,
            "mu0": mu0, "v0": v0, "mu": mu1, "flag": flag}


def run_gpu_gc(m, cfg, parts, bg, gradB, mover, pre_init):
    g = api.Context(cfg, m)
    g.background_upload(*bg)
    g.background_upload_gradB(gradB)
    g.particles_upload(*parts)
    n = parts[0].shape[1]
    mu0 = v0 = None
    if pre_init:
        g.InitiateMagneticMoment(mover)
        d = g.particles_download()
        mu0, v0 = np.empty(n), np.empty((3, n))
        mu0[d["ptrs"]] = g.magnetic_moment_download()
        v0[:, d["ptrs"]] = d["v"]
    st = g.MoveParticles(mover, raise_on_particle_error=False)
    moved = g.particles_download()
    mu_dev = g.magnetic_moment_download()
    nrec, recs = g.exit_records()
    g.sort()
    n_after = g.particle_count()
    g.close()
    mu1 = np.empty(n)
    mu1[moved["ptrs"]] = mu_dev
    flag = np.empty(n, dtype=np.uint8)
    flag[moved["ptrs"]] = (moved["species"] >> 6) & 1
    return {"stats": st, "moved": moved, "records": sorted(recs), "n_records": nrec, "n_after": n_after, "mu0": mu0, "v0": v0, "mu": mu1, "flag": flag}
