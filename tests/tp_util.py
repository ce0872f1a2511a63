"""Test-particle fixtures: protons in an Earth-like dipole (+ a weak convection E) tabulated on the centre nodes of an
open box, like srcEarth/main_lib.cpp:686-813 tabulates T96 and srcMoverTest/main_lib.cpp:127-225 its dipole."""
import ctypes as C
import os

import numpy as np

from amps_b200 import _capi, api, mesh as meshmod
from oracle.oracle_py import Oracle

RE = 6.371e6
QP, MP, CLIGHT = 1.602176634e-19, 1.67262192369e-27, 299792458.0
B0 = 3.1e-5


def dipole(x):
    r2 = (x ** 2).sum(1)
    r5 = r2 ** 2.5
    k = -B0 * RE ** 3
    B = np.empty_like(x)
    B[:, 0] = k * 3.0 * x[:, 2] * x[:, 0] / r5
    B[:, 1] = k * 3.0 * x[:, 2] * x[:, 1] / r5
    B[:, 2] = k * (3.0 * x[:, 2] ** 2 - r2) / r5
    return B


def make_tp_case(n_particles=4096, half_width_re=8.0, n_blocks=8, block_cells=(4, 4, 4), seed=1, dt=0.05, backward=False,
                 interp=_capi.CPLR_LINEAR, boundary=_capi.BOUNDARY_USER_FUNCTION, sphere=True, uniform_B=None, rigidity_gv=(0.5, 20.0)):
    L = half_width_re * RE
    m = meshmod.build_mesh((-L, -L, -L), (L, L, L), (n_blocks,) * 3, block_cells, (1, 1, 1), periodic=False)
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
    # cells from positions (uniform open box)
    N = np.array(block_cells)
    dx_block = 2 * L / n_blocks
    bidx = np.floor((x.T + L) / dx_block).astype(np.int64)
    leaf = np.array([m.find_leaf_ix([int(b[0] * 4096 + 1), int(b[1] * 4096 + 1), int(b[2] * 4096 + 1)]) for b in bidx])
    lo = m.leaf_xmin()[leaf]
    cidx = np.floor((x.T - lo) / (dx_block / N)).astype(np.int64)
    cidx = np.minimum(cidx, N - 1)
    cells = (leaf * int(N.prod()) + cidx[:, 0] + N[0] * (cidx[:, 1] + N[1] * cidx[:, 2])).astype(np.int32)
    cfg = api.make_config(block_cells, (1, 1, 1), (QP,), (MP,), (1.0,), dt, periodic=False, capacity=n_particles + 16, boundary_mode=boundary)
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
