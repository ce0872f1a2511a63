"""Shared driver of the GPU-vs-oracle parity cases (used by tests/ and __graft_entry__.smoke()).

Tolerances (BASELINE.json north_star): particle-to-block/cell assignment and crossing counts
bit-exact over one step; positions, velocities, J and mass-matrix entries within 1e-10 relative.
x and v are compared element-wise; J and M relative to the largest magnitude of the array
(individual entries are sums with cancellation, the summation order on the GPU differs).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from amps_b200 import _capi, api, mesh as meshmod, workload  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

REL_TOL = 1e-10


def rel_elementwise(a, b):
    d = np.abs(a - b)
    s = np.maximum(np.abs(b), 1e-300)
    return float((d / s).max()) if d.size else 0.0


def rel_scaled(a, b):
    s = float(np.abs(b).max()) if b.size else 0.0
    if s == 0.0:
        return float(np.abs(a).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / s)


def make_case(n_cells=(16, 16, 16), ppc=8, seed=3, block_cells=(8, 8, 8), ghost_cells=(1, 1, 1), E_amp=0.01, dt=1.0, vscale=1.0,
              periodic=True, extra_capacity=0, boundary_mode=0, b_mode=0, amr_radii=None, ppc_by_level=None, exact_arithmetic=0,
              four_species=False):
    if amr_radii is not None:  # BASELINE config 4 (scaled): sphere-refined open box, mixed ppc, drifting Maxwellian
        nb = [n_cells[d] // block_cells[d] for d in range(3)]
        m = workload.amr_sphere_box(nb, block_cells, ghost_cells, radii=amr_radii)
        periodic = False
    elif periodic:
        m = meshmod.uniform_periodic_box(n_cells, block_cells, ghost_cells)
    else:
        nb = [n_cells[d] // block_cells[d] for d in range(3)]
        m = meshmod.build_mesh((0.0, 0.0, 0.0), tuple(float(c) for c in n_cells), nb, block_cells, ghost_cells, periodic=False)
    charge, mass, wgt = workload.species_tables(ppc, dt)
    if amr_radii is not None:
        x, v, w, sp, cells = workload.maxwellian_amr(m, ppc_by_level or (ppc,) * (len(amr_radii) + 1), seed=seed)
    else:
        x, v, w, sp, cells = workload.maxwellian_box(m, ppc, seed=seed)
        w[:] = 1.0
    v *= vscale
    if four_species:
        # e-, p+, a heavy ion and a NEUTRAL species (charge 0: pushed ballistically, deposits nothing, counts in energy and cfl):
        # more than two species take the separate diagnostics kernel of the deposit
        charge, mass, wgt = tuple(charge) + (1.0, 0.0), tuple(mass) + (16.0 * mass[1], 4.0 * mass[1]), tuple(wgt) + (wgt[0], 0.5 * wgt[0])
        r4 = np.random.default_rng(seed + 2).uniform(size=sp.shape)
        sp = np.where((sp == 1) & (r4 < 0.3), 2, np.where((sp == 1) & (r4 > 0.7), 3, sp)).astype(np.uint8)
    rng = np.random.default_rng(seed + 1)
    w = w * rng.uniform(0.5, 1.5, size=w.shape)  # exercise the individual weight correction
    cfg = api.make_config(block_cells, ghost_cells, charge, mass, wgt, dt, periodic=periodic, capacity=x.shape[1] + extra_capacity + 16,
                          boundary_mode=boundary_mode)
    cfg.exit_record_capacity = x.shape[1]
    cfg.b_mode = b_mode
    cfg.exact_arithmetic = exact_arithmetic
    E, B = workload.box_fields(m, E_amp=E_amp, b_on_corners=(b_mode == _capi.B_CORNER_BASED))
    Bcur = B * 1.01 + 0.001  # B_cur != B_prev so that a mix-up of the two shows
    return m, cfg, (x, v, w, sp, cells), (E, B, Bcur)


def run_oracle(m, cfg, parts, fields, n_threads=1, kind="parity"):
    x, v, w, sp, cells = parts
    E, B, Bcur = fields
    o = Oracle(cfg, m, kind)
    o.set_fields(E, B, Bcur)
    o.add_particles(x, v, w, sp, cells)
    rc, st, ret, fc = o.move(0, n_threads)
    pp = o.particles()
    nrec, recs = o.exit_records()
    lists_ok = o.check_lists()
    J, M, en, cfl = o.deposit(n_threads)
    o.close()
    return {"rc": rc, "stats": st, "ret": ret, "final_cell": fc, "particles": pp, "J": J, "M": M, "energy": en, "cfl": cfl, "lists": lists_ok,
            "records": sorted(recs), "n_records": nrec}


def run_gpu(m, cfg, parts, fields, ctx=None):
    """ctx: an existing context (of another mesh epoch) to re-use: the mesh is uploaded again and the context stays open"""
    x, v, w, sp, cells = parts
    E, B, Bcur = fields
    if ctx is None:
        g = api.Context(cfg, m)
    else:
        g = ctx
        g.mesh_upload(m)
    g.fields_upload(E, B, Bcur)
    g.particles_upload(x, v, w, sp, cells)
    n0 = g.particle_count()
    st = g.MoveParticles()
    moved = g.particles_download()  # slot i still holds the particle it held before the move
    nrec, recs = g.exit_records()
    g.sort()
    table = g.cell_table()
    srt = g.particles_download()
    en, cfl = g.UpdateJMassMatrix()
    J, M = g.JM_download()
    launches = g.launch_count()
    n_redo = g.last_move_redo()
    if ctx is None:
        g.close()
    return {"n_redo": n_redo, "n0": n0, "stats": st, "moved": moved, "sorted": srt, "table": table, "J": J, "M": M, "energy": en, "cfl": cfl, "launches": launches,
            "records": sorted(recs), "n_records": nrec}


def compare(m, parts, ora, gpu):
    n = parts[0].shape[1]
    res = {"n": n}
    # GPU arrays by ptr
    mv = gpu["moved"]
    ptr = mv["ptrs"]
    gx = np.empty((3, n)), np.empty((3, n))
    gcell = np.full(n, -2, dtype=np.int64)
    gx[0][:, ptr] = mv["x"]
    gx[1][:, ptr] = mv["v"]
    gcell[ptr] = mv["cells"]
    ocell = ora["final_cell"].astype(np.int64)
    alive = ocell >= 0
    res["n_alive_oracle"] = int(alive.sum())
    res["cell_mismatch"] = int((gcell != ocell).sum())
    ox, ov = ora["particles"]["x"], ora["particles"]["v"]
    res["max_rel_x"] = rel_elementwise(gx[0][:, alive], ox[:, alive])
    res["max_rel_v"] = rel_elementwise(gx[1][:, alive], ov[:, alive])
    # the same per particle against the length of the vector: a component that passes through zero has no relative accuracy of
    # its own (the fast mover's contracted arithmetic differs from the reference by ~1e-16 of |v|)
    def relnorm(a, b):
        nb = np.sqrt((b * b).sum(axis=0))
        return float((np.abs(a - b).max(axis=0) / np.maximum(nb, 1e-300)).max()) if b.size else 0.0
    res["max_relnorm_x"] = relnorm(gx[0][:, alive], ox[:, alive])
    res["max_relnorm_v"] = relnorm(gx[1][:, alive], ov[:, alive])
    res["bit_mismatch_xv"] = int((gx[0][:, alive] != ox[:, alive]).sum() + (gx[1][:, alive] != ov[:, alive]).sum())
    res["stats_equal"] = all(ora["stats"][k] == gpu["stats"][k] for k in ora["stats"])
    res["stats_gpu"], res["stats_oracle"] = gpu["stats"], ora["stats"]
    # sorted layout
    s = gpu["sorted"]
    keys = s["cells"].astype(np.int64)
    tab = gpu["table"]
    res["sorted_ok"] = bool((np.diff(keys) >= 0).all()) and len(keys) == int(alive.sum())
    cnt = np.bincount(keys, minlength=m.n_cells) if len(keys) else np.zeros(m.n_cells, dtype=np.int64)
    res["table_ok"] = bool((np.diff(tab) == cnt).all() and tab[0] == 0 and tab[-1] == len(keys))
    # the sort is a permutation of the moved particles
    order = np.argsort(s["ptrs"])
    res["perm_ok"] = bool((np.sort(s["ptrs"]) == np.nonzero(alive)[0]).all()) and bool((s["x"][:, order] == gx[0][:, alive]).all()) and bool(
        (s["cells"][order] == gcell[alive]).all())
    res["max_rel_J"] = rel_scaled(gpu["J"], ora["J"])
    res["max_rel_M"] = rel_scaled(gpu["M"], ora["M"])
    res["rel_energy"] = abs(gpu["energy"] - ora["energy"]) / max(abs(ora["energy"]), 1e-300)
    res["rel_cfl"] = max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(gpu["cfl"], ora["cfl"]))
    res["oracle_lists"] = ora["lists"]
    res["n_records"] = ora["n_records"]
    res["records_equal"] = (gpu["n_records"] == ora["n_records"]) and (gpu["records"] == ora["records"])
    res["gpu_launches"] = gpu["launches"]
    res["n_redo"] = gpu["n_redo"]
    return res


def run_parity_case(**kw):
    steps = kw.pop("steps", 1)
    assert steps == 1
    m, cfg, parts, fields = make_case(**kw)
    ora = run_oracle(m, cfg, parts, fields)
    gpu = run_gpu(m, cfg, parts, fields)
    return compare(m, parts, ora, gpu)
