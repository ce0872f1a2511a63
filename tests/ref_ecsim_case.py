"""The fast-wave case of the reference's own PIC core (oracle/_ref/libref_pic.so, built by oracle/ref_pic/build_ref_pic.sh from the
reference tree) next to the same case expressed for this repo: mesh, unique-node fields, particles, configuration.

The reference is initialised ONCE per process (its state is global), driven through one ECSIM particle phase
   UpdateJMassMatrix (initial positions)  ->  MoveParticles + exchanges  ->  UpdateJMassMatrix
and everything it produced is kept for the comparisons in tests/test_reference_ecsim.py."""
import functools

import numpy as np

from amps_b200 import api, mesh as meshmod
from oracle.ref_pic import ref_pic


def _wrap_index(pos, xmin, n_cells, corner):
    """lattice index of node positions wrapped into the periodic box (dx = 1 in the fast-wave case)"""
    t = pos - xmin - (0.0 if corner else 0.5)
    i = np.rint(t).astype(np.int64)
    assert np.abs(t - i).max() < 1e-9
    return np.mod(i, n_cells)


@functools.lru_cache(maxsize=1)
def case(keep_every=1, prepare=None, do_field=True):
    """prepare(r, p0, info): called once the fields and weights are set, before the first UpdateJMassMatrix (the gyrokinetic variant
    marks its guiding-centre species and sets mu / v_normal there); info["E_cur"] is the current E it must put on the corners itself.
    do_field=False leaves the field half out (it overwrites E and swaps the two B slots)"""
    r = ref_pic.RefPic()
    if keep_every > 1:
        r.thin(keep_every)
    N, g = r.N, r.g
    real = np.nonzero(r.ghost == 0)[0]
    xmin = r.bxmin[real].min(axis=0)
    xmax = r.bxmax[real].max(axis=0)
    n_cells = np.rint(xmax - xmin).astype(np.int64)  # dx = 1
    assert tuple(n_cells) == (32, 16, 8)
    m = meshmod.uniform_periodic_box(tuple(int(c) for c in n_cells), N, g, dx=1.0, origin=tuple(xmin))

    def uid_map(node_x, corner):
        idx = _wrap_index(node_x, xmin, n_cells, corner)
        key = idx[:, 0] + n_cells[0] * (idx[:, 1] + n_cells[1] * idx[:, 2])
        table = np.full(int(np.prod(n_cells)), -1, dtype=np.int64)
        table[key] = np.arange(len(key))
        assert (table >= 0).all()  # every physical node of the periodic lattice has one unique id here
        return table

    ctab, ztab = uid_map(m.corner_x, True), uid_map(m.center_x, False)

    def ref_to_uid(pos, tab, corner):
        idx = _wrap_index(pos.reshape(-1, 3), xmin, n_cells, corner)
        return tab[idx[:, 0] + n_cells[0] * (idx[:, 1] + n_cells[1] * idx[:, 2])].reshape(pos.shape[:-1])

    cu = ref_to_uid(r.corner_positions(), ctab, True)   # [block][k][j][i] -> unique corner
    zu = ref_to_uid(r.center_positions(), ztab, False)

    # fields on the unique nodes (any smooth numbers: both sides get the same doubles), copied to every copy of a node the
    # reference holds (ghost layers, periodic ghost blocks)
    def smooth(x, amp, ph):
        k = 2 * np.pi / (xmax - xmin)
        return np.stack([amp * (1.0 + 0.5 * np.sin(k[0] * x[:, 0] + ph) * np.cos(k[1] * x[:, 1] - 0.3 * d) + 0.25 * np.sin(k[2] * x[:, 2] + d))
                         for d in range(3)], axis=1)
    E_u = smooth(m.corner_x, 0.01, 0.4) - 0.008
    Bp_u = smooth(m.center_x, 0.04, 1.1)
    Bc_u = smooth(m.center_x, 0.041, 0.7) + 0.001
    r.set_corner(1, E_u[cu])      # E at the half step (what Lapenta2017 interpolates)
    r.set_center(1, Bp_u[zu])     # B previous (the mover)
    r.set_center(0, Bc_u[zu])     # B current (ProcessCell)

    p0 = r.particles()
    rng = np.random.default_rng(7)
    w = rng.uniform(0.5, 1.5, size=p0["w"].shape)
    r.set_weight_correction(p0["ptr"], w)
    p0["w"] = w

    extra = prepare(r, p0, {"cu": cu, "zu": zu, "mesh": m, "smooth": smooth}) if prepare is not None else None

    # reference: deposit, move, deposit
    e0 = r.update_JM()
    J0, M0 = r.corner(2), r.corner(3)
    r.move()
    p1 = r.particles()
    e1 = r.update_JM()
    J1, M1 = r.corner(2), r.corner(3)

    # ---- the field half of the next step (ECSIM::TimeStep): E^n, B^n, J, M -> E^{n+theta}, E^{n+1}, B^{n+1} ----
    import ctypes as C

    def corner_u(arr):
        out = np.full((m.n_corners, arr.shape[-1]), np.nan)
        for b in real:
            sl = (b, slice(g[2], g[2] + N[2] + 1), slice(g[1], g[1] + N[1] + 1), slice(g[0], g[0] + N[0] + 1))
            out[cu[sl].reshape(-1)] = arr[sl].reshape(-1, arr.shape[-1])
        return out

    def center_u(arr):
        out = np.full((m.n_centers, 3), np.nan)
        for b in real:
            sl = (b, slice(g[2], g[2] + N[2]), slice(g[1], g[1] + N[1]), slice(g[0], g[0] + N[0]))
            out[zu[sl].reshape(-1)] = arr[sl].reshape(-1, 3)
        return out
    field = {}
    if do_field:
        Efld = smooth(m.corner_x, 0.02, 2.1) - 0.015
        r.set_corner(0, Efld[cu])  # E^n (the initial condition of fast-wave is E = 0: give the operator something to act on)
        field = {"E": corner_u(r.corner(0)), "B": center_u(r.center(0)), "theta": 0.5}
        # the field getters of the guiding-centre movers in ECSIM mode (ECSIM::GetElectricField / GetMagneticField / GetMagneticFieldGradient),
        # sampled before the field step touches E and swaps the B slots: uniform points in the real blocks, a few of them on block faces
        if hasattr(r.lib, "ref_pic_ecsim_fields"):
            rg = np.random.default_rng(11)
            nP = 6000
            gblk = rg.choice(real, nP).astype(np.int32)
            u = rg.uniform(0.0, 1.0, (nP, 3))
            u[:200, 0] = 0.0                       # on the lower x face of the block
            u[200:400, 1] = 1.0                    # on the upper y face (the corner stencil snaps these)
            u[400:600] = np.round(u[400:600] * 8) / 8  # on cell faces / corners (16 x 8 x 4 cells: multiples of 1/8 hit faces in every direction)
            gx = r.bxmin[gblk] + u * (r.bxmax[gblk] - r.bxmin[gblk])
            gE, gB, gG = np.zeros((nP, 3)), np.zeros((nP, 3)), np.zeros((nP, 9))
            with ref_pic.quiet():
                r.lib.ref_pic_ecsim_fields(C.c_long(nP), ref_pic._p(gx), ref_pic._p(gblk), ref_pic._p(gE), ref_pic._p(gB), ref_pic._p(gG))
            field["getters"] = {"x": gx, "block": gblk, "E": gE, "B": gB, "gradB": gG, "E_u": Efld, "B_u": Bc_u}
        rel_res = C.c_double()
        with ref_pic.quiet():
            field["iterations"] = r.lib.ref_pic_field_step(C.c_double(1e-12), 400, C.byref(rel_res))
        field["rel_residual"] = float(rel_res.value)
        r.lib.ref_pic_solver_rows.restype = C.c_long
        nrow = r.lib.ref_pic_solver_rows(0, None, None, None, None)
        blk, ijk, iv, rr = np.zeros(nrow, dtype=np.int32), np.zeros((nrow, 3), dtype=np.int32), np.zeros(nrow, dtype=np.int32), np.zeros(nrow)
        r.lib.ref_pic_solver_rows(nrow, ref_pic._p(blk), ref_pic._p(ijk), ref_pic._p(iv), ref_pic._p(rr))
        ru = cu[blk, ijk[:, 2] + g[2], ijk[:, 1] + g[1], ijk[:, 0] + g[0]]
        field["rhs"] = np.zeros((m.n_corners, 3))
        field["rhs"][ru, iv] = rr
        xin = np.random.default_rng(5).standard_normal((m.n_corners, 3))
        vin, vout = np.ascontiguousarray(xin[ru, iv]), np.zeros(nrow)
        r.lib.ref_pic_matvec(ref_pic._p(vin), ref_pic._p(vout), int(nrow))
        field["matvec_in"], field["matvec_out"] = xin, np.zeros((m.n_corners, 3))
        field["matvec_out"][ru, iv] = vout
        field["E_half"], field["E_new"], field["B_new"] = corner_u(r.corner(1)), corner_u(r.corner(0)), center_u(r.center(0))

        def stencil(kind, p, q):
            ijk3, aa = np.zeros((200, 3), dtype=np.int32), np.zeros(200)
            n = r.lib.ref_pic_stencil(kind, p, q, 200, ref_pic._p(ijk3), ref_pic._p(aa))
            T = np.zeros((3, 3, 3))
            for t in range(n):
                T[ijk3[t, 0] + 1, ijk3[t, 1] + 1, ijk3[t, 2] + 1] += aa[t]
            return T
        field["laplacian"] = [stencil(0, p, 0) for p in range(3)]
        field["graddiv"] = [[stencil(1, p, q) for q in range(3)] for p in range(3)]

    # particles in this repo's numbering: leaf by block position, same cell formula i + Nx (j + Ny k)
    lx = m.leaf_xmin()
    def leaf_of_block(b):
        d = np.abs(lx - r.bxmin[b]).max(axis=1)
        k = int(np.argmin(d))
        assert d[k] == 0.0
        return k
    b2l = np.array([leaf_of_block(b) if r.ghost[b] == 0 else -1 for b in range(r.n_blocks)])
    if "getters" in field:
        field["getters"]["leaf"] = b2l[field["getters"]["block"]].astype(np.int32)
    C = m.cells_per_block
    assert (b2l[p0["block"]] >= 0).all() and (b2l[p1["block"]] >= 0).all()
    cells0 = (b2l[p0["block"]] * C + p0["cell"]).astype(np.int32)
    # after the move, by ParticleBuffer slot
    order = np.argsort(p1["ptr"])
    pos = np.searchsorted(p1["ptr"][order], p0["ptr"])
    assert (p1["ptr"][order][pos] == p0["ptr"]).all()  # periodic box: nobody is deleted
    sel = order[pos]
    after = {"x": p1["x"][:, sel], "v": p1["v"][:, sel], "cells": (b2l[p1["block"][sel]] * C + p1["cell"][sel]).astype(np.int64)}

    cfg = api.make_config(N, g, tuple(r.charge), tuple(r.mass), tuple(r.weight), r.dt, periodic=True, capacity=p0["x"].shape[1] + 16,
                          B_conv=r.B_conv, length_conv=r.length_conv, light_speed=r.light_speed)

    # the reference's J, M per unique corner: every copy of a corner a real block holds (its own corners, i in 0..N) must agree
    def to_unique(arr):
        out = np.full((m.n_corners, arr.shape[-1]), np.nan)
        spread = 0.0
        for b in real:
            a = arr[b, g[2]:g[2] + N[2] + 1, g[1]:g[1] + N[1] + 1, g[0]:g[0] + N[0] + 1].reshape(-1, arr.shape[-1])
            u = cu[b, g[2]:g[2] + N[2] + 1, g[1]:g[1] + N[1] + 1, g[0]:g[0] + N[0] + 1].reshape(-1)
            have = ~np.isnan(out[u, 0])
            if have.any():
                spread = max(spread, float(np.abs(out[u[have]] - a[have]).max()))
            out[u] = a
        return out, spread
    ref = {"J0": to_unique(J0), "M0": to_unique(M0), "J1": to_unique(J1), "M1": to_unique(M1), "energy0": e0, "energy1": e1, "after": after}
    touched = ~np.isnan(ref["J0"][0][:, 0])
    ref["field"] = field
    return {"ref": ref, "mesh": m, "cfg": cfg, "parts": (p0["x"], p0["v"], p0["w"], p0["species"].astype(np.uint8), cells0),
            "fields": (E_u, Bp_u, Bc_u), "touched": touched, "refpic": r, "extra": extra, "ptr0": p0["ptr"], "ptr1_in_order0": p1["ptr"][sel],
            "maps": {"cu": cu, "zu": zu, "b2l": b2l, "real": real, "to_unique": to_unique}}
