"""Synthetic plasmas for tests and bench.py (data = "synthetic").

Config 2/3 ("ECSIM uniform periodic box"): box [0,n)^3 with dx = 1, species e- (q=-1, m=1) and
p+ (q=+1, m=1836) in normalised units (c = 1), `ppc` particles per cell PER SPECIES placed
uniformly at random inside each cell, Maxwellian v_th,e = 0.05, v_th,p = 0.05/sqrt(1836),
w_corr = 1, species weight such that omega_pe*dt = 0.5, B = (0, 0.04 + 0.004 sin(2 pi x/Lx), 0),
dt = 1.  The generator is counter based (Philox) keyed by (seed, cell id), so a particle's state
does not depend on the domain decomposition -- like the fast-wave fixture's pinned rnd_seed
(test/srcFastWave/main.cpp:111-345, :549).
"""
import numpy as np

V_TH_E = 0.05
MASS_RATIO = 1836.0


def species_tables(ppc, dt=1.0, dx=1.0):
    charge = (-1.0, 1.0)
    mass = (1.0, MASS_RATIO)
    # omega_pe^2 = 4 pi n q^2 / m,  n = ppc * weight / dx^3,  omega_pe * dt = 0.5
    wgt = (0.5 / dt) ** 2 / (4.0 * np.pi * ppc) * dx ** 3
    return charge, mass, (wgt, wgt)


def box_fields(mesh, E_amp=0.0, b_on_corners=False):
    """E_half on unique corners, B on unique centres (or on the corners for _PIC_FIELD_SOLVER_B_CORNER_BASED_).
    E_amp != 0 adds a smooth test field."""
    L = mesh.user_xmax - mesh.user_xmin
    xc = mesh.corner_x
    xb = mesh.corner_x if b_on_corners else mesh.center_x
    E = np.zeros((mesh.n_corners, 3))
    if E_amp != 0.0:
        ph = 2.0 * np.pi * (xc - mesh.user_xmin[None, :]) / L[None, :]
        E[:, 0] = E_amp * np.sin(ph[:, 1]) * np.cos(ph[:, 2])
        E[:, 1] = E_amp * np.sin(ph[:, 2]) * np.cos(ph[:, 0])
        E[:, 2] = E_amp * np.sin(ph[:, 0]) * np.cos(ph[:, 1])
    B = np.zeros((xb.shape[0], 3))
    B[:, 1] = 0.04 + 0.004 * np.sin(2.0 * np.pi * (xb[:, 0] - mesh.user_xmin[0]) / L[0])
    if E_amp != 0.0:  # make all B components and gradients non-trivial for parity tests
        ph = 2.0 * np.pi * (xb - mesh.user_xmin[None, :]) / L[None, :]
        B[:, 0] = 0.01 * np.cos(ph[:, 1])
        B[:, 2] = 0.02 * np.sin(ph[:, 1] + ph[:, 2])
    return E, B


def amr_sphere_box(n_blocks, block_cells=(8, 8, 8), ghost_cells=(1, 1, 1), radii=(12.0, 6.0), rank=0, n_ranks=1, decomp=None):
    """BASELINE config 4 geometry (scaled by the caller): open box [0,n)^3 of unit base cells, one more refinement level
    inside every sphere radii[l] (in base cells) about the centre; outer boundary = DELETE."""
    n_blocks = np.asarray(n_blocks, dtype=np.int64)
    N = np.asarray(block_cells, dtype=np.int64)
    hi = (n_blocks * N).astype(np.float64)
    ctr = 0.5 * hi

    def refine(level, lo, up):
        if level >= len(radii):
            return False
        near = np.clip(ctr, lo, up)  # closest point of the block to the centre
        return float(np.linalg.norm(near - ctr)) < radii[level]

    from . import mesh as meshmod
    return meshmod.build_mesh((0.0, 0.0, 0.0), tuple(hi), tuple(int(v) for v in n_blocks), block_cells, ghost_cells, periodic=False,
                              refine=refine, max_level=len(radii), rank=rank, n_ranks=n_ranks, decomp=decomp)


def maxwellian_amr(mesh, ppc_by_level, seed=100, drift=(0.02, 0.0, 0.0)):
    """"mixed ppc": ppc_by_level[l] particles per cell and species on the leaves of level l; w_corr keeps the
    number density uniform (cell volume 8^-l, fewer particles per cell)."""
    lev = mesh.leaf_level()
    parts = []
    for l, ppc in enumerate(ppc_by_level):
        leaves = np.nonzero(lev == l)[0]
        if len(leaves) == 0:
            continue
        x, v, w, sp, cells = maxwellian_box(mesh, ppc, seed=seed + 7919 * l, drift=drift, leaves=leaves)
        w *= (ppc_by_level[0] / ppc) / 8.0 ** l
        parts.append((x, v, w, sp, cells))
    return (np.concatenate([p[0] for p in parts], axis=1), np.concatenate([p[1] for p in parts], axis=1), np.concatenate([p[2] for p in parts]),
            np.concatenate([p[3] for p in parts]), np.concatenate([p[4] for p in parts]))


def maxwellian_box(mesh, ppc, seed=100, drift=(0.0, 0.0, 0.0), leaves=None, chunk_cells=1 << 15):
    """Particles for every cell of the real leaves (or `leaves`).  Returns x[3,n], v[3,n], w[n], species[n], cells[n]."""
    N = np.array(mesh.block_cells)
    C = mesh.cells_per_block
    if leaves is None:
        leaves = mesh.real_leaves()
    leaves = np.asarray(leaves, dtype=np.int64)
    lxmin = mesh.leaf_xmin()[leaves]
    lxmax = mesh.leaf_xmax()[leaves]
    ncell = len(leaves) * C
    npart = ncell * 2 * ppc
    x = np.empty((3, npart))
    v = np.empty((3, npart))
    species = np.empty(npart, dtype=np.uint8)
    cells = np.empty(npart, dtype=np.int32)
    vth = (V_TH_E, V_TH_E / np.sqrt(MASS_RATIO))
    # cell-major generation in chunks: cell c of leaf l owns particles [c*2ppc, (c+1)*2ppc)
    c_all = np.arange(ncell, dtype=np.int64)
    for c0 in range(0, ncell, chunk_cells):
        cc = c_all[c0:c0 + chunk_cells]
        li = cc // C
        ci = cc % C
        k = ci // (N[0] * N[1])
        j = (ci // N[0]) % N[1]
        i = ci % N[0]
        ijk = np.stack([i, j, k], axis=0).astype(np.float64)  # [3,nc]
        dxc = ((lxmax[li] - lxmin[li]) / N[None, :]).T  # [3,nc]
        lo = lxmin[li].T + ijk * dxc
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, c0]))
        nc = len(cc)
        u = rng.random((3, nc, 2 * ppc))
        # keep particles strictly inside their cell so that the upload key is consistent with x
        u = np.clip(u, 1e-9, 1.0 - 1e-9)
        g = rng.standard_normal((3, nc, 2 * ppc))
        sl = slice(c0 * 2 * ppc, (c0 + nc) * 2 * ppc)
        x[:, sl] = (lo[:, :, None] + u * dxc[:, :, None]).reshape(3, -1)
        sp = np.tile(np.repeat(np.arange(2, dtype=np.uint8), ppc), nc)
        species[sl] = sp
        vt = np.where(sp == 0, vth[0], vth[1])
        v[:, sl] = g.reshape(3, -1) * vt[None, :] + np.asarray(drift)[:, None]
        cells[sl] = np.repeat(leaves[li] * C + ci, 2 * ppc).astype(np.int32)
    w = np.ones(npart)
    return x, v, w, species, cells
