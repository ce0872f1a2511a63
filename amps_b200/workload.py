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


def amr_sphere_box(n_blocks, block_cells=(8, 8, 8), ghost_cells=(1, 1, 1), radii=(12.0, 6.0), rank=0, n_ranks=1, decomp=None, leaf_weight=None):
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
                              refine=refine, max_level=len(radii), rank=rank, n_ranks=n_ranks, decomp=decomp, leaf_weight=leaf_weight)


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


# ---- test-particle workloads (BASELINE configs[0] and [4]): protons in an Earth-like dipole (+ a weak convection E) tabulated on the
# centre nodes of an open box, like srcEarth/main_lib.cpp:686-813 tabulates T96 and srcMoverTest/main_lib.cpp:127-225 its dipole
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


def background_analytic(x, uniform_B=None, E_uniform=None, convection=True):
    if uniform_B is None:
        r = np.sqrt((x ** 2).sum(1))
        B = dipole(np.where(r[:, None] < 0.5 * RE, x + 0.5 * RE, x))
        vbg = np.array([-4.0e5 if convection else 0.0, 0.0, 0.0])
        E = -np.cross(np.broadcast_to(vbg, B.shape), B)
    else:
        B = np.broadcast_to(np.asarray(uniform_B, dtype=np.float64), x.shape).copy()
        E = np.broadcast_to(np.asarray((0.0, 0.0, 0.0) if E_uniform is None else E_uniform, dtype=np.float64), x.shape).copy()
    return E, B


def gca_var15(x, h, **kw):
    """b.grad(b), vE.grad(b), b.grad(vE), vE.grad(vE), grad(kappa*B) by central differences of the analytic field
    (the reference tabulates the same five vectors from its data file, pic_datafile.cpp:1164-1340)"""
    def derived(xx):
        E, B = background_analytic(xx, **kw)
        Bn = np.sqrt((B ** 2).sum(1))
        b = B / Bn[:, None]
        vE = np.cross(E, B) / (Bn ** 2)[:, None]
        kappa = 1.0 / np.sqrt(1.0 - (vE ** 2).sum(1) / CLIGHT ** 2)
        return b, vE, kappa * Bn

    b, vE, _ = derived(x)
    grad_b = np.empty((x.shape[0], 3, 3))  # [n][j][i] = d_j b_i
    grad_vE = np.empty((x.shape[0], 3, 3))
    grad_kB = np.empty((x.shape[0], 3))
    for j in range(3):
        e = np.zeros(3)
        e[j] = h
        bp, vp, kp = derived(x + e)
        bm, vm, km = derived(x - e)
        grad_b[:, j, :] = (bp - bm) / (2 * h)
        grad_vE[:, j, :] = (vp - vm) / (2 * h)
        grad_kB[:, j] = (kp - km) / (2 * h)
    out = np.empty((x.shape[0], 15))
    out[:, 0:3] = np.einsum("nj,nji->ni", b, grad_b)
    out[:, 3:6] = np.einsum("nj,nji->ni", vE, grad_b)
    out[:, 6:9] = np.einsum("nj,nji->ni", b, grad_vE)
    out[:, 9:12] = np.einsum("nj,nji->ni", vE, grad_vE)
    out[:, 12:15] = grad_kB
    return out


def gc_gradB(x, h, **kw):
    """gradB[3*i+j] = d B_i / d x_j by central differences of the analytic field (layout of pic.h:8439-8442)"""
    out = np.empty((x.shape[0], 9))
    for j in range(3):
        e = np.zeros(3)
        e[j] = h
        _, Bp = background_analytic(x + e, **kw)
        _, Bm = background_analytic(x - e, **kw)
        d = (Bp - Bm) / (2 * h)
        for i in range(3):
            out[:, 3 * i + j] = d[:, i]
    return out




def dipole_test_particles(n_particles, half_width_re=8.0, n_blocks=8, block_cells=(4, 4, 4), ghost_cells=(1, 1, 1), amr_levels=0, seed=1,
                          rigidity_gv=(0.5, 20.0)):
    """Mesh + protons launched in the shell 1.2..6 R_E with isotropic directions and log-uniform rigidity.  Returns
    (mesh, (x, v, w, species, cells))."""
    from . import mesh as meshmod

    L = half_width_re * RE
    refine = None
    if amr_levels > 0:
        # cell size grows with the distance from the planet, like dx = 0.25 R_E (r/R_E) of input/earth-cutoff-rigidity.input
        def refine(level, lo, hi):
            near = np.clip(np.zeros(3), lo, hi)
            r = float(np.linalg.norm(near)) / RE
            return r < half_width_re * (0.45, 0.12, 0.03)[level]  # nested so that neighbours differ by one level at most
    m = meshmod.build_mesh((-L, -L, -L), (L, L, L), (n_blocks,) * 3, block_cells, ghost_cells, periodic=False, refine=refine, max_level=amr_levels)
    rng = np.random.default_rng(seed)
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
    cells = locate_cells(m, x)
    return m, (x, v, np.ones(n_particles), sp, cells)


def locate_cells(m, x):
    """(block, cell) key of every position: leaf by the findTreeNode lattice (vectorised descent), then the cell inside the leaf"""
    N = np.array(m.block_cells)
    gmin = np.array([m.c.x_global_min[d] for d in range(3)])
    dxr = np.array([m.c.dx_max_refinement[d] for d in range(3)])
    lat = np.floor((x.T - gmin) / dxr).astype(np.int64)
    S = 1 << int(m.c.max_refinement_level)
    nroot = np.array([m.c.n_root[d] for d in range(3)])
    r = lat // S
    node = m.arrays["root_node"][r[:, 0] + nroot[0] * (r[:, 1] + nroot[1] * r[:, 2])].astype(np.int64)
    child = m.arrays["node_child"].reshape(-1, 8)
    imin = m.arrays["node_imin"].reshape(-1, 3)
    isize = m.arrays["node_isize"]
    while True:
        has = child[node, 0] >= 0
        if not has.any():
            break
        h = (isize[node] // 2)[:, None]
        o = (lat - imin[node] >= h).astype(np.int64)
        nxt = child[node, o[:, 0] + 2 * (o[:, 1] + 2 * o[:, 2])]
        node = np.where(has, nxt, node)
    leaf = m.arrays["node_leaf"][node].astype(np.int64)
    lo, hi = m.leaf_xmin()[leaf], m.leaf_xmax()[leaf]
    cidx = np.clip(np.floor((x.T - lo) / ((hi - lo) / N)).astype(np.int64), 0, N - 1)
    return (leaf * int(N.prod()) + cidx[:, 0] + N[0] * (cidx[:, 1] + N[1] * cidx[:, 2])).astype(np.int32)
