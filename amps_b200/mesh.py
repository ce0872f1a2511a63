"""Host-side flattening of the AMR block tree for the device.

In AMPS the tree lives in ``cMeshAMRgeneric`` (src/meshAMR/meshAMRgeneric.h); the
drop-in shim walks ``rootTree`` once per mesh epoch and fills ``amps_gpu_mesh``.
This module builds the same flattened description for synthetic boxes
(uniform periodic / open boxes and sphere-refined AMR boxes) so that tests and
``bench.py`` can run without AMPS.

Geometry conventions follow the reference:
  * periodic mode wraps the user domain in a one-block shell of "ghost" blocks that
    are paired with the real block one period away (src/pic/pic_bc_periodic.cpp:502-571);
  * the lattice used by ``findTreeNode`` has ``1<<max_refinement_level`` points per root
    block (meshAMRgeneric.h:2365, 2392);
  * block-local node numbers are ``_getCornerNodeLocalNumber/_getCenterNodeLocalNumber``
    (meshAMRgeneric.h:74-75).
The forest (``n_root`` root blocks) generalises the reference's single octree so that
boxes whose block count is not 2^m-2 (the reference's periodic constraint,
pic_bc_periodic.cpp:705-747) can be described; ``n_root=(1,1,1)`` is the reference tree.
"""
import ctypes as C

import numpy as np

from . import _capi


class FlatMesh:
    """numpy arrays + the ctypes ``amps_gpu_mesh`` view over them."""

    def __init__(self):
        self.arrays = {}
        self.c = _capi.Mesh()

    def _set(self, name, arr, ctype):
        arr = np.ascontiguousarray(arr)
        self.arrays[name] = arr
        setattr(self.c, name, arr.ctypes.data_as(C.POINTER(ctype)))

    # convenience accessors -------------------------------------------------
    @property
    def n_leaves(self):
        return int(self.c.n_leaves)

    @property
    def n_corners(self):
        return int(self.c.n_corners)

    @property
    def n_centers(self):
        return int(self.c.n_centers)

    @property
    def cells_per_block(self):
        return int(np.prod(self.block_cells))

    @property
    def n_cells(self):
        return self.n_leaves * self.cells_per_block

    def leaf_xmin(self):
        return self.arrays["node_xmin"].reshape(-1, 3)[self.arrays["leaf_node"]]

    def leaf_xmax(self):
        return self.arrays["node_xmax"].reshape(-1, 3)[self.arrays["leaf_node"]]

    def corner_neighbours(self):
        """[n_corners, 27] unique corner at neighbour slot sx+3sy+9sz (per-dimension code 0 -> 0, -1 -> 1, +1 -> 2, the numbering of
        the ECSIM mass matrix, pic_field_solver_ecsim.cpp:625-637), -1 where there is none; from the leaves' corner tables"""
        N, g = np.array(self.block_cells), np.array(self.ghost_cells)
        T = N + 2 * g + 1
        cu = self.arrays["leaf_corner_uid"].reshape(-1, T[2], T[1], T[0]).astype(np.int64)
        nb = np.full((self.n_corners, 27), -1, dtype=np.int64)
        code = {0: 0, -1: 1, 1: 2}
        core = cu[:, g[2]:g[2] + N[2] + 1, g[1]:g[1] + N[1] + 1, g[0]:g[0] + N[0] + 1]
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    s = code[dx] + 3 * code[dy] + 9 * code[dz]
                    sh = cu[:, g[2] + dz:g[2] + dz + N[2] + 1, g[1] + dy:g[1] + dy + N[1] + 1, g[0] + dx:g[0] + dx + N[0] + 1]
                    ok = (core >= 0) & (sh >= 0)
                    nb[core[ok], s] = sh[ok]
        return nb

    def field_solver_tables(self):
        """Node adjacency of the ECSIM field solve on the unique nodes (single-level meshes), from the leaves' node tables:
        corner_nb[n_corners, 27]      = corner_neighbours()
        corner_cells[n_corners, 8]    centre node of the cell at corner index + (a, b, c), a, b, c in {-1, 0}; entry (a+1) + 2 (b+1) + 4 (c+1)
        center_corners[n_centers, 8]  corner node at cell index + (ii, jj, kk) in {0, 1}^3; entry ii + 2 jj + 4 kk
        (-1 where the mesh has no such node: outside an open domain)"""
        N, g = np.array(self.block_cells), np.array(self.ghost_cells)
        T = N + 2 * g
        cu = self.arrays["leaf_corner_uid"].reshape(-1, T[2] + 1, T[1] + 1, T[0] + 1).astype(np.int64)
        zu = self.arrays["leaf_center_uid"].reshape(-1, T[2], T[1], T[0]).astype(np.int64)
        cc = np.full((self.n_corners, 8), -1, dtype=np.int64)
        zc = np.full((self.n_centers, 8), -1, dtype=np.int64)
        core = cu[:, g[2]:g[2] + N[2] + 1, g[1]:g[1] + N[1] + 1, g[0]:g[0] + N[0] + 1]
        for c in (-1, 0):
            for b in (-1, 0):
                for a in (-1, 0):
                    sh = zu[:, g[2] + c:g[2] + c + N[2] + 1, g[1] + b:g[1] + b + N[1] + 1, g[0] + a:g[0] + a + N[0] + 1]
                    ok = (core >= 0) & (sh >= 0)
                    cc[core[ok], (a + 1) + 2 * (b + 1) + 4 * (c + 1)] = sh[ok]
        cells = zu[:, g[2]:g[2] + N[2], g[1]:g[1] + N[1], g[0]:g[0] + N[0]]
        for kk in (0, 1):
            for jj in (0, 1):
                for ii in (0, 1):
                    sh = cu[:, g[2] + kk:g[2] + kk + N[2], g[1] + jj:g[1] + jj + N[1], g[0] + ii:g[0] + ii + N[0]]
                    ok = (cells >= 0) & (sh >= 0)
                    zc[cells[ok], ii + 2 * jj + 4 * kk] = sh[ok]
        return self.corner_neighbours().astype(np.int32), cc.astype(np.int32), zc.astype(np.int32)

    def leaf_level(self):
        return self.arrays["node_level"][self.arrays["leaf_node"]]

    def real_leaves(self):
        """leaves that hold particles on this rank (used, not periodic ghosts, owned)."""
        fl = self.arrays["node_flags"][self.arrays["leaf_node"]]
        ok = ((fl & _capi.NODE_USED) != 0) & ((fl & _capi.NODE_PERIODIC_GHOST) == 0)
        ok &= self.arrays["leaf_owner"] == self.rank
        return np.nonzero(ok)[0]


def _local_numbers(N, g, corner):
    """block-local (i,j,k) grids incl. ghost layers in local-number order."""
    ext = [N[d] + 2 * g[d] + (1 if corner else 0) for d in range(3)]
    k, j, i = np.meshgrid(np.arange(ext[2]) - g[2], np.arange(ext[1]) - g[1], np.arange(ext[0]) - g[0], indexing="ij")
    return i.ravel(), j.ravel(), k.ravel()  # x fastest == local number order


def build_mesh(xmin, xmax, n_blocks, block_cells=(8, 8, 8), ghost_cells=(1, 1, 1), periodic=True,
               max_refinement_level=12, refine=None, max_level=0, rank=0, n_ranks=1, decomp=None, leaf_weight=None):
    """Flatten a box mesh.

    xmin/xmax     user ("original") domain
    n_blocks      level-0 blocks per dimension inside the user domain
    refine        callable(level, xmin[3], xmax[3]) -> bool : split this block? (AMR)
    max_level     deepest level ``refine`` may produce
    rank,n_ranks  this process and the number of processes (one per GPU)
    decomp        (px,py,pz) Cartesian split of the user domain, px*py*pz == n_ranks, or "sfc": the reference's own decomposition
                  (CreateNewParallelDistributionLists / RedistributeParallelLoad, meshAMRgeneric.h:11787-11940): the leaves in the order
                  of the Morton space-filling curve (tree traversal, children 0..7), their load measures normalised so that every rank's
                  share is 1, and the curve cut where the cumulative load passes 1, 2, ... (the leaf that crosses the mark opens the
                  next rank's chunk).  leaf_weight(level, xmin, xmax) -> load measure of a leaf (default 1 = blocks; the reference
                  uses the particle number or the execution time).  Either way every rank owns a contiguous set of blocks;
                  the rank keeps the whole tree but allocates only its own blocks, the blocks around them
                  (DomainBoundaryLayerNodesList, meshAMRgeneric.h:1812) and the real images of adjacent
                  periodic ghost blocks.
    """
    N = np.asarray(block_cells, dtype=np.int64)
    g = np.asarray(ghost_cells, dtype=np.int64)
    nb = np.asarray(n_blocks, dtype=np.int64)
    xmin = np.asarray(xmin, dtype=np.float64)
    xmax = np.asarray(xmax, dtype=np.float64)
    L = int(max_refinement_level)
    S = 1 << L
    shell = 1 if periodic else 0
    n_root = nb + 2 * shell
    dx_root = (xmax - xmin) / nb
    gmin = xmin - shell * dx_root
    gmax = gmin + n_root * dx_root
    if periodic:
        gmax = xmax + shell * dx_root

    m = FlatMesh()
    m.block_cells = tuple(int(v) for v in N)
    m.ghost_cells = tuple(int(v) for v in g)
    m.periodic = bool(periodic)
    m.user_xmin, m.user_xmax = xmin.copy(), xmax.copy()
    m.n_blocks_user = tuple(int(v) for v in nb)

    # ---- tree ------------------------------------------------------------
    parent, child, level, imin, isize, nxmin, nxmax, flags = [], [], [], [], [], [], [], []

    def edge(d, r):  # shared expression so that neighbours agree bitwise
        return gmin[d] + r * dx_root[d]

    def new_node(par, lev, im, sz, lo, hi, fl):
        parent.append(par); child.append([-1] * 8); level.append(lev); imin.append(list(im)); isize.append(sz)
        nxmin.append(list(lo)); nxmax.append(list(hi)); flags.append(fl)
        return len(parent) - 1

    def split(n):
        lo, hi = np.array(nxmin[n]), np.array(nxmax[n])
        mid = 0.5 * (lo + hi)  # bisection as cMeshAMRgeneric::splitTreeNode
        half = isize[n] // 2
        for kk in range(2):
            for jj in range(2):
                for ii in range(2):
                    o = (ii, jj, kk)
                    clo = [lo[d] if o[d] == 0 else mid[d] for d in range(3)]
                    chi = [mid[d] if o[d] == 0 else hi[d] for d in range(3)]
                    cim = [imin[n][d] + o[d] * half for d in range(3)]
                    c = new_node(n, level[n] + 1, cim, half, clo, chi, flags[n])
                    child[n][ii + 2 * (jj + 2 * kk)] = c
                    if refine is not None and level[c] < max_level and refine(level[c], np.array(clo), np.array(chi)):
                        split(c)

    root_node = np.zeros(int(np.prod(n_root)), dtype=np.int32)
    for rk in range(n_root[2]):
        for rj in range(n_root[1]):
            for ri in range(n_root[0]):
                r = (ri, rj, rk)
                lo = [edge(d, r[d]) for d in range(3)]
                hi = [edge(d, r[d] + 1) for d in range(3)]
                ghost = periodic and any(r[d] == 0 or r[d] == n_root[d] - 1 for d in range(3))
                fl = _capi.NODE_USED | (_capi.NODE_PERIODIC_GHOST if ghost else 0)
                n = new_node(-1, 0, [r[d] * S for d in range(3)], S, lo, hi, fl)
                root_node[ri + n_root[0] * (rj + n_root[1] * rk)] = n
                if (not ghost) and refine is not None and max_level > 0 and refine(0, np.array(lo), np.array(hi)):
                    split(n)

    n_nodes = len(parent)
    child = np.array(child, dtype=np.int32)
    level = np.array(level, dtype=np.int32)
    imin = np.array(imin, dtype=np.int32)
    isize = np.array(isize, dtype=np.int32)
    nxmin = np.array(nxmin, dtype=np.float64)
    nxmax = np.array(nxmax, dtype=np.float64)
    flags = np.array(flags, dtype=np.int32)
    is_leaf = (child < 0).all(axis=1)
    gleaf_node = np.nonzero(is_leaf)[0].astype(np.int32)  # global leaf list (same on every rank)
    n_gleaves = len(gleaf_node)
    node_gleaf = -np.ones(n_nodes, dtype=np.int32)
    node_gleaf[gleaf_node] = np.arange(n_gleaves, dtype=np.int32)

    # ---- leaf lookup by lattice point (host mirror of findTreeNode) ---------
    def find_gleaf_ix(ix):
        r = [ix[d] // S for d in range(3)]
        if any(ix[d] < 0 or r[d] >= n_root[d] for d in range(3)):
            return -1
        n = root_node[r[0] + n_root[0] * (r[1] + n_root[1] * r[2])]
        while child[n, 0] >= 0:
            h = isize[n] // 2
            o = [0 if ix[d] - imin[n, d] < h else 1 for d in range(3)]
            n = child[n, o[0] + 2 * (o[1] + 2 * o[2])]
        return int(node_gleaf[n])

    gli = imin[gleaf_node].astype(np.int64)
    gls = isize[gleaf_node].astype(np.int64)
    tot = n_root * S
    gflags = flags[gleaf_node]
    g_is_ghost = (gflags & _capi.NODE_PERIODIC_GHOST) != 0

    # ---- periodic pairing on the global leaf list (findCorrespondingRealBlock, pic_bc_periodic.cpp:502-519)
    gleaf_real = -np.ones(n_gleaves, dtype=np.int32)
    if periodic:
        period = nb * S
        for l in np.nonzero(g_is_ghost)[0]:
            c = gli[l] + gls[l] // 2
            c = (c - S) % period + S
            gleaf_real[l] = find_gleaf_ix([int(v) for v in c])

    # ---- ownership: Cartesian split of the user domain, or the reference's space-filling-curve chunks ----------------------
    gowner = np.zeros(n_gleaves, dtype=np.int32)
    real_mask = ~g_is_ghost
    if isinstance(decomp, str):
        assert decomp == "sfc"
        curve = np.nonzero(real_mask)[0]  # the global leaf list is the tree traversal = the Morton curve
        wts = np.ones(len(curve)) if leaf_weight is None else np.array(
            [float(leaf_weight(int(level[gleaf_node[l]]), nxmin[gleaf_node[l]], nxmax[gleaf_node[l]])) for l in curve])
        norm = wts.sum() / n_ranks  # LoadMeasureNormal: every rank's share is 1
        cum, cur, load_cur = 0.0, 0, 0.0
        for l, wl in zip(curve, wts / norm):
            cum += wl
            if cum > 1.0 + 1e-8 + cur and cur != n_ranks - 1 and load_cur > 0.0:  # RedistributeParallelLoad, :11925-11930
                cur, load_cur = cur + 1, 0.0
            load_cur += wl
            gowner[l] = cur
    else:
        if decomp is None:
            decomp = (n_ranks, 1, 1)
        decomp = np.asarray(decomp, dtype=np.int64)
        assert int(np.prod(decomp)) == n_ranks
        ctr = gli + (gls // 2)[:, None] - shell * S  # block centre on the lattice of the user domain
        ext = nb * S
        rc = np.clip((ctr * decomp[None, :]) // ext[None, :], 0, decomp[None, :] - 1)
        gowner[:] = (rc[:, 0] + decomp[0] * (rc[:, 1] + decomp[1] * rc[:, 2])).astype(np.int32)
    if periodic:
        gh = np.nonzero(g_is_ghost)[0]
        gowner[gh] = gowner[gleaf_real[gh]]

    # ---- local leaves: own + boundary layer + real images of adjacent ghost blocks ----
    own = np.nonzero(real_mask & (gowner == rank))[0]
    if n_ranks == 1:
        local = np.arange(n_gleaves)
    else:
        loc = set(int(v) for v in own)
        extra = []
        for l in own:
            lo, sz = gli[l], int(gls[l])
            # probe one lattice point beyond every face/edge/corner (and the mid points for finer neighbours)
            probes = (-1, 0, sz // 2, sz - 1, sz)
            for a in probes:
                for b in probes:
                    for c3 in probes:
                        if 0 <= a < sz and 0 <= b < sz and 0 <= c3 < sz:
                            continue
                        g2 = find_gleaf_ix([int(lo[0] + a), int(lo[1] + b), int(lo[2] + c3)])
                        if g2 >= 0 and g2 not in loc:
                            loc.add(g2)
                            extra.append(g2)
        for g2 in list(extra):
            r2 = int(gleaf_real[g2])
            if r2 >= 0 and r2 not in loc:
                loc.add(r2)
                extra.append(r2)
        local = np.concatenate([own, np.array(sorted(extra), dtype=np.int64)]) if extra else own
    local = np.asarray(local, dtype=np.int64)
    n_leaves = len(local)
    n_own = len(own) if n_ranks > 1 else n_leaves
    leaf_node = gleaf_node[local].astype(np.int32)
    g2l = -np.ones(n_gleaves, dtype=np.int32)
    g2l[local] = np.arange(n_leaves, dtype=np.int32)
    node_leaf = -np.ones(n_nodes, dtype=np.int32)
    node_leaf[leaf_node] = np.arange(n_leaves, dtype=np.int32)
    leaf_owner = gowner[local].astype(np.int32)
    leaf_real = np.where(gleaf_real[local] >= 0, g2l[np.maximum(gleaf_real[local], 0)], -1).astype(np.int32)

    def find_leaf_ix(ix):
        g2 = find_gleaf_ix(ix)
        return -1 if g2 < 0 else int(g2l[g2])

    m.find_leaf_ix = find_leaf_ix

    # ---- face boundary flags (GetNeibFace(f,0,0)==NULL) ---------------------
    li = gli[local]
    ls = gls[local]
    leaf_face_boundary = np.zeros(n_leaves, dtype=np.int32)
    for d in range(3):
        leaf_face_boundary |= np.where(li[:, d] == 0, 1 << (2 * d), 0).astype(np.int32)
        leaf_face_boundary |= np.where(li[:, d] + ls == tot[d], 1 << (2 * d + 1), 0).astype(np.int32)

    # ---- unique corner / centre nodes -------------------------------------------------
    # integer node keys: corner = imin*N + i*isize  (units: 1/N of a lattice step),
    #                    centre = 2*imin*N + (2i+1)*isize
    # Only leaves that hold particles on this rank (own, or every leaf for a single rank) get node tables.
    has_nodes = np.zeros(n_leaves, dtype=bool)
    has_nodes[:n_own] = True

    def keys_for(corner):
        i, j, k = _local_numbers(N, g, corner)
        loc3 = np.stack([i, j, k], axis=1).astype(np.int64)  # [nloc,3]
        if corner:
            key = li[:, None, :] * N[None, None, :] + loc3[None, :, :] * ls[:, None, None]
            span = nb * S * N
            org = shell * S * N
        else:
            key = 2 * li[:, None, :] * N[None, None, :] + (2 * loc3[None, :, :] + 1) * ls[:, None, None]
            span = 2 * nb * S * N
            org = 2 * shell * S * N
        inside_block = np.ones(loc3.shape[0], dtype=bool)
        for d in range(3):
            hi = N[d] + (1 if corner else 0)
            inside_block &= (loc3[:, d] >= 0) & (loc3[:, d] < hi)
        valid = np.ones(key.shape[:2], dtype=bool)
        if periodic:
            key = (key - org) % span  # identify periodic images
        else:
            for d in range(3):
                valid &= (key[:, :, d] >= 0) & (key[:, :, d] <= span[d])
        enc = (key[:, :, 2] * (span[1] + 1) + key[:, :, 1]) * (span[0] + 1) + key[:, :, 0]
        return enc, valid, inside_block, key, span

    amr = bool((level[gleaf_node] > 0).any())

    def uid_table(corner):
        enc, valid, inside_block, key, span = keys_for(corner)
        valid = valid & has_nodes[:, None]
        if n_ranks == 1 and not amr:
            # nodes that blocks really own; ghost-layer positions only resolve to such nodes
            pool = np.unique(enc[:, inside_block].ravel())
        else:
            # AMR: the ghost cells of a block next to a coarser/finer one are nodes of their own (the reference
            # allocates them per block and the coupler fills them, srcEarth/main_lib.cpp:686-800)
            # every node an own block's tile can touch is kept locally (its value arrives with the field upload)
            pool = np.unique(enc[valid])
        pos = np.searchsorted(pool, enc)
        pos_c = np.minimum(pos, len(pool) - 1)
        found = (pool[pos_c] == enc) & valid
        uid = np.where(found, pos_c, -1).astype(np.int32)
        first = np.full(len(pool), -1, dtype=np.int64)
        flat_ok = found.ravel()
        idx = np.nonzero(flat_ok)[0]
        first[pos_c.ravel()[idx][::-1]] = idx[::-1]
        kk = key.reshape(-1, 3)[first]
        x0 = xmin if periodic else gmin
        den = (S * N) if corner else (2 * S * N)
        xx = x0[None, :] + kk / den[None, :] * dx_root[None, :]
        # deposit targets: in-block corners of own leaves (global keys, for the cross-rank corner exchange)
        # (centres: the cells of own leaves -- every cell belongs to exactly one rank)
        targets = np.unique(enc[:n_own][:, inside_block].ravel())
        return uid, len(pool), xx, pool, targets

    corner_uid, n_corners, corner_x, corner_gkey, corner_targets = uid_table(True)
    center_uid, n_centers, center_x, center_gkey, center_own = uid_table(False)
    m.center_own_gkeys = center_own
    m.corner_x, m.center_x = corner_x, center_x
    m.corner_gkey, m.center_gkey = corner_gkey, center_gkey
    m.corner_target_gkeys = corner_targets
    m.rank, m.n_ranks, m.n_own_leaves = rank, n_ranks, n_own
    m.leaf_global = local.astype(np.int32)
    m.n_global_leaves = n_gleaves

    node_thread = np.zeros(n_nodes, dtype=np.int32)
    node_thread[gleaf_node] = gowner

    c = m.c
    for d in range(3):
        c.n_root[d] = int(n_root[d])
        c.x_global_min[d] = gmin[d]
        c.x_global_max[d] = gmax[d]
        c.dx_max_refinement[d] = (gmax[d] - gmin[d]) / (int(n_root[d]) * S)  # == (xmax-xmin)/(1<<L) for one root
        c.dx_root_block[d] = (gmax[d] - gmin[d]) / int(n_root[d])
    c.max_refinement_level = L
    eps = min(0.0001 * c.dx_root_block[d] / float(N[d]) / S for d in range(3))  # meshAMRgeneric.h:2340-2351
    c.eps = eps
    c.n_nodes = n_nodes
    c.n_leaves = n_leaves
    c.n_corners = n_corners
    c.n_centers = n_centers
    c.this_rank = rank
    c.n_ranks = n_ranks
    c.n_global_leaves = n_gleaves
    m._set("node_parent", np.array(parent, dtype=np.int32), C.c_int32)
    m._set("node_child", child, C.c_int32)
    m._set("node_level", level, C.c_int32)
    m._set("node_imin", imin, C.c_int32)
    m._set("node_isize", isize, C.c_int32)
    m._set("node_xmin", nxmin, C.c_double)
    m._set("node_xmax", nxmax, C.c_double)
    m._set("node_leaf", node_leaf, C.c_int32)
    m._set("node_flags", flags, C.c_int32)
    m._set("node_thread", node_thread, C.c_int32)
    m._set("root_node", root_node, C.c_int32)
    m._set("leaf_node", leaf_node, C.c_int32)
    m._set("leaf_real", leaf_real, C.c_int32)
    m._set("leaf_face_boundary", leaf_face_boundary, C.c_int32)
    m._set("leaf_corner_uid", corner_uid, C.c_int32)
    m._set("leaf_center_uid", center_uid, C.c_int32)
    m._set("leaf_owner", leaf_owner, C.c_int32)
    m._set("leaf_global_id", m.leaf_global, C.c_int32)
    m._set("global_leaf_to_local", g2l, C.c_int32)
    return m


def shared_corner_lists(m, all_targets):
    """Per peer rank: local uids of the corners whose J/M this rank AND the peer deposit into, both sides in the
    order of the global corner key.  all_targets[r] = corner_target_gkeys of rank r (gathered by the caller)."""
    out = {}
    mine = m.corner_target_gkeys
    for r, other in enumerate(all_targets):
        if r == m.rank:
            continue
        common = np.intersect1d(mine, other, assume_unique=True)
        if len(common):
            out[r] = np.searchsorted(m.corner_gkey, common).astype(np.int32)
    return out


def field_halo_lists(m, gathered):
    """Halo lists of the device field solve on several ranks.  gathered[r] = (corner_gkey, corner_target_gkeys, center_gkey,
    center_own_gkeys) of rank r.  A corner is computed by every rank that deposits into it (its row needs only local data once the
    shared J / M sums are exchanged); its PRIMARY rank -- the lowest of them -- counts it in the inner products and sends its value to
    every other rank that holds the corner (as a target or in a ghost layer), after every operator product.  A centre (B) belongs to
    the rank that owns its cell.  Returns (primary_mask[n_corners] uint8, {peer: (corner_send, corner_recv, center_send, center_recv)}),
    all lists as local unique ids in the order of the global key, so that both sides of a pair agree."""
    me = m.rank
    R = len(gathered)
    keys = np.concatenate([g[1] for g in gathered])
    ranks = np.concatenate([np.full(len(g[1]), r, dtype=np.int32) for r, g in enumerate(gathered)])
    order = np.lexsort((ranks, keys))
    ks, rs = keys[order], ranks[order]
    first = np.ones(len(ks), dtype=bool)
    first[1:] = ks[1:] != ks[:-1]
    prim_keys, prim_rank = ks[first], rs[first]          # primary rank of every target corner of the whole mesh

    def primary_of(k):
        pos = np.searchsorted(prim_keys, k)
        pos = np.minimum(pos, len(prim_keys) - 1)
        return np.where(prim_keys[pos] == k, prim_rank[pos], -1)

    mask = np.zeros(m.n_corners, dtype=np.uint8)
    mine_t = m.corner_target_gkeys
    mask[np.searchsorted(m.corner_gkey, mine_t[primary_of(mine_t) == me])] = 1
    lists = {}
    for p in range(R):
        if p == me:
            continue
        common = np.intersect1d(m.corner_gkey, gathered[p][0], assume_unique=True)
        pr = primary_of(common)
        c_send = np.searchsorted(m.corner_gkey, common[pr == me]).astype(np.int32)
        c_recv = np.searchsorted(m.corner_gkey, common[pr == p]).astype(np.int32)
        z_send = np.searchsorted(m.center_gkey, np.intersect1d(m.center_own_gkeys, gathered[p][2], assume_unique=True)).astype(np.int32)
        z_recv = np.searchsorted(m.center_gkey, np.intersect1d(m.center_gkey, gathered[p][3], assume_unique=True)).astype(np.int32)
        if len(c_send) or len(c_recv) or len(z_send) or len(z_recv):
            lists[p] = (c_send, c_recv, z_send, z_recv)
    return mask, lists


def uniform_periodic_box(n_cells, block_cells=(8, 8, 8), ghost_cells=(1, 1, 1), dx=1.0, origin=(0.0, 0.0, 0.0),
                         max_refinement_level=12, rank=0, n_ranks=1, decomp=None, leaf_weight=None):
    """BASELINE config 2/3 geometry: [origin, origin+n_cells*dx) periodic, single AMR level."""
    n_cells = np.asarray(n_cells, dtype=np.int64)
    N = np.asarray(block_cells, dtype=np.int64)
    assert (n_cells % N == 0).all(), "n_cells must be a multiple of block_cells"
    xmin = np.asarray(origin, dtype=np.float64)
    xmax = xmin + n_cells * dx
    return build_mesh(xmin, xmax, n_cells // N, block_cells, ghost_cells, True, max_refinement_level, rank=rank, n_ranks=n_ranks,
                      decomp=decomp, leaf_weight=leaf_weight)
