"""Pins the restated mesh functions against the REFERENCE's own AMR mesh class, compiled from /root/reference/src/meshAMR by
oracle/Makefile into oracle/_ref/libref_mesh.so (with a single-process mpi.h stand-in; no reference source is copied):

  a14  findTreeNode / FindCellIndex                 meshAMRgeneric.h:2793-2882, 2256-2323
  a4   neibNodeFace/Edge/Corner, SetNeibRefinmentLevelLimits (the inputs of the AMR interpolation)  :505-725, :1018-1048
  EPS, dx_max_refinment                              :2340-2365

The reference builds its tree with init() + buildMesh() from a resolution function; the flattened description the product and the
oracle use is generated to mirror that tree, then every query is compared point by point.  Skipped when the library is absent
(it cannot be rebuilt without /root/reference)."""
import ctypes as C
import os

import numpy as np
import pytest

from amps_b200 import api, mesh as meshmod
from oracle.oracle_py import Oracle

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_mesh.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_mesh.so not built (needs /root/reference)")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefMesh:
    def __init__(self, L=40.0, dx0=2.0, radii=(12.0, 6.0)):
        self.lib = C.CDLL(LIB)
        self.lib.ref_mesh_eps.restype = C.c_double
        self.lib.ref_find_cell_index.restype = C.c_long
        self.lib.ref_mesh_build.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p]
        for f in ("ref_find_tree_node", "ref_find_cell_index", "ref_neib", "ref_neib_levels"):
            getattr(self.lib, f).argtypes = None
        lo, hi, r = np.zeros(3), np.full(3, L), np.array(radii, dtype=np.float64)
        assert self.lib.ref_mesh_build(_p(lo), _p(hi), C.c_double(dx0), len(radii), _p(r)) == 0
        self.L = L

    def leaf(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lo, hi = np.zeros(3), np.zeros(3)
        lev = C.c_int()
        r = self.lib.ref_find_tree_node(_p(x), _p(lo), _p(hi), C.byref(lev))
        return None if r < 0 else (lo, hi, lev.value)

    def cell(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        ijk = np.zeros(3, dtype=np.int32)
        nd = self.lib.ref_find_cell_index(_p(x), _p(ijk))
        return int(nd), ijk

    def neib(self, x, kind, idx):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lo, hi = np.zeros(3), np.zeros(3)
        lev = C.c_int()
        r = self.lib.ref_neib(_p(x), kind, idx, _p(lo), _p(hi), C.byref(lev))
        return None if r < 0 else (lo, hi, lev.value)

    def neib_levels(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        mm = np.zeros(2, dtype=np.int32)
        assert self.lib.ref_neib_levels(_p(x), _p(mm)) == 0
        return int(mm[0]), int(mm[1])


@pytest.fixture(scope="module")
def pair():
    ref = RefMesh()
    N = ref.lib.ref_mesh_block_cells()
    G = ref.lib.ref_mesh_ghost_cells()
    Lmax = ref.lib.ref_mesh_max_refinement_level()

    def refine(level, lo, hi):            # split wherever the reference tree is deeper than this node
        return ref.leaf(0.5 * (lo + hi))[2] > level

    m = meshmod.build_mesh((0.0,) * 3, (ref.L,) * 3, (1, 1, 1), (N,) * 3, (G,) * 3, periodic=False, max_refinement_level=Lmax, refine=refine,
                           max_level=8)
    cfg = api.make_config((N,) * 3, (G,) * 3, (1.0,), (1.0,), (1.0,), 1.0, periodic=False, capacity=16)
    return ref, m, Oracle(cfg, m)


def test_tree_constants(pair):
    ref, m, o = pair
    lev = m.leaf_level()
    assert len(set(lev.tolist())) >= 3 and m.c.n_leaves > 100          # a real multi-level tree
    assert m.c.eps == ref.lib.ref_mesh_eps()                            # EPS, :2340-2351
    dx = np.zeros(3)
    ref.lib.ref_mesh_dx_max_refinement(_p(dx))
    assert all(m.c.dx_max_refinement[d] == dx[d] for d in range(3))     # :2365


def test_find_tree_node_and_cell_index_match_the_reference(pair):
    ref, m, o = pair
    rng = np.random.default_rng(0)
    lo_all, hi_all, lev_all = m.leaf_xmin(), m.leaf_xmax(), m.leaf_level()
    pts = rng.uniform(0.0, ref.L, size=(4000, 3))
    # plus points exactly on block faces / cell faces / the domain boundary and just outside
    k = rng.integers(0, 41, size=(1500, 3)).astype(np.float64)
    pts = np.vstack([pts, k, k + rng.choice([0.0, 0.5, 1e-13, -1e-13], size=k.shape), [[-1e-9, 5, 5], [40.0, 3, 3], [40.0 + 1e-9, 3, 3], [0, 0, 0]]])
    n_in = 0
    for x in pts:
        r = ref.leaf(x)
        leaf = o.find_tree_node(x)
        if r is None:
            assert leaf < 0, x
            continue
        assert leaf >= 0, x
        n_in += 1
        assert (lo_all[leaf] == r[0]).all() and (hi_all[leaf] == r[1]).all() and lev_all[leaf] == r[2], (x, r)   # bit-equal block bounds
        nd_ref, ijk_ref = ref.cell(x)
        nd, ijk = o.find_cell_index(x, leaf)
        assert nd == nd_ref and (nd < 0 or (ijk == ijk_ref).all()), (x, nd, nd_ref, ijk, ijk_ref)
    assert n_in > 5000


def test_neighbours_and_level_limits_match_the_reference(pair):
    ref, m, o = pair
    lo_all, hi_all = m.leaf_xmin(), m.leaf_xmax()
    checked = coarser = finer = 0
    for leaf in range(m.c.n_leaves):
        xc = 0.5 * (lo_all[leaf] + hi_all[leaf])
        assert o.neib_levels(leaf) == ref.neib_levels(xc), leaf
        for kind, count in ((0, 6), (1, 12), (2, 8)):
            for idx in range(count):
                a, b = o.neib(leaf, kind, idx), ref.neib(xc, kind, idx)
                assert (a is None) == (b is None), (leaf, kind, idx)
                if a is not None:
                    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2], (leaf, kind, idx)
                    lv = m.leaf_level()[leaf]
                    coarser += a[2] < lv
                    finer += a[2] > lv
                    checked += 1
    assert checked > 2000 and coarser > 50 and finer > 50


def test_specfunc_helpers_match_the_reference():
    """::Relativistic::GetGyroFrequency (it sets the sub-cycle length of Relativistic::Boris, specfunc.h:1290) and Vector3D::Normalize
    (:969-981) as restated in the oracle, against the reference's header compiled into libref_mesh.so: bit-identical"""
    ref = C.CDLL(LIB)
    ref.ref_gyro_frequency.restype = C.c_double
    ref.ref_gyro_frequency.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    ref.ref_speed_of_light.restype = C.c_double
    ref.ref_normalize.argtypes = [C.c_void_p]
    from oracle.oracle_py import load as _load

    ora = _load("parity")
    ora.oracle_probe_gyro_frequency.restype = C.c_double
    ora.oracle_probe_gyro_frequency.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_double]
    ora.oracle_probe_normalize.argtypes = [C.c_void_p]
    c = ref.ref_speed_of_light()
    assert c == 299792458.0
    rng = np.random.default_rng(5)
    qp, mp = 1.602176634e-19, 1.67262192369e-27
    for i in range(2000):
        d = rng.standard_normal(3)
        v = np.ascontiguousarray(d / np.linalg.norm(d) * c * rng.uniform(1e-4, 0.9999))
        B = np.ascontiguousarray(rng.standard_normal(3) * 10.0 ** rng.uniform(-9, -4))
        q = qp * (1 if i % 3 else -1)
        a = ref.ref_gyro_frequency(v.ctypes.data, mp, q, B.ctypes.data)
        b = ora.oracle_probe_gyro_frequency(v.ctypes.data, mp, q, B.ctypes.data, c)
        assert a == b and a > 0
        x1 = np.ascontiguousarray(rng.standard_normal(3) * 10.0 ** rng.uniform(-12, 12))
        x2 = x1.copy()
        ref.ref_normalize(x1.ctypes.data)
        ora.oracle_probe_normalize(x2.ctypes.data)
        assert (x1 == x2).all() and abs(np.linalg.norm(x1) - 1.0) < 1e-15
    z1, z2 = np.zeros(3), np.zeros(3)
    ref.ref_normalize(z1.ctypes.data), ora.oracle_probe_normalize(z2.ctypes.data)
    assert (z1 == 0).all() and (z2 == 0).all()


def test_cpp_flattener_matches_the_python_builder(pair):
    """amps_b200/host/amps_gpu_host_mesh.hpp (the product's UploadMesh: a template over the reference's cMeshAMRgeneric / cTreeNodeAMR)
    run on the reference's own mesh object gives the same flattened description as the Python builder that mirrors the tree:
    nodes, lattice indices, geometry, leaves, boundary faces, unique corner / centre ids and their positions."""
    ref, m, _ = pair
    lib = ref.lib
    N, G = lib.ref_mesh_block_cells(), lib.ref_mesh_ghost_cells()
    sizes = np.zeros(6, dtype=np.int32)
    assert lib.ref_flatten(N, G, _p(sizes)) == 0
    n_nodes, n_leaves, n_corners, n_centers, ncl, nzl = (int(v) for v in sizes)
    assert (n_nodes, n_leaves, n_corners, n_centers) == (m.c.n_nodes, m.c.n_leaves, m.c.n_corners, m.c.n_centers)
    child, level, isize, leaf_node, face = (np.zeros(s, dtype=np.int32) for s in ((n_nodes, 8), n_nodes, n_nodes, n_leaves, n_leaves))
    imin = np.zeros((n_nodes, 3), dtype=np.int32)
    xmin, xmax = np.zeros((n_nodes, 3)), np.zeros((n_nodes, 3))
    cu, zu = np.zeros((n_leaves, ncl), dtype=np.int32), np.zeros((n_leaves, nzl), dtype=np.int32)
    cx, zx = np.zeros((n_corners, 3)), np.zeros((n_centers, 3))
    lib.ref_flatten_arrays(_p(child), _p(level), _p(imin), _p(isize), _p(xmin), _p(xmax), _p(leaf_node), _p(face), _p(cu), _p(zu), _p(cx), _p(zx))
    a = m.arrays
    assert np.array_equal(child, a["node_child"].reshape(-1, 8)) and np.array_equal(level, a["node_level"])
    assert np.array_equal(imin, a["node_imin"].reshape(-1, 3)) and np.array_equal(isize, a["node_isize"])
    assert np.array_equal(xmin, a["node_xmin"].reshape(-1, 3)) and np.array_equal(xmax, a["node_xmax"].reshape(-1, 3))
    assert np.array_equal(leaf_node, a["leaf_node"]) and np.array_equal(face, a["leaf_face_boundary"])
    assert np.array_equal(cu, a["leaf_corner_uid"].reshape(n_leaves, -1)) and np.array_equal(zu, a["leaf_center_uid"].reshape(n_leaves, -1))
    assert np.abs(cx - m.corner_x).max() <= 1e-12 * ref.L and np.abs(zx - m.center_x).max() <= 1e-12 * ref.L
