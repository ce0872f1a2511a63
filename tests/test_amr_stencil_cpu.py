"""a4, AMR branch of CellCentered::Linear::InitStencil (pic_interpolation_routines.cpp:224-1070) as restated by the oracle.

The reference's design properties pin the restatement: the weights sum to 1 and the stencil reproduces a linear field
exactly (coarse-lattice trilinear interpolation; a coarse centre covered by a finer block is the average of its 2x2x2
fine cells; the coarse and the fine stencil are blended linearly between half a cell and one cell from the interface),
so the interpolant is also continuous across coarse/fine interfaces."""
import numpy as np

from amps_b200 import _capi, api, workload
from oracle.oracle_py import Oracle


def _setup(ghost=(1, 1, 1)):
    m = workload.amr_sphere_box((4, 4, 4), (4, 4, 4), ghost, radii=(5.0, 2.5))
    cfg = api.make_config((4, 4, 4), ghost, (1.0,), (1.0,), (1.0,), 1.0, periodic=False, capacity=16)
    cfg.coupler_interpolation = _capi.CPLR_LINEAR
    return m, cfg, Oracle(cfg, m)


def test_neighbour_level_limits():
    m, cfg, o = _setup()
    lev = m.leaf_level()
    assert sorted(np.unique(lev)) == [0, 1, 2]
    mm = np.array([o.neib_levels(l) for l in range(m.c.n_leaves)])
    assert (mm[:, 0] <= mm[:, 1]).all() and (np.abs(mm[:, 0] - lev) <= 1).all() and (np.abs(mm[:, 1] - lev) <= 1).all()   # 2:1 balanced
    assert ((mm[:, 0] < lev) | (mm[:, 1] > lev)).sum() > 100                                                              # interfaces exist
    o.close()


def test_linear_field_is_reproduced_exactly_and_weights_sum_to_one():
    for ghost in ((1, 1, 1), (2, 2, 2)):
        m, cfg, o = _setup(ghost)
        a = np.array([0.3, -1.2, 0.7])
        fc = 2.0 + m.center_x @ a
        rng = np.random.default_rng(0)
        lo, hi = m.leaf_xmin(), m.leaf_xmax()
        lens, n_pts, worst = {}, 0, 0.0
        for l in range(m.c.n_leaves):
            for _ in range(40):
                x = lo[l] + rng.random(3) * (hi[l] - lo[l])
                if (x < 1.6).any() or (x > 16 - 1.6).any():
                    continue  # next to the domain boundary the stencil is truncated and renormalised
                n, ids, w = o.coupler_stencil(x, l)
                assert n > 0
                lens[n] = lens.get(n, 0) + 1
                n_pts += 1
                assert abs(w.sum() - 1.0) < 1e-13 and (w >= 0).all()
                worst = max(worst, abs((w * fc[ids]).sum() - (2.0 + x @ a)))
        assert worst < 1e-12, worst
        assert n_pts > 10000 and max(lens) > 8 and lens.get(8, 0) > 0, lens     # both the plain and the multi-block stencils occur
        o.close()


def test_continuity_across_a_coarse_fine_interface():
    m, cfg, o = _setup()
    rng = np.random.default_rng(1)
    f = rng.standard_normal(m.n_centers)            # arbitrary (non-smooth) data: continuity must not rely on linearity
    lev = m.leaf_level()
    lo, hi = m.leaf_xmin(), m.leaf_xmax()
    n_checked = 0
    for l in np.nonzero(lev == 1)[0]:
        # step across the +x face into the neighbour if that one is coarser
        xq = np.array([hi[l, 0], 0.5 * (lo[l, 1] + hi[l, 1]) + 0.13, 0.5 * (lo[l, 2] + hi[l, 2]) - 0.21])
        gmin = np.array([m.c.x_global_min[d] for d in range(3)])
        dxr = np.array([m.c.dx_max_refinement[d] for d in range(3)])
        nb = m.find_leaf_ix([int(v) for v in np.floor((xq + np.array([1e-6, 0, 0]) - gmin) / dxr)])
        if nb < 0 or lev[nb] != 0 or xq[0] > 14.0:
            continue
        vals = []
        for eps, leaf in ((-1e-9, l), (1e-9, nb)):
            x = xq + np.array([eps, 0.0, 0.0])
            n, ids, w = o.coupler_stencil(x, int(leaf))
            assert n > 0
            vals.append((w * f[ids]).sum())
        assert abs(vals[0] - vals[1]) < 1e-6, (l, vals)
        n_checked += 1
    assert n_checked >= 4
    o.close()
