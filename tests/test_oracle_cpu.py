"""CPU tests of the oracle (no GPU): structural invariants the reference states, agreement of its serial and
OpenMP variants (the reference asserts all CPU variants agree to 2e-14, MakefileTest/Table:2533), and an
independent numpy statement of the same equations."""
import numpy as np
import pytest

from tests import numpy_ref, parity_util as pu
from oracle.oracle_py import Oracle

CASE = dict(n_cells=(16, 16, 16), ppc=4, seed=21)


@pytest.fixture(scope="module")
def case():
    m, cfg, parts, fields = pu.make_case(**CASE)
    return m, cfg, parts, fields, pu.run_oracle(m, cfg, parts, fields)


def test_lists_and_counts(case):
    m, cfg, parts, fields, ora = case
    assert ora["rc"] == 0 and ora["lists"] == 0
    st = ora["stats"]
    assert st["n_moved"] == parts[0].shape[1] and st["n_left_domain"] == 0 and st["n_error"] == 0
    # periodic box: everything stays inside the user domain after the wrap
    x = ora["particles"]["x"]
    for d in range(3):
        assert x[d].min() >= m.user_xmin[d] and x[d].max() < m.user_xmax[d]
    # final cell is consistent with the final position
    lx = m.leaf_xmin()
    cells = ora["final_cell"]
    C = m.cells_per_block
    leaf, cin = cells // C, cells % C
    N = np.array(m.block_cells)
    ijk = np.stack([cin % N[0], (cin // N[0]) % N[1], cin // (N[0] * N[1])])
    lo = lx[leaf].T + ijk
    assert (x >= lo - 1e-12).all() and (x <= lo + 1 + 1e-12).all()


def test_openmp_variant_agrees(case):
    m, cfg, parts, fields, ora = case
    omp = pu.run_oracle(m, cfg, parts, fields, n_threads=4)
    assert (omp["final_cell"] == ora["final_cell"]).all()
    assert (omp["particles"]["x"] == ora["particles"]["x"]).all() and (omp["particles"]["v"] == ora["particles"]["v"]).all()
    assert omp["stats"] == ora["stats"]
    assert np.abs(omp["M"] - ora["M"]).max() <= 2e-14 * max(1.0, np.abs(ora["M"]).max())
    assert np.abs(omp["J"] - ora["J"]).max() <= 2e-14 * max(1.0, np.abs(ora["J"]).max())


def test_corner_weights_sum_to_one(case):
    m, cfg, parts, fields, _ = case
    o = Oracle(cfg, m)
    rng = np.random.default_rng(0)
    leaf = int(m.real_leaves()[3])
    lo, hi = m.leaf_xmin()[leaf], m.leaf_xmax()[leaf]
    for _ in range(200):
        x = lo + rng.random(3) * (hi - lo)
        n, xs, W, ids, wn = o.corner_stencil(x, leaf)
        assert n == 8
        assert abs(W.sum() - 1.0) < 4e-16 and abs(wn.sum() - 1.0) < 4e-16 and (W >= 0).all()
        nc, cid, cw = o.center_stencil(x, leaf)
        assert nc == 8 and abs(cw.sum() - 1.0) < 4e-16
    # the snap rule: a point within 1e-10 dx of xmax is moved to xmax - 1e-10 dx (mutates the caller's x)
    x = np.array([hi[0] - 1e-13, lo[1] + 0.5, lo[2] + 0.5])
    n, xs, W, ids, wn = o.corner_stencil(x, leaf)
    assert xs[0] == hi[0] - 1e-10 * (hi[0] - lo[0]) / m.block_cells[0]
    o.close()


def test_mass_matrix_symmetry(case):
    """both (c,c') and (c',c) receive the same 3x3 block (pic_field_solver_ecsim.cpp:2411-2420):
    M[g][slot(n)] == M[g+n][slot(-n)] on the periodic corner lattice"""
    m, cfg, parts, fields, ora = case
    n = np.array(CASE["n_cells"])
    M = ora["M"].reshape(n[2], n[1], n[0], 27, 9)  # uid = x fastest
    f = lambda d: 0 if d == 0 else (1 if d < 0 else 2)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                s1 = f(dx) + 3 * f(dy) + 9 * f(dz)
                s2 = f(-dx) + 3 * f(-dy) + 9 * f(-dz)
                shifted = np.roll(M[:, :, :, s2, :], shift=(-dz, -dy, -dx), axis=(0, 1, 2))
                assert np.abs(M[:, :, :, s1, :] - shifted).max() <= 1e-13 * np.abs(M).max()


def test_against_independent_numpy_statement(case):
    m, cfg, parts, fields, ora = case
    x, v, w, sp, cells = parts
    E, B, Bcur = fields
    n = np.array(CASE["n_cells"])
    # uid tables of the uniform periodic box are x-fastest lattices
    Ec = E.reshape(n[2], n[1], n[0], 3).transpose(2, 1, 0, 3)
    Bp = B.reshape(n[2], n[1], n[0], 3).transpose(2, 1, 0, 3)
    Bc = Bcur.reshape(n[2], n[1], n[0], 3).transpose(2, 1, 0, 3)
    q = np.array([cfg.charge[s] for s in range(2)])[sp]
    ms = np.array([cfg.mass[s] for s in range(2)])[sp]
    xn, vn = numpy_ref.push(x.T.copy(), v.T.copy(), q / ms, cfg.time_step[0], Ec, Bp, n)
    ox, ov = ora["particles"]["x"].T, ora["particles"]["v"].T
    assert np.abs(vn - ov).max() <= 1e-12 * np.abs(ov).max()
    d = np.abs(xn - ox)
    d = np.minimum(d, np.abs(d - n[None, :]))  # the clamp at the wrap may differ by the 1e-10 L rule
    assert d.max() <= 1e-9
    wt = np.array([cfg.species_weight[s] for s in range(2)])[sp] * w
    J, M = numpy_ref.deposit(ox, ov, q * wt, ms * wt, cfg.ecsim_dt_total, Bc, n)
    Jo = ora["J"].reshape(n[2], n[1], n[0], 3).transpose(2, 1, 0, 3)
    Mo = ora["M"].reshape(n[2], n[1], n[0], 27, 9).transpose(2, 1, 0, 3, 4)
    assert np.abs(J - Jo).max() <= 1e-10 * np.abs(Jo).max()
    assert np.abs(M - Mo).max() <= 1e-10 * np.abs(Mo).max()


def test_deposit_is_linear_in_the_particle_set(case):
    m, cfg, parts, fields, ora = case
    x, v, w, sp, cells = parts
    half = x.shape[1] // 2
    res = []
    for sl in (slice(0, half), slice(half, None)):
        o = Oracle(cfg, m)
        o.set_fields(*fields)
        o.add_particles(x[:, sl], v[:, sl], w[sl], sp[sl], cells[sl])
        res.append(o.deposit(1))
        o.close()
    o = Oracle(cfg, m)
    o.set_fields(*fields)
    o.add_particles(x, v, w, sp, cells)
    full = o.deposit(1)
    o.close()
    assert np.abs(res[0][1] + res[1][1] - full[1]).max() <= 1e-12 * np.abs(full[1]).max()
    assert np.abs(res[0][0] + res[1][0] - full[0]).max() <= 1e-12 * np.abs(full[0]).max()
    assert abs(res[0][2] + res[1][2] - full[2]) <= 1e-12 * abs(full[2])


def test_find_tree_node_and_cell_index(case):
    m, cfg, parts, fields, _ = case
    o = Oracle(cfg, m)
    rng = np.random.default_rng(5)
    gmin = np.array(m.c.x_global_min[:])
    gmax = np.array(m.c.x_global_max[:])
    lx, hx = m.leaf_xmin(), m.leaf_xmax()
    for _ in range(300):
        x = gmin + rng.random(3) * (gmax - gmin)
        leaf = o.find_tree_node(x, int(rng.integers(0, m.n_leaves)))
        assert leaf >= 0 and (x >= lx[leaf]).all() and (x < hx[leaf]).all()
        r, ijk = o.find_cell_index(x, leaf)
        assert r >= 0 and (ijk == np.floor(x - lx[leaf]).astype(int)).all()
    assert o.find_tree_node(gmax + 1.0, 0) == -1 and o.find_tree_node(gmin - 1e-9, 0) == -1
    # a point exactly on a block face belongs to the upper block (x >= xmax -> ix++)
    leaf = int(m.real_leaves()[0])
    xf = np.array([hx[leaf][0], lx[leaf][1] + 0.5, lx[leaf][2] + 0.5])
    up = o.find_tree_node(xf, leaf)
    assert up != leaf and lx[up][0] == hx[leaf][0]
    o.close()
