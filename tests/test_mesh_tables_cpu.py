"""Host-side tables of amps_b200.mesh that the packed J/M rows rely on (no GPU): the 27-slot neighbour table of the unique corners."""
import numpy as np

from amps_b200 import mesh as meshmod


def _opp(s):
    code, inv = {0: 0, -1: 1, 1: 2}, {0: 0, 1: -1, 2: 1}
    d = (inv[s % 3], inv[(s // 3) % 3], inv[s // 9])
    return code[-d[0]] + 3 * code[-d[1]] + 9 * code[-d[2]], d


def test_corner_neighbours_periodic_box():
    m = meshmod.uniform_periodic_box((8, 16, 8), (4, 8, 4), (1, 1, 1))
    nb = m.corner_neighbours()
    x = np.asarray(m.corner_x)
    L = np.array([8.0, 16.0, 8.0])
    real = nb[:, 0] >= 0                                  # corners of depositing blocks (ghost-position nodes have no row)
    assert real.sum() == 8 * 16 * 8 and (nb[real, 0] == np.nonzero(real)[0]).all()
    for s in range(27):
        so, d = _opp(s)
        assert (nb[real, s] >= 0).all()                   # periodic: every neighbour exists
        back = nb[nb[real, s], so]
        assert (back == np.nonzero(real)[0]).all()        # the slot pairing of the packed rows: (c, d) <-> (c + d, -d)
        dx = x[nb[real, s]] - x[real] - np.array(d, dtype=float)
        dx -= L * np.round(dx / L)                        # the neighbour sits one cell away in direction d (modulo the period)
        assert np.abs(dx).max() < 1e-12


def test_corner_neighbours_open_box():
    m = meshmod.build_mesh((0.0, 0.0, 0.0), (8.0, 8.0, 8.0), (2, 2, 2), (4, 4, 4), (1, 1, 1), periodic=False)
    nb = m.corner_neighbours()
    x = np.asarray(m.corner_x)
    inside = (x >= -1e-12).all(1) & (x <= 8 + 1e-12).all(1)
    assert inside.sum() == 9 ** 3
    for s in range(27):
        so, d = _opp(s)
        tgt = x[inside] + np.array(d, dtype=float)
        exists = ((tgt >= -1e-12) & (tgt <= 8 + 1e-12)).all(1)
        have = nb[inside, s] >= 0
        # a neighbour inside the domain is always found; outside there may be a ghost-layer node, never a domain corner
        assert have[exists].all()
        ok = nb[inside, s][exists]
        assert np.abs(x[ok] - tgt[exists]).max() < 1e-12
