"""f3: PIC::Sampling::SamplingManager / ProcessCell (pic.cpp:705-1082) on the device store.  CPU: the sampled datums against
direct numpy sums; GPU: the collecting buffer within 1e-12 of the oracle's over two sampling passes with a move in between."""
import numpy as np
import pytest

from amps_b200 import api
from oracle.oracle_py import Oracle
from tests import parity_util as pu


def test_oracle_samples_are_the_cell_sums():
    m, cfg, parts, fields = pu.make_case(n_cells=(8, 8, 8), ppc=5, seed=81)
    x, v, w, sp, cells = parts
    o = Oracle(cfg, m)
    o.add_particles(x, v, w, sp, cells)
    s, cnt = o.sample_cells()
    o.close()
    for q in range(cfg.n_species):
        sel = sp == q
        assert cnt[q] == sel.sum()
        ww = cfg.species_weight[q] * w[sel]
        assert np.allclose(np.bincount(cells[sel], weights=ww, minlength=s.shape[0]), s[:, q, 0], rtol=1e-13)
        assert np.allclose(np.bincount(cells[sel], minlength=s.shape[0]), s[:, q, 1])
        assert np.allclose(s[:, q, 2], s[:, q, 0] / 1.0, rtol=1e-13)  # unit cells: Measure = 1
        for d in range(3):
            assert np.allclose(np.bincount(cells[sel], weights=ww * v[d, sel], minlength=s.shape[0]), s[:, q, 3 + d], rtol=1e-12, atol=1e-18)
            assert np.allclose(np.bincount(cells[sel], weights=ww * v[d, sel] ** 2, minlength=s.shape[0]), s[:, q, 6 + d], rtol=1e-12)
            assert np.allclose(np.bincount(cells[sel], weights=ww * v[d, sel] * v[(d + 1) % 3, sel], minlength=s.shape[0]), s[:, q, 10 + d], rtol=1e-11,
                               atol=1e-18)
        assert np.allclose(np.bincount(cells[sel], weights=ww * np.sqrt((v[:, sel] ** 2).sum(0)), minlength=s.shape[0]), s[:, q, 9], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(n_cells=(16, 16, 16), ppc=8, seed=83), dict(n_cells=(32, 16, 8), ppc=5, seed=85, block_cells=(16, 8, 4)),
                                dict(n_cells=(16, 16, 16), ppc=6, seed=87, periodic=False), dict(n_cells=(16, 16, 16), ppc=6, seed=89, four_species=True)])
def test_gpu_sampling_matches_oracle(kw):
    m, cfg, parts, fields = pu.make_case(**kw)
    x, v, w, sp, cells = parts
    E, B, Bcur = fields
    o = Oracle(cfg, m)
    o.set_fields(E, B, Bcur)
    o.add_particles(x, v, w, sp, cells)
    o.sample_cells()
    o.move(0, 1)
    ref, rcnt = o.sample_cells()
    o.close()
    g = api.Context(cfg, m)
    g.fields_upload(E, B, Bcur)
    g.particles_upload(x, v, w, sp, cells)
    g.SampleCells()
    g.MoveParticles()
    g.sort()
    g.SampleCells()
    got, cnt = g.sample_download(clear=True)
    zero, zcnt = g.sample_download()
    g.close()
    assert (cnt == rcnt).all() and (zcnt == 0).all() and not zero.any()
    assert (got[:, :, 1] == ref[:, :, 1]).all()                     # particle numbers: exact
    for k in range(13):
        assert pu.rel_scaled(got[:, :, k], ref[:, :, k]) <= 1e-12, k
    assert np.abs(ref).max() > 0
