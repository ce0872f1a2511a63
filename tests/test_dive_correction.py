"""f4 (second half): the particle shift of the div-E correction, ECSIM::CorrectParticleLocation
(pic_field_solver_ecsim.cpp:4440-4688), and the per-species corner moments it reads (the
_PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_ part of ProcessCell / UpdateJMassMatrix, :2270-2300, :2384-2392, :3874-3879).
CPU: properties of the restatement; GPU: moments within 1e-10, positions within 1e-12 of a cell, cells bit-exact."""
import numpy as np
import pytest

from amps_b200 import api
from oracle.oracle_py import Oracle
from tests import parity_util as pu


def _smooth_phi(m, amp):
    """a smooth periodic potential sampled on the unique centre nodes"""
    xc = np.asarray(m.center_x)  # [n_centers, 3]
    L = np.array([m.c.x_global_max[d] - m.c.x_global_min[d] for d in range(3)])
    k = 2 * np.pi / L
    return amp * (np.sin(k[0] * xc[:, 0]) * np.cos(k[1] * xc[:, 1]) + 0.5 * np.sin(2 * k[2] * xc[:, 2] + 0.3))


def _oracle_case(kw, amp):
    m, cfg, parts, fields = pu.make_case(**kw)
    x, v, w, sp, cells = parts
    o = Oracle(cfg, m)
    o.add_particles(x, v, w, sp, cells)
    return m, cfg, parts, o, _smooth_phi(m, amp)


def test_species_moments_cpu():
    m, cfg, parts, o, phi = _oracle_case(dict(n_cells=(8, 8, 8), ppc=6, seed=51), 1e-3)
    x, v, w, sp, cells = parts
    mom = o.species_moments()
    o.close()
    for s in range(cfg.n_species):
        sel = sp == s
        mw = cfg.mass[s] * cfg.species_weight[s] * w[sel]
        # unit cells, periodic box: the corner weights of a particle sum to one, nothing is dropped
        assert abs(mom[:, s, 0].sum() - mw.sum()) <= 1e-10 * mw.sum()
        for d in range(3):
            assert abs(mom[:, s, 1 + d].sum() - (mw * v[d, sel]).sum()) <= 1e-10 * (mw * np.abs(v[d, sel])).sum()
            assert abs(mom[:, s, 4 + d].sum() - (mw * v[d, sel] ** 2).sum()) <= 1e-10 * (mw * v[d, sel] ** 2).sum()
        assert (mom[:, s, 0] >= 0).all() and (mom[:, s, 4:7] >= 0).all()


def test_correct_particle_location_cpu():
    m, cfg, parts, o, phi = _oracle_case(dict(n_cells=(8, 8, 8), ppc=6, seed=53), 2e-4)
    x, v, w, sp, cells = parts
    o.species_moments()
    o.set_phi(phi)
    rc, n_disp, n_del, fc = o.correct_particle_location(1.0, 1.0)
    after = o.particles()
    assert rc == 0 and o.check_lists() == 0
    assert n_disp == int((sp == 0).sum()) and n_del == 0
    d = after["x"] - x
    assert np.abs(d[:, sp != 0]).max() == 0.0                       # only species 0 moves
    dx_cell = 1.0
    assert 0 < np.abs(d[:, sp == 0]).max() <= 0.1 * dx_cell * (1 + 1e-12)  # limited to a tenth of a cell (:4625-4634)
    assert (fc >= 0).all()
    # a constant potential moves nothing
    o.set_phi(np.full(m.n_centers, 0.37))
    x1 = o.particles()["x"].copy()
    o.species_moments()
    o.correct_particle_location(1.0, 1.0)
    assert np.abs(o.particles()["x"] - x1).max() == 0.0
    o.close()


CASES = [dict(n_cells=(16, 16, 16), ppc=8, seed=55), dict(n_cells=(32, 16, 8), ppc=5, seed=57, block_cells=(16, 8, 4)),
         dict(n_cells=(16, 16, 16), ppc=4, seed=59, ghost_cells=(2, 2, 2)), dict(n_cells=(16, 16, 16), ppc=6, seed=61, periodic=False),
         dict(n_cells=(16, 16, 16), ppc=6, seed=63, four_species=True)]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_gpu_species_moments_match_oracle(kw):
    m, cfg, parts, o, phi = _oracle_case(kw, 1e-3)
    ref = o.species_moments()
    o.close()
    x, v, w, sp, cells = parts
    g = api.Context(cfg, m)
    g.particles_upload(x, v, w, sp, cells)
    mom = g.ComputeSpeciesMoments()
    g.close()
    for s in range(cfg.n_species):
        for k in range(10):
            assert pu.rel_scaled(mom[:, s, k], ref[:, s, k]) <= pu.REL_TOL, (s, k)
    assert np.abs(ref).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("amp", [2e-4, 5e-2])  # small shifts / shifts cut at a tenth of a cell
def test_gpu_correct_particle_location_matches_oracle(kw, amp):
    m, cfg, parts, o, phi = _oracle_case(kw, amp)
    x, v, w, sp, cells = parts
    o.species_moments()
    o.set_phi(phi)
    rc, n_disp, n_del, fc = o.correct_particle_location(0.9, 1.1)
    ref = o.particles()
    o.close()
    assert rc == 0
    g = api.Context(cfg, m)
    g.particles_upload(x, v, w, sp, cells)
    g.ComputeSpeciesMoments(download=False)
    g.SetPhi(phi)
    nd, nx = g.CorrectParticleLocation(0.9, 1.1)
    g.sort()
    got = g.particles_download()
    g.close()
    assert (nd, nx) == (n_disp, n_del)
    alive = fc >= 0
    order = np.argsort(got["ptrs"])
    assert (np.sort(got["ptrs"]) == np.nonzero(alive)[0]).all()
    # the species moments differ in the last bits (summation order), so the shift does too: 1e-12 of a cell
    assert np.abs(got["x"][:, order] - ref["x"][:, alive]).max() <= 1e-12
    same = got["cells"][order] == fc[alive]
    if not same.all():
        # a particle within 1e-12 of a cell face may be filed on the other side
        xr = ref["x"][:, alive][:, ~same]
        dist = np.abs(xr - np.round(xr)).min(axis=0)
        assert (dist <= 1e-11).all() and (~same).sum() <= 2
    assert np.abs(got["x"][:, order] - x[:, alive]).max() > 0
