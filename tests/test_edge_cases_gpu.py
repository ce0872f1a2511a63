"""Edge cases of the ECSIM path through the C ABI: empty store, single particle, particles that land exactly on cell / block /
periodic faces (the fast mover must hand them to the exact kernel), capacity errors, state errors."""
import numpy as np
import pytest

from amps_b200 import _capi, api, mesh as meshmod, workload
from tests import parity_util as pu

pytestmark = pytest.mark.gpu


def _ctx(n_cells=(16, 16, 16), capacity=1024, **kw):
    m = meshmod.uniform_periodic_box(n_cells, (8, 8, 8), (1, 1, 1))
    charge, mass, wgt = workload.species_tables(8, 1.0)
    cfg = api.make_config((8, 8, 8), (1, 1, 1), charge, mass, wgt, 1.0, periodic=True, capacity=capacity, **kw)
    E, B = workload.box_fields(m, E_amp=0.0)
    return m, cfg, (E, B, B)


def test_empty_store_steps_and_deposits_zero():
    m, cfg, f = _ctx()
    g = api.Context(cfg, m)
    g.fields_upload(*f)
    z = np.zeros((3, 0))
    g.particles_upload(z, z, np.zeros(0), np.zeros(0, dtype=np.uint8), np.zeros(0, dtype=np.int32))
    assert g.particle_count() == 0
    st = g.MoveParticles()
    assert st["n_moved"] == 0
    g.sort()
    en, cfl = g.UpdateJMassMatrix()
    J, M = g.JM_download()
    assert en == 0.0 and not J.any() and not M.any()
    g.step()
    J, M = np.ones((m.n_corners, 3)), np.ones((m.n_corners, 243))
    g.step_JM(J, M)
    assert not J.any() and not M.any() and g.particle_count() == 0
    # the other particle passes on an empty store: zeros, no error
    assert not g.ComputeNetCharge(1.0).any()
    assert not g.ComputeSpeciesMoments().any()
    g.SetPhi(np.linspace(0.0, 1.0, m.n_centers))
    assert g.CorrectParticleLocation(1.0, 1.0) == (0, 0)
    g.sort()
    g.SampleCells()
    s, cnt = g.sample_download(clear=True)
    assert not s.any() and not cnt.any()
    g.close()


def test_only_ions_are_not_shifted_and_one_species_cells():
    """CorrectParticleLocation moves species 0 only; cells that hold one species leave the other species' moments and samples zero"""
    m, cfg, parts, fields = pu.make_case(n_cells=(16, 16, 16), ppc=3, seed=91)
    x, v, w, sp, cells = parts
    ions = sp == 1
    g = api.Context(cfg, m)
    g.particles_upload(x[:, ions], v[:, ions], w[ions], sp[ions], cells[ions])
    mom = g.ComputeSpeciesMoments()
    assert not mom[:, 0, :].any() and mom[:, 1, 0].min() >= 0 and mom[:, 1, 0].sum() > 0
    g.SetPhi(np.sin(np.arange(m.n_centers) * 0.37))
    assert g.CorrectParticleLocation(1.0, 1.0) == (0, 0)
    g.sort()
    d = g.particles_download()
    order = np.argsort(d["ptrs"])
    assert (d["x"][:, order] == x[:, ions]).all() and (d["cells"][order] == cells[ions]).all()
    g.SampleCells()
    s, cnt = g.sample_download()
    assert cnt[0] == 0 and cnt[1] == ions.sum() and not s[:, 0, :].any()
    g.close()


def test_single_particle_matches_oracle():
    m, cfg, f = _ctx()
    x = np.array([[3.25], [9.5], [12.75]])
    v = np.array([[0.3], [-0.2], [0.1]])
    cells = workload.locate_cells(m, x)
    parts = (x, v, np.ones(1), np.zeros(1, dtype=np.uint8), cells)
    cfg.exit_record_capacity = 4
    ora = pu.run_oracle(m, cfg, parts, f)
    gpu = pu.run_gpu(m, cfg, parts, f)
    res = pu.compare(m, parts, ora, gpu)
    assert res["cell_mismatch"] == 0 and res["stats_equal"] and res["max_rel_x"] <= pu.REL_TOL and res["max_rel_M"] <= pu.REL_TOL, res


@pytest.mark.parametrize("exact", [0, 1])
def test_particles_landing_exactly_on_faces(exact):
    """E = B = 0: x' = x + v exactly.  Every particle lands on a cell face, many on block faces and on the periodic boundary:
    the integer part of (x'-xmin)/dx decides the cell, so the contracted-arithmetic kernel may not finish any of them."""
    m, cfg, _ = _ctx(n_cells=(16, 16, 16), capacity=20000)
    cfg.exact_arithmetic = exact
    E = np.zeros((m.n_corners, 3))
    B = np.zeros((m.n_centers, 3))
    rng = np.random.default_rng(5)
    n = 12000
    xi = rng.integers(0, 16, size=(3, n)).astype(np.float64)
    x = xi + 0.5
    s = rng.choice([-1.5, -0.5, 0.5, 1.5], size=(3, n))          # x' = integer: on a face in all three dimensions
    s[2, : n // 2] = rng.uniform(-0.4, 0.4, n // 2)              # ... or in two of them
    v = s.copy()
    cells = workload.locate_cells(m, x)
    parts = (x, v, np.ones(n), rng.integers(0, 2, n).astype(np.uint8), cells)
    f = (E, B, B)
    ora = pu.run_oracle(m, cfg, parts, f)
    gpu = pu.run_gpu(m, cfg, parts, f)
    res = pu.compare(m, parts, ora, gpu)
    assert res["cell_mismatch"] == 0 and res["stats_equal"], res
    assert res["bit_mismatch_xv"] == 0, res                      # nothing to round: both paths give the same doubles
    assert res["stats_oracle"]["n_periodic_wrap"] > 0 and res["stats_oracle"]["n_cross_block"] > 0
    if not exact:
        assert res["n_redo"] == n                                # all of them went through the exact kernel
    assert res["max_rel_J"] <= pu.REL_TOL and res["max_rel_M"] <= pu.REL_TOL


def test_capacity_and_state_errors_are_reported():
    m, cfg, f = _ctx(capacity=100)
    g = api.Context(cfg, m)
    n = 200
    x = np.full((3, n), 4.5)
    with pytest.raises(api.AmpsGpuError):                        # more particles than PIC::ParticleBuffer::MaxNPart
        g.particles_upload(x, x * 0, np.ones(n), np.zeros(n, dtype=np.uint8), workload.locate_cells(m, x))
    with pytest.raises(api.AmpsGpuError):                        # cell id outside the mesh
        g.particles_upload(x[:, :4], x[:, :4] * 0, np.ones(4), np.zeros(4, dtype=np.uint8), np.full(4, m.n_cells, dtype=np.int32))
    g.particles_upload(x[:, :4], x[:, :4] * 0, np.ones(4), np.zeros(4, dtype=np.uint8), workload.locate_cells(m, x[:, :4]))
    with pytest.raises(api.AmpsGpuError):                        # Lapenta2017 before the fields were uploaded
        g.MoveParticles()
    g.fields_upload(*f)
    g.MoveParticles()
    with pytest.raises(api.AmpsGpuError):                        # the deposit needs the sorted layout
        g.UpdateJMassMatrix()
    g.sort()
    g.UpdateJMassMatrix()
    with pytest.raises(api.AmpsGpuError):                        # unknown mover id
        g.MoveParticles(99)
    g.close()


def test_wrong_cell_on_upload_is_an_error_like_the_reference_exit():
    # a particle filed under a cell that does not contain it: the reference exit()s ("the point is out of block")
    m, cfg, f = _ctx()
    g = api.Context(cfg, m)
    g.fields_upload(*f)
    x = np.array([[3.5], [3.5], [3.5]])
    far = workload.locate_cells(m, x + 8.0)
    g.particles_upload(x, x * 0, np.ones(1), np.zeros(1, dtype=np.uint8), far)
    with pytest.raises(api.AmpsGpuError):
        g.MoveParticles()
    g.close()


@pytest.mark.gpu
def test_mesh_reupload_starts_a_new_epoch():
    """amps_gpu_mesh_upload on a live context (UpdateBlockTable after a mesh modification): the step on the new mesh matches the
    oracle of the new mesh; one context goes through three meshes (block size, ghost layers and boundary mode are properties of
    the context, like the reference's compile-time switches)"""
    # (same ppc: the species weights are part of the context's configuration)
    kws = [dict(n_cells=(16, 16, 16), ppc=5, seed=71), dict(n_cells=(32, 16, 24), ppc=5, seed=73), dict(n_cells=(16, 8, 8), ppc=5, seed=75)]
    cases = [pu.make_case(**kw) for kw in kws]
    cap = max(c[1].capacity for c in cases)
    g = None
    for m, cfg, parts, fields in cases:
        cfg.capacity = cap
        cfg.exit_record_capacity = cap
        if g is None:
            g = api.Context(cfg, m)
        ora = pu.run_oracle(m, cfg, parts, fields)
        gpu = pu.run_gpu(m, cfg, parts, fields, ctx=g)
        res = pu.compare(m, parts, ora, gpu)
        assert res["cell_mismatch"] == 0 and res["max_rel_x"] <= pu.REL_TOL and res["max_rel_J"] <= pu.REL_TOL and res["max_rel_M"] <= pu.REL_TOL, res
    g.close()


@pytest.mark.gpu
def test_aos_round_trip_carries_the_reduced_state():
    """amps_gpu_particles_upload_aos / _download_aos on 89-byte records (the packed basic record + magnetic moment + v_parallel,
    picParticleDataMacro.h): x, v, w, the species byte with its InitFlag, mu and v_parallel come back in the caller's slots after a
    sort, and the rebuilt cell lists hold every particle once"""
    import ctypes as C

    OFF_NEXT, OFF_PREV, OFF_SPEC, OFF_V, OFF_X, OFF_W, OFF_MU, OFF_VPAR, STRIDE = 0, 8, 16, 17, 41, 65, 73, 81, 89
    m, cfg, parts, fields = pu.make_case(n_cells=(16, 16, 16), ppc=3, seed=95)
    x, v, w, sp, cells = parts
    n = x.shape[1]
    cfg.carry_magnetic_moment, cfg.carry_v_parallel = 1, 1
    rng = np.random.default_rng(7)
    slots = rng.permutation(n + 50)[:n].astype(np.int64)          # the particles sit in scattered ParticleBuffer slots
    mu, vpar = rng.uniform(0.0, 1.0, n), rng.standard_normal(n)
    spb = (sp | np.where(rng.uniform(size=n) < 0.5, 0x40, 0)).astype(np.uint8)  # InitFlag on half of them
    buf = np.zeros((n + 50, STRIDE), dtype=np.uint8)
    for off, arr in ((OFF_V, v.T), (OFF_X, x.T)):
        buf[slots[:, None], off + np.arange(24)[None, :]] = np.ascontiguousarray(arr).view(np.uint8).reshape(n, 24)
    for off, arr in ((OFF_W, w), (OFF_MU, mu), (OFF_VPAR, vpar)):
        buf[slots[:, None], off + np.arange(8)[None, :]] = np.ascontiguousarray(arr).view(np.uint8).reshape(n, 8)
    buf[slots, OFF_SPEC] = spb | 0x80                                # allocated bit
    lay = _capi.AosLayout()
    lay.stride, lay.off_species, lay.off_v, lay.off_x, lay.off_w, lay.off_mu, lay.off_next, lay.off_prev, lay.off_vpar = (
        STRIDE, OFF_SPEC, OFF_V, OFF_X, OFF_W, OFF_MU, OFF_NEXT, OFF_PREV, OFF_VPAR)
    g = api.Context(cfg, m)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    g._ck(g.lib.amps_gpu_particles_upload_aos(g._h, vp(buf), vp(slots), vp(np.ascontiguousarray(cells, dtype=np.int32)), n, C.byref(lay)))
    g.sort()
    out = np.zeros_like(buf)
    out[:, OFF_SPEC] = 0x80
    n_cells = m.c.n_leaves * int(np.prod(m.block_cells))
    first = np.zeros(n_cells, dtype=np.int64)
    nn = C.c_int64()
    g._ck(g.lib.amps_gpu_particles_download_aos(g._h, vp(out), vp(first), n + 50, C.byref(lay), C.byref(nn)))
    g.close()
    assert nn.value == n
    for off, ln in ((OFF_V, 24), (OFF_X, 24), (OFF_W, 8), (OFF_MU, 8), (OFF_VPAR, 8)):
        assert (out[slots, off:off + ln] == buf[slots, off:off + ln]).all(), off
    assert (out[slots, OFF_SPEC] == buf[slots, OFF_SPEC]).all()
    # walk the lists: every slot exactly once, in its cell
    nxt = out[:, OFF_NEXT:OFF_NEXT + 8].copy().view(np.int64).ravel()
    seen = np.zeros(n + 50, dtype=np.int32)
    cell_of = np.full(n + 50, -1, dtype=np.int64)
    cell_of[slots] = cells
    for c in np.nonzero(first >= 0)[0]:
        p = first[c]
        while p != -1:
            seen[p] += 1
            assert cell_of[p] == c
            p = nxt[p]
    assert (seen[slots] == 1).all() and seen.sum() == n


@pytest.mark.gpu
def test_reduced_state_follows_the_fused_step():
    """amps_gpu_step reorders the store inside the deposit (gather through the permutation): the optional magnetic moment and
    v_parallel arrays must follow their particles there as they do in the stand-alone sort"""
    m, cfg, parts, fields = pu.make_case(n_cells=(16, 16, 16), ppc=6, seed=97, vscale=6.0)
    x, v, w, sp, cells = parts
    n = x.shape[1]
    cfg.carry_magnetic_moment, cfg.carry_v_parallel = 1, 1
    mu0, vp0 = 1.0 + 1e-3 * np.arange(n), -2.0 - 1e-3 * np.arange(n)
    g = api.Context(cfg, m)
    g.fields_upload(*fields)
    g.particles_upload(x, v, w, sp, cells)
    g.magnetic_moment_upload(mu0)
    g.v_parallel_upload(vp0)
    for it in range(3):
        g.step()
    d = g.particles_download()
    mu, vp = g.magnetic_moment_download(), g.v_parallel_download()
    J = np.empty((m.n_corners, 3))
    M = np.empty((m.n_corners, 243))
    g.step_JM(J, M)  # the ranged deposit of the pipelined variant gathers too
    d2 = g.particles_download()
    mu2, vp2 = g.magnetic_moment_download(), g.v_parallel_download()
    g.close()
    assert len(d["ptrs"]) == n and (np.sort(d["ptrs"]) == np.arange(n)).all()
    assert (mu == mu0[d["ptrs"]]).all() and (vp == vp0[d["ptrs"]]).all()
    assert (mu2 == mu0[d2["ptrs"]]).all() and (vp2 == vp0[d2["ptrs"]]).all()
    assert (np.diff(d2["cells"].astype(np.int64)) >= 0).all()
    assert (d["ptrs"] != np.arange(n)).any()          # the steps really permuted the store
