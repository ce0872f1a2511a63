"""The headline path against the REFERENCE'S OWN CODE ("parity unpinned on the headline path").

oracle/_ref/libref_pic.so is the reference's src/pic, src/general, src/meshAMR, ... compiled from /root/reference for its own
ECSIM test `input/test/fast-wave.input` (recipe: oracle/ref_pic/build_ref_pic.sh; 118 of the reference's translation units, our
stand-ins only for mpi.h and three un-vendored SWMF headers).  On the fast-wave plasma (783 360 particles, 32x16x8 cells in
16x8x4-cell blocks, periodic, individual weight corrections, non-trivial E, B_prev, B_cur) it runs
    ECSIM::UpdateJMassMatrix  ->  PIC::Mover::MoveParticles (Lapenta2017) + ExchangeParticleData + Periodic::ExchangeParticles
    ->  ECSIM::UpdateJMassMatrix
and the CPU oracle (and, with a GPU, the CUDA path through the C ABI) must reproduce it:
  * x', v' of every particle BIT-IDENTICAL, (block, cell) identical  (rows a3, a4, a5, a14, a16 of SURVEY 8a)
  * J and the mass matrix <= 2e-14 of the array maximum, the reference's own cross-variant tolerance (MakefileTest/Table:2533);
    the particle energy <= 1e-12                                     (rows a10, a11, a12)
The same is asserted against tests/golden/ref_fastwave.npz, the reference's results for every 32nd particle (committed, so the
pin survives where the library cannot be built: tests/golden/make_ref_fastwave.py wrote it)."""
import os

import numpy as np
import pytest

from amps_b200 import api, mesh as meshmod
from oracle.oracle_py import Oracle
from oracle.ref_pic import ref_pic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fastwave.npz")
TOL_JM = 2e-14
needs_ref = pytest.mark.skipif(not ref_pic.available(), reason="oracle/_ref/libref_pic.so not built (needs /root/reference at build time)")


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def oracle_phase(m, cfg, parts, fields):
    x, v, w, sp, cells = parts
    o = Oracle(cfg, m, "parity")
    o.set_fields(*fields)
    o.add_particles(x, v, w, sp, cells)
    J0, M0, e0, _ = o.deposit(1)
    rc, st, ret, fc = o.move(0, 1)
    pp = o.particles()
    J1, M1, e1, _ = o.deposit(1)
    o.close()
    return {"J0": J0, "M0": M0, "J1": J1, "M1": M1, "energy": (e0, e1), "x": pp["x"], "v": pp["v"], "cells": fc.astype(np.int64), "stats": st}


def gpu_phase(m, cfg, parts, fields, exact):
    x, v, w, sp, cells = parts
    cfg.exact_arithmetic = 1 if exact else 0
    g = api.Context(cfg, m)
    g.fields_upload(*fields)
    g.particles_upload(x, v, w, sp, cells)
    e0, _ = g.UpdateJMassMatrix()
    J0, M0 = g.JM_download()
    st = g.MoveParticles()
    mv = g.particles_download()  # slot i still holds the particle it held before the move
    g.sort()
    e1, _ = g.UpdateJMassMatrix()
    J1, M1 = g.JM_download()
    g.close()
    n = x.shape[1]
    gx, gv, gc = np.empty((3, n)), np.empty((3, n)), np.empty(n, dtype=np.int64)
    gx[:, mv["ptrs"]], gv[:, mv["ptrs"]], gc[mv["ptrs"]] = mv["x"], mv["v"], mv["cells"]
    return {"J0": J0, "M0": M0, "J1": J1, "M1": M1, "energy": (e0, e1), "x": gx, "v": gv, "cells": gc, "stats": st}


# ---------------- the live reference ----------------
@pytest.fixture(scope="module")
def live():
    from tests import ref_ecsim_case as rc

    return rc.case()


def check_live(c, got, bitwise, tol_jm=TOL_JM):
    ref, t = c["ref"], c["touched"]
    assert max(ref[k][1] for k in ("J0", "M0", "J1", "M1")) == 0.0  # the copies of a corner the reference holds agree
    assert t.all()
    a = ref["after"]
    assert int((got["cells"] != a["cells"]).sum()) == 0
    if bitwise:
        assert int((got["x"] != a["x"]).sum()) == 0 and int((got["v"] != a["v"]).sum()) == 0
    else:
        nx, nv = np.sqrt((a["x"] ** 2).sum(axis=0)), np.sqrt((a["v"] ** 2).sum(axis=0))
        assert (np.abs(got["x"] - a["x"]).max(axis=0) / nx).max() <= 1e-10 and (np.abs(got["v"] - a["v"]).max(axis=0) / nv).max() <= 1e-10
    for k in ("J0", "M0", "J1", "M1"):
        assert rel(got[k], ref[k][0]) <= tol_jm, (k, rel(got[k], ref[k][0]))
    assert abs(got["energy"][0] - ref["energy0"]) <= 1e-12 * ref["energy0"] and abs(got["energy"][1] - ref["energy1"]) <= 1e-12 * ref["energy1"]
    assert got["stats"]["n_moved"] == a["x"].shape[1] and got["stats"]["n_periodic_wrap"] > 1000 and got["stats"]["n_cross_block"] > 1000


@needs_ref
def test_oracle_reproduces_the_reference_compiled_here(live):
    got = oracle_phase(live["mesh"], live["cfg"], live["parts"], live["fields"])
    check_live(live, got, bitwise=True)


@needs_ref
def test_oracle_cell_sampling_reproduces_the_reference(live):
    """PIC::Sampling::SamplingManager -> ProcessCell (pic.cpp:1049, :705) of the reference build on the moved fast-wave plasma: weight,
    particle number, number density, velocity, velocity^2 and speed sums of every real cell and species in the collecting buffer
    (the velocity tensor is off in this configuration)"""
    from oracle.oracle_py import Oracle

    r = live["refpic"]
    if not hasattr(r.lib, "ref_pic_sample_cells"):
        pytest.skip("oracle/_ref/libref_pic.so predates ref_pic_sample_cells (rebuild: make -C oracle)")
    ref, cnt = r.sample_cells()       # the collecting buffer after one more sample ...
    ref2, cnt2 = r.sample_cells()     # ... and after a second one: the difference is exactly one sample
    one = ref2 - ref
    m, cfg = live["mesh"], live["cfg"]
    x, v, w, sp, cells0 = live["parts"]
    a = live["ref"]["after"]
    o = Oracle(cfg, m)
    o.add_particles(a["x"], a["v"], w, sp, a["cells"].astype(np.int32))
    got, n_sampled = o.sample_cells()
    o.close()
    b2l, real = live["maps"]["b2l"], live["maps"]["real"]
    C = m.cells_per_block
    got = got.reshape(m.n_leaves, C, cfg.n_species, 13)
    assert int(cnt2.sum()) == x.shape[1] and list(n_sampled[: cfg.n_species]) == list(cnt2)
    worst = 0.0
    for b in real:
        for q in range(10):
            d = np.abs(got[b2l[b], :, :, q] - one[b, :, :, q]).max()
            worst = max(worst, d / max(np.abs(one[:, :, :, q]).max(), 1e-300))
    assert worst <= 1e-14, worst
    assert np.abs(one[real][..., 1].sum() - x.shape[1]) == 0  # particle numbers are exact


@needs_ref
def test_oracle_ecsim_field_getters_reproduce_the_reference(live):
    """ECSIM::GetElectricField / GetMagneticField / GetMagneticFieldGradient (pic_field_solver_ecsim.cpp:7440-7547), what the
    guiding-centre movers read when the field solver is ECSIM (cfg.gc_fields_ecsim): the oracle's restatement against the reference's own
    functions at 6000 points of the fast-wave box, some of them on block and cell faces"""
    from oracle.oracle_py import Oracle

    gt = live["ref"]["field"].get("getters")
    if gt is None:
        pytest.skip("oracle/_ref/libref_pic.so predates ref_pic_ecsim_fields (rebuild: make -C oracle)")
    E_u, Bp_u, Bc_u = live["fields"]
    o = Oracle(live["cfg"], live["mesh"])
    o.set_fields(E_u, Bp_u, gt["B_u"])
    o.set_E_current(gt["E_u"])
    E, B, G, bad = o.ecsim_fields(gt["x"], gt["leaf"])
    o.close()
    assert bad == 0
    assert (E == gt["E"]).all() and (B == gt["B"]).all() and (G == gt["gradB"]).all()  # bit for bit


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("exact", [True, False])
def test_gpu_reproduces_the_reference_compiled_here(live, exact):
    got = gpu_phase(live["mesh"], live["cfg"], live["parts"], live["fields"], exact)
    check_live(live, got, bitwise=exact, tol_jm=1e-10 if not exact else 1e-13)


# ---------------- the committed vectors of the reference ----------------
def gold_case():
    z = np.load(GOLD)
    m = meshmod.uniform_periodic_box(tuple(int(c) for c in z["n_cells"]), tuple(int(c) for c in z["block_cells"]),
                                     tuple(int(c) for c in z["ghost_cells"]), dx=1.0, origin=tuple(z["origin"]))
    n = z["x"].shape[1]
    cfg = api.make_config(tuple(int(c) for c in z["block_cells"]), tuple(int(c) for c in z["ghost_cells"]), tuple(z["charge"]), tuple(z["mass"]),
                          tuple(z["species_weight"]), float(z["dt"]), periodic=True, capacity=n + 16, B_conv=float(z["unit"][0]),
                          length_conv=float(z["unit"][1]), light_speed=float(z["unit"][2]))
    return z, m, cfg, (z["x"], z["v"], z["w"], z["species"], z["cells"]), (z["E_half"], z["B_prev"], z["B_cur"])


def check_gold(z, got, bitwise, tol_jm=TOL_JM):
    assert int((got["cells"] != z["cells_after"]).sum()) == 0
    if bitwise:
        assert int((got["x"] != z["x_after"]).sum()) == 0 and int((got["v"] != z["v_after"]).sum()) == 0
    else:
        nx, nv = np.sqrt((z["x_after"] ** 2).sum(axis=0)), np.sqrt((z["v_after"] ** 2).sum(axis=0))
        assert (np.abs(got["x"] - z["x_after"]).max(axis=0) / nx).max() <= 1e-10
        assert (np.abs(got["v"] - z["v_after"]).max(axis=0) / nv).max() <= 1e-10
    sub = z["M_corners"]
    for k in ("0", "1"):
        assert rel(got["J" + k], z["J" + k]) <= tol_jm
        assert rel(got["M" + k][sub], z["M" + k + "_sub"]) <= tol_jm
        assert rel(got["M" + k].sum(axis=1), z["M" + k + "_rowsum"]) <= 50 * tol_jm  # 243-term sums
    assert abs(got["energy"][0] - z["energy"][0]) <= 1e-12 * z["energy"][0] and abs(got["energy"][1] - z["energy"][1]) <= 1e-12 * z["energy"][1]


def test_oracle_reproduces_the_committed_reference_vectors():
    z, m, cfg, parts, fields = gold_case()
    check_gold(z, oracle_phase(m, cfg, parts, fields), bitwise=True)


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [True, False])
def test_gpu_reproduces_the_committed_reference_vectors(exact):
    z, m, cfg, parts, fields = gold_case()
    check_gold(z, gpu_phase(m, cfg, parts, fields, exact), bitwise=exact, tol_jm=1e-10 if not exact else 1e-13)
