"""The guiding-centre branches of the ECSIM path against the reference's OWN code built with its gyrokinetic model on
(oracle/_ref/libref_pic_gk.so, REF_PIC_VARIANT=gk of oracle/ref_pic/build_ref_pic.sh; vectors committed in
tests/golden/ref_gyrokinetic.npz by tests/golden/make_ref_gyrokinetic.py: every 31st particle of the fast-wave box, electrons =
guiding-centre species):

  * ProcessCell with use_gc_species (pic_field_solver_ecsim.cpp:2084, :2205-2256, :2310, closure :1828 called :2376)  -> cfg.gc_species_mask
  * PIC::GYROKINETIC::Mover -> GuidingCenter::Mover_FirstOrder on ECSIM::GetElectricField / GetMagneticField / GetMagneticFieldGradient
    with InitiateMagneticMoment (pic_mover_guiding_center.cpp:103, :179-184, :629-849), Lapenta2017 for the ions  -> cfg.gc_fields_ecsim

The oracle is checked on the CPU, the kernels on the GPU, both against the same reference numbers."""
import os
import subprocess
import sys

import numpy as np
import pytest

from amps_b200 import _capi, api, mesh as meshmod
from oracle.oracle_py import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ref_gyrokinetic.npz")
LIB = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_pic_gk.so")
GC1 = _capi.MOVER_GC_FIRST_ORDER


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def gold_case():
    z = np.load(GOLD)
    N, g = tuple(int(c) for c in z["block_cells"]), tuple(int(c) for c in z["ghost_cells"])
    m = meshmod.uniform_periodic_box(tuple(int(c) for c in z["n_cells"]), N, g, dx=1.0, origin=tuple(z["origin"]))
    n = z["x"].shape[1]
    cfg = api.make_config(N, g, tuple(z["charge"]), tuple(z["mass"]), tuple(z["species_weight"]), float(z["dt"]), periodic=True, capacity=n + 16,
                          B_conv=float(z["unit"][0]), length_conv=float(z["unit"][1]), light_speed=float(z["unit"][2]))
    cfg.gc_species_mask = 1
    cfg.carry_magnetic_moment = 1
    cfg.gc_fields_ecsim = 1
    cfg.ideal_mhd = 1  # _PIC__IDEAL_MHD_MODE_ of the reference's configuration (picGlobal.dfn:339)
    return z, m, cfg


def check_deposit(z, k, J, M, energy, tol):
    sub = z["M_corners"]
    assert rel(J, z["J" + k]) <= tol, rel(J, z["J" + k])
    assert rel(M[sub], z["M" + k + "_sub"]) <= tol
    assert rel(M.sum(axis=1), z["M" + k + "_rowsum"]) <= 50 * tol
    assert abs(energy - z["energy"][int(k)]) <= 1e-12 * z["energy"][int(k)]


def test_oracle_gc_species_deposit_matches_the_reference():
    z, m, cfg = gold_case()
    for k, (x, v, cells, mu) in (("0", (z["x"], z["v"], z["cells"], z["mu0"])), ("1", (z["x_after"], z["v_after"], z["cells_after"], z["mu_after"]))):
        o = Oracle(cfg, m)
        o.set_fields(z["E_half"], z["B_prev"], z["B_cur"])
        o.add_particles(x, v, z["w"], z["species"], cells.astype(np.int32))
        o.set_reduced_state(mu, np.zeros_like(mu))
        o.set_v_normal(z["vnormal"])
        J, M, en, _ = o.deposit(1)
        o.close()
        check_deposit(z, k, J, M, en, 2e-14)


def mover_config(z, cfg, s):
    """GuidingCenter::Mover_FirstOrder / InitiateMagneticMoment read PIC::MolecularData::GetElectricCharge / GetMass, the raw species
    tables (pic_mover_guiding_center.cpp:137, :216-217, :643), where Lapenta2017 and ProcessCell convert them with picunits::si2no_*:
    the context that runs the guiding-centre species gets the raw tables"""
    if s == 0:
        for i in range(2):
            cfg.charge[i], cfg.mass[i] = float(z["charge_table"][i]), float(z["mass_table"][i])
    return cfg


def oracle_move(z, m, cfg, sel, mover, start=("x", "v", "cells", "mu0"), global_stencil=0):
    o = Oracle(cfg, m)
    o.set_global_stencil_length(global_stencil)
    o.set_fields(z["E_half"], z["B_prev"], z["B_cur"])
    o.set_E_current(z["E_cur"])
    o.add_particles(z[start[0]][:, sel], z[start[1]][:, sel], z["w"][sel], z["species"][sel], z[start[2]][sel].astype(np.int32))
    o.set_reduced_state(z[start[3]][sel], np.zeros(int(sel.sum())))
    rc, st, ret, fc = o.move(mover, 1)
    pp = o.particles()
    mu, flag = o.magnetic_moment()
    o.close()
    assert rc == 0
    return pp["x"], pp["v"], fc.astype(np.int64), mu, flag


def test_oracle_gyrokinetic_mover_matches_the_reference():
    z, m, cfg = gold_case()
    sp = z["species"]
    for s, mover in ((0, GC1), (1, _capi.MOVER_LAPENTA2017)):
        sel = sp == s
        z, m, cfg = gold_case()
        x, v, cells, mu, flag = oracle_move(z, m, mover_config(z, cfg, s), sel, mover)
        assert (cells == z["cells_after"][sel]).all()
        if s == 1:  # Lapenta2017: bit for bit, as in test_reference_ecsim.py
            assert (x == z["x_after"][:, sel]).all() and (v == z["v_after"][:, sel]).all()
        else:  # the guiding-centre mover: the reference writes |B| as pow(B.B, 0.5), the oracle as well; a few ulp from summation order
            nx, nv = np.abs(z["x_after"][:, sel]).max(), np.abs(z["v_after"][:, sel]).max()
            assert np.abs(x - z["x_after"][:, sel]).max() <= 1e-13 * nx
            assert np.abs(v - z["v_after"][:, sel]).max() <= 1e-12 * nv
            assert np.abs(mu - z["mu_after"][sel]).max() <= 1e-13 * np.abs(z["mu_after"][sel]).max()
            assert (flag == z["init_flag_after"][sel]).all()
            print("guiding-centre species: x words equal", int((x == z["x_after"][:, sel]).sum()), "of", x.size)


def dive_conv(z):
    """ComputeNetCharge and CorrectParticleLocation multiply the RAW species tables by charge_conv / mass_conv (:4704, :4449-4452), the
    sampled moments use ProcessCell's si2no masses: with the si2no tables in the configuration the two factors carry the ratio"""
    return float(z["conv"][0] * z["charge_table"][0] / z["charge"][0]), float(z["conv"][1] * z["mass_table"][0] / z["mass"][0])


def test_oracle_second_order_guiding_centre_mover_matches_the_reference():
    """GuidingCenter::Mover_SecondOrder (pic_mover_guiding_center.cpp:292-619) on ECSIM's fields, from the state the div-E correction left"""
    sp = np.load(GOLD)["species"]
    for s, mover in ((0, _capi.MOVER_GC_SECOND_ORDER), (1, _capi.MOVER_LAPENTA2017)):
        sel = sp == s
        z, m, cfg = gold_case()
        # ComputeNetCharge ran before this move: the reference's global StencilTable holds an 8-cell stencil, so the B stencils of
        # Lapenta2017 and of ECSIM::GetMagneticField are no longer normalised (pic_interpolation_routines.cpp:903)
        x, v, cells, mu, flag = oracle_move(z, m, mover_config(z, cfg, s), sel, mover, start=("x_corrected", "v_after", "cells_corrected", "mu_after"),
                                            global_stencil=8)
        assert (cells == z["cells_second"][sel]).all()
        if s == 1:
            assert (x == z["x_second"][:, sel]).all() and (v == z["v_second"][:, sel]).all()
        else:
            nx, nv = np.abs(z["x_second"][:, sel]).max(), np.abs(z["v_second"][:, sel]).max()
            assert np.abs(x - z["x_second"][:, sel]).max() <= 1e-13 * nx
            assert np.abs(v - z["v_second"][:, sel]).max() <= 1e-12 * nv
            assert np.abs(mu - z["mu_second"][sel]).max() <= 1e-13 * np.abs(z["mu_second"][sel]).max()
            print("second order: x words equal", int((x == z["x_second"][:, sel]).sum()), "of", x.size)


def inner_cells(z, m):
    """the centres that are not in the outermost layer of the periodic box: there the reference's ComputeNetCharge leaves the contributions
    of the periodic images in its ghost blocks and SetBoundaryChargeDivE (:4960-5009) zeroes the cell afterwards, while the library's one
    unique centre per periodic image holds the folded sum"""
    lo = np.asarray(z["origin"], dtype=np.float64)
    hi = lo + np.asarray(z["n_cells"], dtype=np.float64)
    xc = np.asarray(m.center_x)
    return ((xc > lo + 1.0) & (xc < hi - 1.0)).all(axis=1)


def test_oracle_dive_correction_passes_match_the_reference():
    """ECSIM::ComputeNetCharge (:4690), the corner species moments of ProcessCell (_PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_, :2270-2300)
    and CorrectParticleLocation (:4440-4688) of the same reference build, on the plasma after the move"""
    z, m, cfg = gold_case()
    cc, mc = dive_conv(z)
    o = Oracle(cfg, m)
    o.set_fields(z["E_half"], z["B_prev"], z["B_cur"])
    o.add_particles(z["x_after"], z["v_after"], z["w"], z["species"], z["cells_after"].astype(np.int32))
    rho = o.net_charge(cc)
    inner = inner_cells(z, m)
    assert inner.mean() > 0.6 and rel(rho[inner], z["net_charge"][inner]) <= 1e-13
    mom = o.species_moments()
    assert float(z["species_moments_spread"]) == 0.0
    for s in range(2):
        for k in range(10):
            assert rel(mom[:, s, k], z["species_moments"][:, s, k]) <= 1e-13, (s, k)
    o.set_phi(z["phi"])
    rc, n_disp, n_del, fc = o.correct_particle_location(cc, mc)
    after = o.particles()
    o.close()
    assert rc == 0 and n_del == 0 and n_disp == int((z["species"] == 0).sum())
    moved = np.abs(z["x_corrected"] - z["x_after"]).max(axis=0) > 0
    assert moved.sum() == n_disp
    # the shift is a ratio of interpolated sums: it inherits their summation order
    assert np.abs(after["x"] - z["x_corrected"]).max() <= 1e-12
    inside = z["cells_corrected"] >= 0
    assert (fc[inside] == z["cells_corrected"][inside]).all()


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_pic_gk.so not built (REF_PIC_VARIANT=gk, needs /root/reference)")
def test_committed_vectors_are_what_the_reference_library_produces(tmp_path):
    out = str(tmp_path / "gk.npz")
    env = dict(os.environ, AMPS_REF_PIC_LIB=LIB)
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_ref_gyrokinetic.py"), out], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    a, b = np.load(out), np.load(GOLD)
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
def test_gpu_gc_species_deposit_and_gyrokinetic_mover_match_the_reference():
    z, m, cfg = gold_case()
    sp = z["species"]
    n = z["x"].shape[1]
    # deposit before and after the move
    for k, (x, v, cells, mu) in (("0", (z["x"], z["v"], z["cells"], z["mu0"])), ("1", (z["x_after"], z["v_after"], z["cells_after"], z["mu_after"]))):
        g = api.Context(cfg, m)
        g.fields_upload(z["E_half"], z["B_prev"], z["B_cur"])
        g.particles_upload(x, v, z["w"], sp, cells.astype(np.int32))
        g.magnetic_moment_upload(mu)
        g.v_normal_upload(z["vnormal"])
        g.sort()
        en, _ = g.UpdateJMassMatrix()
        J, M = g.JM_download()
        g.close()
        check_deposit(z, k, J, M, en, 1e-10)
    # the movers, species by species as PIC::GYROKINETIC::Mover routes them
    for s, mover in ((0, GC1), (1, _capi.MOVER_LAPENTA2017)):
        sel = sp == s
        ns = int(sel.sum())
        z, m, cfg = gold_case()
        cfg = mover_config(z, cfg, s)
        cfg.exact_arithmetic = 1
        g = api.Context(cfg, m)
        g.fields_upload(z["E_half"], z["B_prev"], z["B_cur"])
        g.E_upload(z["E_cur"])
        g.particles_upload(z["x"][:, sel], z["v"][:, sel], z["w"][sel], sp[sel], z["cells"][sel])
        g.magnetic_moment_upload(z["mu0"][sel])
        st = g.MoveParticles(mover)
        mv = g.particles_download()
        mu_dev = g.magnetic_moment_download()
        g.close()
        gx, gv, gc, gmu = np.empty((3, ns)), np.empty((3, ns)), np.empty(ns, dtype=np.int64), np.empty(ns)
        gx[:, mv["ptrs"]], gv[:, mv["ptrs"]], gc[mv["ptrs"]], gmu[mv["ptrs"]] = mv["x"], mv["v"], mv["cells"], mu_dev
        assert (gc == z["cells_after"][sel]).all()
        if s == 1:
            assert (gx == z["x_after"][:, sel]).all() and (gv == z["v_after"][:, sel]).all()
        else:
            nx, nv = np.abs(z["x_after"][:, sel]).max(), np.abs(z["v_after"][:, sel]).max()
            assert np.abs(gx - z["x_after"][:, sel]).max() <= 1e-12 * nx
            assert np.abs(gv - z["v_after"][:, sel]).max() <= 1e-10 * nv
            assert np.abs(gmu - z["mu_after"][sel]).max() <= 1e-12 * np.abs(z["mu_after"][sel]).max()
    # the second move (after ComputeNetCharge): second-order guiding centre for the electrons, Lapenta2017 for the ions.  The reference
    # stopped normalising full B stencils once ComputeNetCharge had filled its global StencilTable (pic_interpolation_routines.cpp:903,
    # see the oracle): amps_gpu_global_stencil_set puts the kernels into the same state
    for s, mover in ((0, _capi.MOVER_GC_SECOND_ORDER), (1, _capi.MOVER_LAPENTA2017)):
        sel = sp == s
        ns = int(sel.sum())
        z, m, cfg = gold_case()
        cfg = mover_config(z, cfg, s)
        cfg.exact_arithmetic = 1
        g = api.Context(cfg, m)
        g.fields_upload(z["E_half"], z["B_prev"], z["B_cur"])
        g.E_upload(z["E_cur"])
        g.particles_upload(z["x_corrected"][:, sel], z["v_after"][:, sel], z["w"][sel], sp[sel], z["cells_corrected"][sel].astype(np.int32))
        g.magnetic_moment_upload(z["mu_after"][sel])
        g.global_stencil_set(True)
        g.MoveParticles(mover)
        mv = g.particles_download()
        g.close()
        gx, gv, gc = np.empty((3, ns)), np.empty((3, ns)), np.empty(ns, dtype=np.int64)
        gx[:, mv["ptrs"]], gv[:, mv["ptrs"]], gc[mv["ptrs"]] = mv["x"], mv["v"], mv["cells"]
        assert (gc == z["cells_second"][sel]).all()
        nx, nv = np.abs(z["x_second"][:, sel]).max(), np.abs(z["v_second"][:, sel]).max()
        if s == 1:  # the exact Lapenta2017 kernel in the reference's post-ComputeNetCharge state: bit for bit
            assert (gx == z["x_second"][:, sel]).all() and (gv == z["v_second"][:, sel]).all()
        else:
            assert np.abs(gx - z["x_second"][:, sel]).max() <= 1e-12 * nx and np.abs(gv - z["v_second"][:, sel]).max() <= 1e-10 * nv
    # the particle passes of the div-E correction on the moved plasma
    z, m, cfg = gold_case()
    cc, mc = dive_conv(z)
    g = api.Context(cfg, m)
    g.fields_upload(z["E_half"], z["B_prev"], z["B_cur"])
    g.particles_upload(z["x_after"], z["v_after"], z["w"], sp, z["cells_after"].astype(np.int32))
    g.sort()
    rho = g.ComputeNetCharge(cc)
    inner = inner_cells(z, m)
    assert rel(rho[inner], z["net_charge"][inner]) <= 1e-10
    mom = g.ComputeSpeciesMoments()
    for s in range(2):
        for k in range(10):
            assert rel(mom[:, s, k], z["species_moments"][:, s, k]) <= 1e-10, (s, k)
    g.SetPhi(z["phi"])
    nd, nx = g.CorrectParticleLocation(cc, mc)
    got = g.particles_download()
    g.close()
    assert nx == 0 and nd == int((sp == 0).sum())
    gx = np.empty((3, n))
    gx[:, got["ptrs"]] = got["x"]
    assert np.abs(gx - z["x_corrected"]).max() <= 1e-10
