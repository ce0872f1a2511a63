"""Known-answer test of the reference for the test-particle path: the vertical Størmer cutoff of a centred dipole,
Rc = R0 cos^4(lambda) / r^2, srcEarth/test/C1 (tests/golden/stormer_tables.json holds that test's table, written by
tests/golden/make_stormer_tables.py from the reference tree).
Protons are traced backward in time with PIC::Mover::Relativistic::Boris (a7) through the dipole tabulated on an AMR mesh, like
the reference's Mode3D MESH variant: a vertical arrival 1.6 x above the cutoff connects to the outer boundary, one 0.6 x below
does not (the reference accepts 5-35 % around Rc, run_C1.py:322-337).  srcEarth/test/C4/reference_C4_invariants.csv (same fixture
file) lists, for factors 0.5 and 2, whether the trajectory must reach the outer box, and bounds the rigidity change
along it by 1e-6 (E = 0: the magnetic force does no work); both are asserted for every row.
srcEarth/test/C2/reference_C2_stormer_symmetry.csv (the cutoff of a centred dipole does not depend on the longitude: 12 longitudes per
latitude) and the BORIS rows of srcEarth/test/C12/reference_C12_stormer_movers.csv (two shells, latitudes every 20 degrees) are
bracketed the same way by the oracle."""
import json
import math
import os

import numpy as np
import pytest

from amps_b200 import _capi, api, mesh as meshmod, workload
from amps_b200.workload import B0, CLIGHT, MP, QP, RE
from oracle.oracle_py import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def _fixture():
    with open(os.path.join(HERE, "golden", "stormer_tables.json")) as f:
        return json.load(f)


def _table():
    return [(r["alt_km"], r["lat_deg"], r["Rc_stormer_GV"]) for r in _fixture()["C1"]]


def test_table_is_the_stormer_formula():
    # run_C1.py:43-46,105-106: R0 = 0.299792458 * 0.25 * B_eq * Re with B_eq = 3.12e-5 T, Re = 6371.2 km
    R0 = 0.299792458 * 0.25 * 3.12e-5 * (6371.2 * 1000.0)
    for alt, lat, rc in _table():
        r_re = (6371.2 + alt) / 6371.2
        assert abs(R0 * math.cos(math.radians(lat)) ** 4 / r_re ** 2 - rc) <= 1e-9 * rc


def _table_c4():
    return [(r["alt_km"], r["lat_deg"], r["factor"], r["R_GV"], r["Rc_stormer_GV"], r["expected_allowed"], r["rel_dR_limit"]) for r in _fixture()["C4"]]


def test_c4_table_is_consistent_with_c1():
    c1 = {(a, l): rc for a, l, rc in _table()}
    for alt, lat, factor, R, rc, allowed, lim in _table_c4():
        assert abs(rc - c1[(alt, lat)]) <= 1e-9 * rc and abs(R - factor * rc) <= 1e-9 * R and allowed == (factor > 1.0)


def _case(rows, factors):
    """one proton per (table row, factor): launched at the row's point, arriving vertically with rigidity factor * Rc of OUR dipole
    (B_eq = workload.B0 at workload.RE; the cutoff scales linearly with B_eq Re)"""
    L = 16.0 * RE

    def refine(level, lo, hi):  # dx = 1, 0.5, 0.25, 0.125 Re towards the planet (2:1 balanced)
        near = np.clip(np.zeros(3), lo, hi)
        return float(np.linalg.norm(near)) / RE < (9.0, 5.0, 3.0)[level]

    m = meshmod.build_mesh((-L, -L, -L), (L, L, L), (8, 8, 8), (4, 4, 4), (1, 1, 1), periodic=False, refine=refine, max_level=3)
    xc = m.center_x
    r = np.sqrt((xc ** 2).sum(1))
    B = workload.dipole(np.where(r[:, None] < 0.5 * RE, xc + 0.5 * RE, xc))
    E = np.zeros_like(B)
    R0 = 0.299792458 * 0.25 * B0 * RE  # GV
    xs, vs, expect, rig = [], [], [], []
    for row in rows:  # (alt_km, lat_deg, Rc) or (alt_km, lat_deg, Rc, lon_deg)
        alt, lat = row[0], row[1]
        phi = math.radians(row[3]) if len(row) > 3 else 0.0
        rr = RE + alt * 1e3
        lam = math.radians(lat)
        pos = rr * np.array([math.cos(lam) * math.cos(phi), math.cos(lam) * math.sin(phi), math.sin(lam)])
        rc = R0 * math.cos(lam) ** 4 / (rr / RE) ** 2
        for f in factors:
            p = f * rc * 1e9 * QP / CLIGHT
            gamma = math.sqrt(1.0 + (p / (MP * CLIGHT)) ** 2)
            xs.append(pos)
            vs.append(-pos / rr * (p / (gamma * MP)))  # arrival velocity: vertically down
            expect.append(f > 1.0)
            rig.append(f * rc)
    x, v = np.array(xs).T.copy(), np.array(vs).T.copy()
    n = x.shape[1]
    cells = workload.locate_cells(m, x)
    cfg = api.make_config((4, 4, 4), (1, 1, 1), (QP,), (MP,), (1.0,), 5.0e-4, periodic=False, capacity=n + 16, boundary_mode=_capi.BOUNDARY_USER_FUNCTION)
    cfg.time_step_mode = _capi.DT_SPECIES_GLOBAL
    cfg.coupler_interpolation = _capi.CPLR_LINEAR
    cfg.backward_time_integration = 1
    cfg.speed_of_light = CLIGHT
    cfg.internal_sphere_radius = RE
    cfg.exit_record_capacity = n
    return m, cfg, (x, v, np.ones(n), np.zeros(n, dtype=np.uint8), cells), (E, B), np.array(expect), np.array(rig)


def _classify(n, records):
    """True: left through the outer boundary (allowed); False: hit the planet or still inside (forbidden)"""
    out = np.zeros(n, dtype=bool)
    for ptr, spec, face, leaf, xx, vv in records:
        if face != _capi.EXIT_SPHERE:
            out[ptr] = True
    return out


def _rigidity_gv(v):
    """rigidity of a proton with velocity v: p c / q in GV"""
    v = np.asarray(v)
    b2 = float((v ** 2).sum()) / CLIGHT ** 2
    return MP * math.sqrt(float((v ** 2).sum())) / math.sqrt(1.0 - b2) * CLIGHT / QP * 1e-9


def _check_c4(recs, expect, rig, lim):
    got = _classify(len(expect), recs)
    assert (got == expect).all(), (got, expect)
    seen = 0
    for ptr, spec, face, leaf, xx, vv in recs:  # every trajectory that ended (outer box or planet): |rel_dR| <= rel_dR_limit
        assert abs(_rigidity_gv(vv) - rig[ptr]) <= lim * rig[ptr], (ptr, _rigidity_gv(vv), rig[ptr])
        seen += 1
    assert seen >= int(expect.sum())


N_STEPS = 4000  # 2 s of flight at dt = 5e-4 s (150 km per step at the speed of light): a path of 94 R_E


def test_oracle_brackets_the_vertical_cutoff():
    rows = [r for r in _table() if r[0] == 9000.0 and abs(r[1]) <= 30.0]
    m, cfg, parts, bg, expect, rig = _case(rows, (0.6, 1.6))
    o = Oracle(cfg, m)
    o.set_background(*bg)
    o.add_particles(*parts)
    for it in range(N_STEPS):
        rc, st, ret, fc = o.move(_capi.MOVER_RELATIVISTIC_BORIS, 1)
        assert rc == 0
        if (fc < 0).all():
            break
    nrec, recs = o.exit_records()
    o.close()
    got = _classify(len(expect), recs)
    assert (got == expect).all(), (got, expect)


@pytest.mark.gpu
def test_gpu_brackets_the_vertical_cutoff():
    rows = _table()
    m, cfg, parts, bg, expect, rig = _case(rows, (0.6, 1.6))
    g = api.Context(cfg, m)
    g.background_upload(*bg)
    g.particles_upload(*parts)
    for it in range(4 * N_STEPS):  # (the 0.26 GV protons of the 60 degree rows move at a quarter of the speed of light)
        g.MoveParticles(_capi.MOVER_RELATIVISTIC_BORIS, stats=False)
        g.sort()
        if it % 100 == 99 and g.particle_count() == 0:
            break
    nrec, recs = g.exit_records()
    g.close()
    got = _classify(len(expect), recs)
    assert (got == expect).all(), (got, expect)


def _c4_case():
    tab = _table_c4()
    rows = sorted({(alt, lat, rc) for alt, lat, f, R, rc, a, lim in tab})
    m, cfg, parts, bg, expect, rig = _case(rows, (0.5, 2.0))
    # our dipole is B_eq = 3.1e-5 T at 6371 km instead of 3.12e-5 T at 6371.2 km: the table's R_GV scale by the same 0.6 %
    want = {(alt, lat, f): bool(a) for alt, lat, f, R, rc, a, lim in tab}
    k = 0
    for alt, lat, rc in rows:
        for f in (0.5, 2.0):
            assert expect[k] == want[(alt, lat, f)]
            k += 1
    return m, cfg, parts, bg, expect, rig, max(r[6] for r in tab)


def test_oracle_c4_exits_and_rigidity_conservation():
    m, cfg, parts, bg, expect, rig, lim = _c4_case()
    keep = np.array([abs(math.degrees(math.asin(parts[0][2, i] / np.linalg.norm(parts[0][:, i])))) <= 31.0 for i in range(len(expect))])
    sub = tuple(a[..., keep] if a.ndim > 1 else a[keep] for a in parts)  # the fast rows keep the CPU suite short
    o = Oracle(cfg, m)
    o.set_background(*bg)
    o.add_particles(*sub)
    for it in range(N_STEPS):
        rc, st, ret, fc = o.move(_capi.MOVER_RELATIVISTIC_BORIS, 1)
        assert rc == 0
        if (fc < 0).all():
            break
    nrec, recs = o.exit_records()
    o.close()
    _check_c4(recs, expect[keep], rig[keep], lim)


@pytest.mark.gpu
def test_gpu_c4_exits_and_rigidity_conservation():
    m, cfg, parts, bg, expect, rig, lim = _c4_case()
    g = api.Context(cfg, m)
    g.background_upload(*bg)
    g.particles_upload(*parts)
    for it in range(6 * N_STEPS):
        g.MoveParticles(_capi.MOVER_RELATIVISTIC_BORIS, stats=False)
        g.sort()
        if it % 100 == 99 and g.particle_count() == 0:
            break
    nrec, recs = g.exit_records()
    g.close()
    _check_c4(recs, expect, rig, lim)


def _run_oracle_bracket(rows, n_steps):
    m, cfg, parts, bg, expect, rig = _case(rows, (0.6, 1.6))
    o = Oracle(cfg, m)
    o.set_background(*bg)
    o.add_particles(*parts)
    for it in range(n_steps):
        rc, st, ret, fc = o.move(_capi.MOVER_RELATIVISTIC_BORIS, 1)
        assert rc == 0
        if (fc < 0).all():
            break
    nrec, recs = o.exit_records()
    o.close()
    return _classify(len(expect), recs), expect


def test_c2_and_c12_tables_are_the_stormer_formula():
    fx = _fixture()
    R0 = 0.299792458 * 0.25 * 3.12e-5 * (6371.2 * 1000.0)
    assert len(fx["C2"]) == 60 and len(fx["C12"]) == 14
    for r in fx["C2"] + fx["C12"]:
        r_re = (6371.2 + r["alt_km"]) / 6371.2
        assert abs(R0 * math.cos(math.radians(r["lat_deg"])) ** 4 / r_re ** 2 - r["Rc_stormer_GV"]) <= 1e-9 * r["Rc_stormer_GV"]
    by_lat = {}
    for r in fx["C2"]:  # run_C2.py: one cutoff per latitude whatever the longitude
        by_lat.setdefault(r["lat_deg"], set()).add(r["Rc_stormer_GV"])
    assert all(len(v) == 1 for v in by_lat.values()) and len(by_lat) == 5


def test_oracle_cutoff_is_independent_of_longitude():
    # C2: the twelve longitudes of the equatorial and the +-30 degree rows (the slow 60 degree rows stay with the GPU test above)
    rows = [(r["alt_km"], r["lat_deg"], r["Rc_stormer_GV"], r["lon_deg"]) for r in _fixture()["C2"] if abs(r["lat_deg"]) <= 30.0]
    assert len(rows) == 36
    got, expect = _run_oracle_bracket(rows, N_STEPS)
    assert (got == expect).all(), (got, expect)


def test_oracle_brackets_the_c12_boris_rows():
    # C12, BORIS: 9000 km and 25000 km shells, latitudes 0, +-20, +-40 (new latitudes and a second shell compared with C1)
    rows = [(r["alt_km"], r["lat_deg"], r["Rc_stormer_GV"]) for r in _fixture()["C12"] if abs(r["lat_deg"]) <= 40.0]
    assert len(rows) == 10
    got, expect = _run_oracle_bracket(rows, 3 * N_STEPS)
    assert (got == expect).all(), (got, expect)


def _run_gpu_bracket(rows, n_steps):
    m, cfg, parts, bg, expect, rig = _case(rows, (0.6, 1.6))
    g = api.Context(cfg, m)
    g.background_upload(*bg)
    g.particles_upload(*parts)
    for it in range(n_steps):
        g.MoveParticles(_capi.MOVER_RELATIVISTIC_BORIS, stats=False)
        g.sort()
        if it % 100 == 99 and g.particle_count() == 0:
            break
    nrec, recs = g.exit_records()
    g.close()
    return _classify(len(expect), recs), expect


@pytest.mark.gpu
def test_gpu_cutoff_is_independent_of_longitude():
    rows = [(r["alt_km"], r["lat_deg"], r["Rc_stormer_GV"], r["lon_deg"]) for r in _fixture()["C2"] if abs(r["lat_deg"]) <= 30.0]
    got, expect = _run_gpu_bracket(rows, N_STEPS)
    assert (got == expect).all(), (got, expect)


@pytest.mark.gpu
def test_gpu_brackets_the_c12_boris_rows():
    rows = [(r["alt_km"], r["lat_deg"], r["Rc_stormer_GV"]) for r in _fixture()["C12"] if abs(r["lat_deg"]) <= 40.0]
    got, expect = _run_gpu_bracket(rows, 3 * N_STEPS)
    assert (got == expect).all(), (got, expect)
