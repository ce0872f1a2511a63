"""BASELINE configs[0] at its stated size: a cutoff-rigidity map of 150 x 150 arrival directions at 4 test points
(input/earth-cutoff-rigidity.input: BlockCells 4,4,4, GhostCells 1,1,1, backward time integration, exit through the user function =
Earth::CutoffRigidity::ProcessOutsideDomainParticles, srcEarth/CutoffRigidity.cpp:129-230: the lowest rigidity that escapes is the
cutoff of its direction).  Protons with 8 log-spaced rigidities per direction (0.5 - 20 GV) are traced backward with
PIC::Mover::Relativistic::Boris through the dipole tabulated on a 4-level AMR mesh: 720 000 trajectories on the GPU through the C ABI.

  * every 10th direction in both angles (15 x 15 x 4 directions, 7 200 trajectories) is also traced by the CPU oracle: the fate of
    every one of those trajectories (escaped / absorbed or trapped, exit face, exit velocity) is IDENTICAL, hence so is the cutoff
    of each of those directions (the movers are bit-exact, tests/test_relativistic_boris.py);
  * the whole map is checked statistically: the vertical cutoff of each point against Stormer's formula (the reference's C1 table
    accepts 5-35 %), the east-west asymmetry of a positive particle's cutoff, monotone decrease with latitude."""
import math
import os

import numpy as np
import pytest

from amps_b200 import _capi, api, mesh as meshmod, workload
from amps_b200.workload import B0, CLIGHT, MP, QP, RE
from oracle.oracle_py import Oracle

N_ZEN, N_AZ = 150, 150
POINTS = [(500.0, 0.0, 0.0), (500.0, 30.0, 90.0), (500.0, -45.0, 180.0), (500.0, 60.0, 270.0)]  # altitude km, latitude, longitude
RIG = np.exp(np.linspace(np.log(0.5), np.log(20.0), 8))  # GV
N_CALLS = 6000  # mover calls of dt = 5e-4 s: 3 s of flight


def build():
    L = 16.0 * RE

    def refine(level, lo, hi):
        near = np.clip(np.zeros(3), lo, hi)
        return float(np.linalg.norm(near)) / RE < (9.0, 5.0, 3.0)[level]

    m = meshmod.build_mesh((-L, -L, -L), (L, L, L), (8, 8, 8), (4, 4, 4), (1, 1, 1), periodic=False, refine=refine, max_level=3)
    xc = m.center_x
    r = np.sqrt((xc ** 2).sum(1))
    B = workload.dipole(np.where(r[:, None] < 0.5 * RE, xc + 0.5 * RE, xc))
    return m, (np.zeros_like(B), B)


def particles(m):
    """arrival directions on the upper hemisphere of every point: cos(zenith) uniform in (0, 1], azimuth uniform (0 = local east,
    90 = local north); index = ((point * N_ZEN + iz) * N_AZ + ia) * len(RIG) + ir"""
    cz = (np.arange(N_ZEN) + 0.5) / N_ZEN
    az = 2.0 * np.pi * (np.arange(N_AZ) + 0.5) / N_AZ
    xs, vs = [], []
    for alt, lat, lon in POINTS:
        lam, phi = math.radians(lat), math.radians(lon)
        up = np.array([math.cos(lam) * math.cos(phi), math.cos(lam) * math.sin(phi), math.sin(lam)])
        east = np.array([-math.sin(phi), math.cos(phi), 0.0])
        north = np.cross(up, east)
        pos = (RE + alt * 1e3) * up
        CZ, AZ = np.meshgrid(cz, az, indexing="ij")
        sz = np.sqrt(1.0 - CZ ** 2)
        # the direction the particle ARRIVES FROM; its velocity at the point is the opposite
        d = CZ[..., None] * up + (sz * np.cos(AZ))[..., None] * east + (sz * np.sin(AZ))[..., None] * north
        p = RIG * 1e9 * QP / CLIGHT
        speed = p / (np.sqrt(1.0 + (p / (MP * CLIGHT)) ** 2) * MP)
        v = -d[:, :, None, :] * speed[None, None, :, None]
        xs.append(np.broadcast_to(pos, v.shape).reshape(-1, 3))
        vs.append(v.reshape(-1, 3))
    x, v = np.concatenate(xs).T.copy(), np.concatenate(vs).T.copy()
    n = x.shape[1]
    return x, v, np.ones(n), np.zeros(n, dtype=np.uint8), workload.locate_cells(m, x)


def config(n):
    cfg = api.make_config((4, 4, 4), (1, 1, 1), (QP,), (MP,), (1.0,), 5.0e-4, periodic=False, capacity=n + 16, boundary_mode=_capi.BOUNDARY_USER_FUNCTION)
    cfg.time_step_mode = _capi.DT_SPECIES_GLOBAL
    cfg.coupler_interpolation = _capi.CPLR_LINEAR
    cfg.backward_time_integration = 1
    cfg.speed_of_light = CLIGHT
    cfg.internal_sphere_radius = RE
    cfg.exit_record_capacity = n
    return cfg


def fates(n, records):
    """per trajectory: exit face (-1: still inside after N_CALLS = trapped), exit velocity"""
    face = np.full(n, -1, dtype=np.int64)
    vout = np.zeros((n, 3))
    for ptr, spec, f, leaf, xx, vv in records:
        face[ptr], vout[ptr] = f, vv
    return face, vout


@pytest.mark.gpu
@pytest.mark.timeout(1500)
def test_cutoff_map_150x150x4_gpu_and_oracle_subsample():
    m, bg = build()
    parts = particles(m)
    n = parts[0].shape[1]
    assert n == len(POINTS) * N_ZEN * N_AZ * len(RIG)
    g = api.Context(config(n), m)
    g.background_upload(*bg)
    g.particles_upload(*parts)
    for it in range(N_CALLS):
        g.MoveParticles(_capi.MOVER_RELATIVISTIC_BORIS, stats=False)
        g.sort()
        if it % 200 == 199 and g.particle_count() == 0:
            break
    nrec, recs = g.exit_records(max_records=n)
    g.close()
    face, vout = fates(n, recs)
    allowed = (face >= 0) & (face != _capi.EXIT_SPHERE)
    A = allowed.reshape(len(POINTS), N_ZEN, N_AZ, len(RIG))
    # cutoff of a direction = the lowest sampled rigidity that escapes (inf: none)
    first = np.where(A.any(axis=3), A.argmax(axis=3), len(RIG))
    cutoff = np.where(first < len(RIG), RIG[np.minimum(first, len(RIG) - 1)], np.inf)

    # ---- the oracle on every 10th direction ----
    sel = np.zeros((len(POINTS), N_ZEN, N_AZ, len(RIG)), dtype=bool)
    sel[:, 4::10, 4::10, :] = True
    idx = np.nonzero(sel.reshape(-1))[0]
    sub = tuple(a[..., idx] if a.ndim > 1 else a[idx] for a in parts)
    o = Oracle(config(idx.size), m)
    o.set_background(*bg)
    o.add_particles(*sub)
    for it in range(N_CALLS):
        rc, st, ret, fc = o.move(_capi.MOVER_RELATIVISTIC_BORIS, os.cpu_count() or 1)
        assert rc == 0
        if (fc < 0).all():
            break
    nrec_o, recs_o = o.exit_records(max_records=idx.size)
    o.close()
    face_o, vout_o = fates(idx.size, recs_o)
    assert np.array_equal(face_o, face[idx]), int((face_o != face[idx]).sum())
    assert np.array_equal(vout_o, vout[idx])
    print("oracle subsample:", idx.size, "trajectories identical;", int((face_o >= 0).sum()), "ended,", int((face_o < 0).sum()), "trapped")

    # ---- statistics of the whole map ----
    R0 = 0.299792458 * 0.25 * B0 * RE  # GV, vertical Stormer cutoff at the equator on the surface
    vertical = cutoff[:, -1, :]        # the ring of directions closest to the zenith
    for ip, (alt, lat, lon) in enumerate(POINTS):
        rc = R0 * math.cos(math.radians(lat)) ** 4 / ((RE + alt * 1e3) / RE) ** 2
        med = float(np.median(vertical[ip][np.isfinite(vertical[ip])]))
        # the sampled rigidities are 1.69 apart: the first escaping sample lies in [Rc, 1.7 Rc] (+ the reference's 35 % band)
        print(f"point {ip}: lat {lat:+.0f}  Stormer {rc:.2f} GV  map (zenith ring median) {med:.2f} GV")
        assert 0.65 * rc <= med <= 1.7 * 1.35 * rc or med == RIG[0]
    # a proton from the west has the lower cutoff (east-west effect), at every point below 60 degrees
    low = cutoff[:, N_ZEN // 4, :]  # zenith angle ~ 75 degrees
    for ip, (alt, lat, lon) in enumerate(POINTS):
        if abs(lat) > 50:
            continue
        from_west = low[ip][(np.arange(N_AZ) > 0.375 * N_AZ) & (np.arange(N_AZ) < 0.625 * N_AZ)]   # azimuth ~ 180: arriving from the west
        from_east = np.concatenate([low[ip][: N_AZ // 8], low[ip][-N_AZ // 8:]])                  # azimuth ~ 0: from the east
        assert np.mean(np.minimum(from_west, 40.0)) < np.mean(np.minimum(from_east, 40.0))  # (means: the 8 sampled rigidities are 1.69 apart)
    # the vertical cutoff falls with |latitude|
    med_by_lat = sorted((abs(lat), float(np.median(np.minimum(vertical[ip], 40.0)))) for ip, (alt, lat, lon) in enumerate(POINTS))
    assert all(a[1] >= b[1] for a, b in zip(med_by_lat, med_by_lat[1:])), med_by_lat
