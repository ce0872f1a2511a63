"""Writes tests/golden/ref_gyrokinetic.npz: the reference's OWN PIC core built with its gyrokinetic model switched on
(oracle/_ref/libref_pic_gk.so: REF_PIC_VARIANT=gk of oracle/ref_pic/build_ref_pic.sh -- the fast-wave configuration with
_PIC_GYROKINETIC_MODEL_MODE_ and _USE_MAGNETIC_MOMENT_ on and PIC::GYROKINETIC::Mover as the particle mover), run on every 31st particle
of the fast-wave initial condition with species 0 (the electrons) as the guiding-centre species:

  * ECSIM::UpdateJMassMatrix with ProcessCell's use_gc_species branch (q v_eff, no mass matrix, the magnetisation current, v_normal in
    the energy) for given mu and v_normal,
  * PIC::Mover::MoveParticles -> PIC::GYROKINETIC::Mover: GuidingCenter::Mover_FirstOrder on ECSIM's own E, B, grad B for species 0
    (InitiateMagneticMoment included: the InitFlag of the particles is off), Lapenta2017 for species 1, then the periodic exchange,
  * UpdateJMassMatrix again with the magnetic moments the mover left,
  * ComputeNetCharge, the corner species moments and CorrectParticleLocation of the div-E correction on the moved plasma,
  * one more MoveParticles with GuidingCenter::Mover_SecondOrder for species 0.

Run it in its own process (the library's state is global and must not share a process with libref_pic.so):
    AMPS_REF_PIC_LIB=oracle/_ref/libref_pic_gk.so python tests/golden/make_ref_gyrokinetic.py [out.npz]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("AMPS_REF_PIC_LIB", os.path.join(ROOT, "oracle", "_ref", "libref_pic_gk.so"))
from tests import ref_ecsim_case as rc  # noqa: E402


def prepare(r, p0, info):
    assert r.gyrokinetic(), "this script needs the gyrokinetic variant of the reference library"
    r.set_gc_species(0, True)
    r.set_gc_species(1, False)
    m = info["mesh"]
    E_cur = info["smooth"](m.corner_x, 0.015, 1.7) - 0.004  # the current E (slot 0 of the corner data): what ECSIM::GetElectricField reads
    r.set_corner(0, E_cur[info["cu"]])
    n = p0["ptr"].size
    rng = np.random.default_rng(23)
    mu = rng.uniform(0.5, 2.0, n) * 1.0e-4
    vn = rng.uniform(0.0, 1.0, n) * 0.02
    r.set_reduced(p0["ptr"], mu=mu, vnormal=vn, init_flag=np.zeros(n, dtype=np.int32))
    return {"E_cur": E_cur, "mu0": mu, "vnormal": vn}


c = rc.case(keep_every=31, prepare=prepare, do_field=False)  # an odd stride: the particle list alternates the two species
ref, m, cfg, r = c["ref"], c["mesh"], c["cfg"], c["refpic"]
x, v, w, sp, cells = c["parts"]
E, Bp, Bc = c["fields"]
mu1, vn1, flag1 = r.get_reduced(c["ptr0"])  # periodic box: every particle keeps its ParticleBuffer slot
# ---- the particle passes of ECSIM::divECorrection on the moved plasma (the library samples the species moments on the corners) ----
assert r.samples_species_on_corners()
mp = c["maps"]
N, g = r.N, r.g
zu, real, b2l = mp["zu"], mp["real"], mp["b2l"]


def center_scalar_unique(arr):
    out = np.full(m.n_centers, np.nan)
    for b in real:
        sl = (b, slice(g[2], g[2] + N[2]), slice(g[1], g[1] + N[1]), slice(g[0], g[0] + N[0]))
        out[zu[sl].reshape(-1)] = arr[sl].reshape(-1)
    assert not np.isnan(out).any()
    return out


rho_u = center_scalar_unique(r.compute_net_charge(False))       # ComputeNetCharge(false)
mom = r.species_moments()                                        # left on the corners by the last UpdateJMassMatrix
mom_u, mom_spread = mp["to_unique"](mom.reshape(mom.shape[:-2] + (-1,)))
xc = m.center_x
L = np.array([32.0, 16.0, 8.0])
kk = 2 * np.pi / L
phi_u = 2.0e-4 * (np.sin(kk[0] * xc[:, 0]) * np.cos(kk[1] * xc[:, 1]) + 0.5 * np.sin(2 * kk[2] * xc[:, 2] + 0.3))
r.set_center_scalar(9, phi_u[zu])                                # phi of the Poisson solve, every copy of a centre node
r.correct_particle_location()                                    # CorrectParticleLocation + ExchangeParticleData
p2 = r.particles()
order = np.argsort(p2["ptr"])
pos = np.searchsorted(p2["ptr"][order], c["ptr0"])
assert (p2["ptr"][order][pos] == c["ptr0"]).all()
sel2 = order[pos]
x_corr = p2["x"][:, sel2]
leaf2 = b2l[p2["block"][sel2]]
cells_corr = np.where(leaf2 >= 0, leaf2 * m.cells_per_block + p2["cell"][sel2], -1).astype(np.int64)  # -1: in a periodic ghost block
# ---- one more MoveParticles with GuidingCenter::Mover_SecondOrder for the guiding-centre species (Lapenta2017 for the ions) ----
r.set_mover_mode(1)
r.move()
p3 = r.particles()
order3 = np.argsort(p3["ptr"])
pos3 = np.searchsorted(p3["ptr"][order3], c["ptr0"])
assert (p3["ptr"][order3][pos3] == c["ptr0"]).all()
sel3 = order3[pos3]
leaf3 = b2l[p3["block"][sel3]]
assert (leaf3 >= 0).all()
mu3, _, _ = r.get_reduced(c["ptr0"])
second = {"x_second": p3["x"][:, sel3], "v_second": p3["v"][:, sel3], "cells_second": (leaf3 * m.cells_per_block + p3["cell"][sel3]).astype(np.int64),
          "mu_second": mu3}
sub = np.arange(0, m.n_corners, 16)
out = {
    "block_cells": np.array(cfg.block_cells[:3]), "ghost_cells": np.array(cfg.ghost_cells[:3]), "n_cells": np.array([32, 16, 8]),
    "origin": np.array([-16.0, -8.0, -4.0]), "charge": np.array(cfg.charge[:2]), "mass": np.array(cfg.mass[:2]),
    "species_weight": np.array(cfg.species_weight[:2]), "dt": np.array(cfg.ecsim_dt_total),
    # PIC::MolecularData::GetElectricCharge / GetMass: the raw species tables the guiding-centre mover reads (the ECSIM kernels use the
    # si2no values above)
    "charge_table": np.array(r.charge_si), "mass_table": np.array(r.mass_si),
    "unit": np.array([cfg.ecsim_B_conv, cfg.ecsim_length_conv, cfg.ecsim_light_speed]),
    "E_half": E, "B_prev": Bp, "B_cur": Bc, "E_cur": c["extra"]["E_cur"], "x": x, "v": v, "w": w, "species": sp, "cells": cells,
    "mu0": c["extra"]["mu0"], "vnormal": c["extra"]["vnormal"],
    "x_after": ref["after"]["x"], "v_after": ref["after"]["v"], "cells_after": ref["after"]["cells"], "mu_after": mu1, "vnormal_after": vn1,
    "init_flag_after": flag1,
    "J0": ref["J0"][0], "J1": ref["J1"][0], "M_corners": sub, "M0_sub": ref["M0"][0][sub], "M1_sub": ref["M1"][0][sub],
    "M0_rowsum": ref["M0"][0].sum(axis=1), "M1_rowsum": ref["M1"][0].sum(axis=1), "energy": np.array([ref["energy0"], ref["energy1"]]),
    "conv": np.array([r.charge_conv, r.mass_conv]), "net_charge": rho_u, "species_moments": mom_u.reshape(m.n_corners, r.n_species, 10),
    "species_moments_spread": np.array(mom_spread), "phi": phi_u, "x_corrected": x_corr, "cells_corrected": cells_corr,
}
out.update(second)
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_gyrokinetic.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes,", x.shape[1], "particles", "species-0:", int((sp == 0).sum()))
