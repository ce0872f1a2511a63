"""Writes tests/golden/ref_fastwave.npz: the reference's OWN PIC core (oracle/_ref/libref_pic.so = the reference's src/pic compiled
for its fast-wave ECSIM test, see oracle/ref_pic/) run on every 32nd particle of the fast-wave initial condition:
inputs (unique-node fields, particles) and what the reference computed from them -- x', v', (block, cell) after
PIC::Mover::MoveParticles (Lapenta2017) + the periodic exchange, and J / the mass matrix of ECSIM::UpdateJMassMatrix before and
after the move (J everywhere, M on every 16th corner plus the per-corner sum of its 243 entries everywhere).

    python tests/golden/make_ref_fastwave.py        (needs /root/reference at build time of the library)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import ref_ecsim_case as rc  # noqa: E402

c = rc.case(keep_every=32)
ref, m, cfg = c["ref"], c["mesh"], c["cfg"]
x, v, w, sp, cells = c["parts"]
E, Bp, Bc = c["fields"]
sub = np.arange(0, m.n_corners, 16)
out = {
    "block_cells": np.array(cfg.block_cells[:3]), "ghost_cells": np.array(cfg.ghost_cells[:3]), "n_cells": np.array([32, 16, 8]),
    "origin": np.array([-16.0, -8.0, -4.0]), "charge": np.array(cfg.charge[:2]), "mass": np.array(cfg.mass[:2]),
    "species_weight": np.array(cfg.species_weight[:2]), "dt": np.array(cfg.ecsim_dt_total),
    "unit": np.array([cfg.ecsim_B_conv, cfg.ecsim_length_conv, cfg.ecsim_light_speed]),
    "E_half": E, "B_prev": Bp, "B_cur": Bc, "x": x, "v": v, "w": w, "species": sp, "cells": cells,
    "x_after": ref["after"]["x"], "v_after": ref["after"]["v"], "cells_after": ref["after"]["cells"],
    "J0": ref["J0"][0], "J1": ref["J1"][0], "M_corners": sub, "M0_sub": ref["M0"][0][sub], "M1_sub": ref["M1"][0][sub],
    "M0_rowsum": ref["M0"][0].sum(axis=1), "M1_rowsum": ref["M1"][0].sum(axis=1), "energy": np.array([ref["energy0"], ref["energy1"]]),
}
path = os.path.join(ROOT, "tests", "golden", "ref_fastwave.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes,", x.shape[1], "particles")
