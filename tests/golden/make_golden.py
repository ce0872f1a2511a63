"""Generates tests/golden/ecsim_box_16x8x8.npz : inputs and oracle outputs of one ECSIM particle phase.

The reference's own golden outputs for this path are not vendored, and the reference cannot be built
here, so these vectors are produced by the parity build of the oracle (oracle/_build/liboracle_parity.so,
-O2 -ffp-contract=off) after it was cross-checked against tests/numpy_ref.py.  They pin the oracle against
accidental change and give the GPU tests a fixture that does not need the oracle at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import parity_util as pu  # noqa: E402

CASE = dict(n_cells=(16, 8, 8), ppc=3, seed=2024, vscale=4.0)

if __name__ == "__main__":
    m, cfg, parts, fields = pu.make_case(**CASE)
    ora = pu.run_oracle(m, cfg, parts, fields)
    out = os.path.join(ROOT, "tests", "golden", "ecsim_box_16x8x8.npz")
    np.savez_compressed(
        out, x=parts[0], v=parts[1], w=parts[2], species=parts[3], cells=parts[4], E_half=fields[0], B_prev=fields[1], B_cur=fields[2],
        x_out=ora["particles"]["x"], v_out=ora["particles"]["v"], final_cell=ora["final_cell"],
        stats=np.array([ora["stats"][k] for k in sorted(ora["stats"])]), stats_keys=np.array(sorted(ora["stats"])),
        J=ora["J"], M=ora["M"].astype(np.float64), energy=ora["energy"], cfl=np.array(ora["cfl"]))
    print("wrote", out, os.path.getsize(out), "bytes;", parts[0].shape[1], "particles")
