"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tests/mp_parity.py under torchrun."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world,decomp", [(2, "cart"), (4, "cart"), (8, "cart"), (2, "sfc")])
def test_sharded_step_matches_single_domain_oracle(world, decomp):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + world + (20 if decomp == "sfc" else 0)), os.path.join(ROOT, "tests", "mp_parity.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, MP_PARITY_DECOMP=decomp), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("MP_PARITY ")]
    assert r.returncode == 0 and line, r.stdout[-3000:]
    out = json.loads(line[-1][len("MP_PARITY "):])
    print(out)
    assert out["ok"], out
