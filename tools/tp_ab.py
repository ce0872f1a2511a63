"""A/B of the test-particle mover side bench (bench.bench_test_particle_movers).  AMPS_GPU_LIB selects the library."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

r = bench.bench_test_particle_movers(torch, only=sys.argv[1] if len(sys.argv) > 1 else None)
print(os.environ.get("AMPS_GPU_LIB", "default"), json.dumps({k: (round(v["ms_per_move"], 3), v["sub_steps"], v["errors"]) for k, v in r.items()}))
