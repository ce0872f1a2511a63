"""One case of bench.bench_test_particle_movers (argv[1]: relativistic_boris_dipole_amr | relativistic_gca_dipole), for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

print(bench.bench_test_particle_movers(torch, only=sys.argv[1] if len(sys.argv) > 1 else "relativistic_boris_dipole_amr"))
