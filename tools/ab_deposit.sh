#!/bin/bash
# A/B of library variants under build_variants/ (AMPS_GPU_LIB): the event-timed phases of bench.py, two passes each
for pass in 1 2; do
for L in build_variants/libamps_gpu_*.so; do
  echo -n "$L : "
  AMPS_GPU_LIB=$L python bench.py --steps 10 --warmup 3 --no-cpu --no-tp --no-large --no-amr --gca-particles 0 --e2e-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['phases_ms_per_step'].items()})"
done
done
