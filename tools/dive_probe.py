"""time the div-E correction particle passes (ComputeNetCharge, species corner moments, CorrectParticleLocation) on the bench workload"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from amps_b200 import api

m, cfg, parts, fields = bench.build_box((64, 64, 64), 64)
g = api.Context(cfg, m)
g.fields_upload(*fields)
g.particles_upload(*parts)
g.step()
n = parts[0].shape[1]
phi = 1e-3 * np.sin(np.asarray(m.center_x)[:, 0] * 2 * np.pi / 64.0)
g.SetPhi(phi)
for name, fn in [("net_charge", lambda: g.ComputeNetCharge(1.0)), ("species_moments", lambda: g.ComputeSpeciesMoments(download=False)),
                 ("correct_particle_location", lambda: g.CorrectParticleLocation(1.0, 1.0))]:
    ts = []
    for it in range(4):
        if name == "correct_particle_location":
            g.sort()
            g.ComputeSpeciesMoments(download=False)
        g.synchronize()
        t0 = time.perf_counter()
        r = fn()
        g.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"{name}: {1e3 * min(ts):.3f} ms for {n} particles (host-timed incl. the result copy)", r if name.startswith("correct") else "")
g.close()
