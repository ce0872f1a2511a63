#!/usr/bin/env python
"""bench.py -- particle push+deposit updates/s of the ECSIM particle phase on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over the resident plasma: PIC::Mover::MoveParticles (Lapenta2017)
+ counting sort by (block,cell) [the list hand-off] + ECSIM::UpdateJMassMatrix, i.e. amps_gpu_step().
N=1 workload = BASELINE configs[1]: ECSIM uniform periodic box, 64^3 cells, 64 ppc per species (e/p),
single AMR level, synthetic Maxwellian (3.36e7 particles; inputs are >> L2 so no L2 flush is needed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the CPU restatement of the reference path (oracle port, all host threads) on a
bounded sample of the same workload; the real AMPS cannot be built here.
"""
import argparse
import json
import os

# the shared-corner exchange is a few 10-50 MB point-to-point messages per step; NCCL's default of 2 channels per peer
# moves them at ~45 GB/s.  NCCL reads these once per process (torch initialises it first here), see amps_gpu_comm_init.
os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "16")
os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle push+deposit updates/s"
UNIT = "updates/s"
ALG_BYTES_FIXED = 105.0  # 57 B read (x,v,w,species) + 48 B write (x',v') per update
ALG_BYTES_PER_CELL = 4000.0  # zero+flush of J[3],M[243] per corner/cell (3936 B) + E/B tile (64 B)
ALG_FLOP_PER_UPDATE = 1050.0
# share of the per-update algorithmic bytes that each kernel must move at minimum ("kernels")
KERNEL_ALG_BYTES = {
    "move": lambda P: 57.0 + 48.0 + 64.0 / P,  # read state, write x',v'; E/B tiles
    "sort": lambda P: 0.0,  # implementation overhead, not counted by SURVEY 8d
    "deposit": lambda P: 57.0 + 3936.0 / P,  # re-read state; zero + flush J,M
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


DECOMP = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def build_box(n_cells, ppc, block_cells=(8, 8, 8), seed=100, capacity_slack=1.02, rank=0, world=1, particles=True):
    from amps_b200 import api, mesh as meshmod, workload

    m = meshmod.uniform_periodic_box(n_cells, block_cells, (1, 1, 1), rank=rank, n_ranks=world, decomp=DECOMP[world])
    charge, mass, wgt = workload.species_tables(ppc, 1.0)
    n = len(m.real_leaves()) * m.cells_per_block * 2 * ppc
    parts = workload.maxwellian_box(m, ppc, seed=seed) if particles else None
    cfg = api.make_config(block_cells, (1, 1, 1), charge, mass, wgt, 1.0, periodic=True,
                          capacity=int(n * (capacity_slack if world == 1 else 1.10)) + 1024)
    E, B = workload.box_fields(m, E_amp=0.0)
    return m, cfg, parts, (E, B, B.copy())


def upload_in_slabs(ctx, m, ppc, seed, leaves_per_slab=512):
    """The plasma of a large box is generated and handed over in slabs of blocks (512 blocks of 8^3 cells = the particles of a
    64^3-cell box, 2 GB of host memory) so that 2.7e8 particles per rank never sit in host memory at once."""
    from amps_b200 import workload

    leaves = m.real_leaves()
    n = 0
    for k, l0 in enumerate(range(0, len(leaves), leaves_per_slab)):
        parts = workload.maxwellian_box(m, ppc, seed=seed + 1000 * k, leaves=leaves[l0:l0 + leaves_per_slab])
        (ctx.particles_upload if k == 0 else ctx.particles_append)(*parts)
        n += parts[0].shape[1]
    return n


def timed_steps(ctx, torch, dist, world, local, K, W):
    """W warm-up steps, then exactly K steps between barriers; CUDA-event time on the library's stream, max over ranks"""
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(W):
        ctx.step()
    barrier()
    ctx.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        ctx.step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    phases = ctx.profile(False)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    return ms, {p: phases[p][0] / max(1, K) for p in ("move", "sort", "deposit", "exchange")}


def bench_large_box(torch, dist, rank, world, local, cells_per_gpu, ppc, K=10, W=3):
    """BASELINE configs[2] series: cells_per_gpu^3 cells on every GPU (8 GPUs: the 256^3-cell, 64-ppc box the north-star's
    efficiency target is quoted on; 2.7e8 particles per GPU).  At world > 1 rank 0 afterwards times the same per-GPU box alone
    (no exchange) on its own GPU, so that the line carries the efficiency of the sharded run against world x one GPU."""
    from amps_b200 import api

    dec = DECOMP[world]
    n_cells = tuple(cells_per_gpu * dec[d] for d in range(3))
    t0 = time.time()
    m, cfg, _, fields = build_box(n_cells, ppc, seed=300 + rank, rank=rank, world=world, particles=False)
    cfg.device = local
    ctx = api.Context(cfg, m)
    if world > 1:
        ctx.comm_init(dist)
    ctx.fields_upload(*fields)
    n_part = upload_in_slabs(ctx, m, ppc, seed=300 + 7919 * rank)
    gen_s = time.time() - t0
    ms, phases = timed_steps(ctx, torch, dist, world, local, K, W)
    n_after = ctx.particle_count()
    ctx.close()
    if world > 1:
        nn = torch.tensor([n_part, n_after], dtype=torch.float64, device="cuda")
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
        n_total, n_after = float(nn[0].item()), float(nn[1].item())
    else:
        n_total = float(n_part)
    out = {"workload": f"ECSIM uniform periodic box {n_cells[0]}x{n_cells[1]}x{n_cells[2]} cells, {ppc} ppc/species, 8^3-cell blocks, "
                       f"{cells_per_gpu}^3 cells per GPU (BASELINE configs[2] at 8 GPUs)",
           "value": n_total * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "steps": K, "warmup": W, "particles_total": n_total,
           "particles_after": n_after, "phases_ms_per_step": phases, "setup_s": round(gen_s, 1)}
    if world > 1:
        single = None
        if rank == 0:
            m1, cfg1, _, f1 = build_box((cells_per_gpu,) * 3, ppc, seed=300, particles=False)
            cfg1.device = local
            c1 = api.Context(cfg1, m1)
            c1.fields_upload(*f1)
            n1 = upload_in_slabs(c1, m1, ppc, seed=300)
            ms1, ph1 = timed_steps(c1, torch, None, 1, local, K, W)
            c1.close()
            single = {"value": n1 * K / (ms1 * 1e-3), "ms_per_step": ms1 / K, "phases_ms_per_step": ph1}
        dist.barrier()
        if rank == 0:
            out["single_gpu_same_box"] = single
            out["parallel_efficiency"] = out["value"] / (world * single["value"])
    return out


def cpu_port_rate(n_cells, ppc, steps, threads, warmup=1, min_seconds=0.0, max_steps=400):
    """oracle (CPU port of the reference path) on a bounded sample: move + list swap + periodic wrap + deposit.
    At least `steps` steps, and further ones until `min_seconds` of timed CPU work are reached (at most max_steps)."""
    from oracle.oracle_py import Oracle

    m, cfg, parts, fields = build_box(n_cells, ppc)
    try:
        o = Oracle(cfg, m, "fast")
    except OSError:
        o = Oracle(cfg, m, "parity")
    o.set_fields(*fields)
    o.add_particles(*parts)
    n = parts[0].shape[1]
    for _ in range(max(1, warmup)):  # warm-up steps (page faults, list order)
        o.move_fast(0, threads)
        o.deposit(threads, want_arrays=False)
    per_step = []
    while len(per_step) < steps or (sum(per_step) < min_seconds and len(per_step) < max_steps):
        t0 = time.perf_counter()
        o.move_fast(0, threads)
        o.deposit(threads, want_arrays=False)
        per_step.append(time.perf_counter() - t0)
    o.close()
    return n, per_step


def reference_code_rate(steps=6, procs=None):
    """The reference's OWN mover + deposit (PIC::Mover::MoveParticles -> Lapenta2017, ECSIM::UpdateJMassMatrix -> ProcessCell), compiled
    from /root/reference at -O3 into oracle/_ref/libref_pic.so (the library that also pins the oracle) (oracle/ref_pic/build_ref_pic.sh), timed on the box it is built for:
    the reference's fast-wave test (16x8x4-cell blocks, 783 360 particles, one rank).  The reference parallelises with MPI ranks, the
    library is one rank: `procs` (default: every host core) independent instances run the same steps at the same time (they start on a
    common clock after their set-up), which is what a domain-decomposed run costs without its exchanges.  Each runs in a child process
    (the reference's state is global and it prints to stdout).  None when the library is not there."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libref_pic.so")
    if not os.path.exists(lib):
        return None
    procs = procs or (os.cpu_count() or 1)
    t_start = time.time() + 12.0  # set-up of one instance takes 1-3 s
    code = (
        "import sys, os, time, json, ctypes as C, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import oracle.ref_pic.ref_pic as rp\n"
        f"rp.LIB = {lib!r}\n"
        "r = rp.RefPic(); e = np.zeros(1)\n"
        "with rp.quiet():\n"
        "    r.lib.ref_pic_move(); r.lib.ref_pic_update_JM(e.ctypes.data_as(C.c_void_p))\n"
        f"late = time.time() > {t_start!r}\n"
        f"while time.time() < {t_start!r}: time.sleep(0.005)\n"
        "t0 = time.time()\n"
        f"for it in range({steps}):\n"
        "    with rp.quiet():\n"
        "        r.lib.ref_pic_move(); r.lib.ref_pic_update_JM(e.ctypes.data_as(C.c_void_p))\n"
        "t1 = time.time()\n"
        "sys.stderr.write('REFCODE ' + json.dumps({'n': int(r.n_particles), 't0': t0, 't1': t1, 'late': late}) + '\\n')\n"
    )
    try:
        env = dict(os.environ, OMP_NUM_THREADS="1")
        ps = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env) for _ in range(procs)]
        recs = []
        for q in ps:
            _, err = q.communicate(timeout=600)
            recs.append(json.loads([l for l in err.splitlines() if l.startswith("REFCODE ")][-1][8:]))
    except Exception as exc:  # noqa: BLE001
        return {"error": repr(exc)[:200]}
    span = max(r["t1"] for r in recs) - min(r["t0"] for r in recs)
    v = sum(r["n"] for r in recs) * steps / span
    one = recs[0]["n"] * steps / min(r["t1"] - r["t0"] for r in recs)
    return {"value": v, "unit": UNIT, "cores": procs, "kind": "reference", "fastest_instance": one, "late_starts": sum(1 for r in recs if r["late"]),
            "sample": f"the reference's own code (oracle/_ref/libref_pic.so, g++ -O3): {procs} one-rank instances at the same time, each on the "
                      f"reference's fast-wave test box ({recs[0]['n']} particles, 16x8x4-cell blocks), {steps} steps of MoveParticles + UpdateJMassMatrix; "
                      "value = all particle updates / the span from the first start to the last end"}


def best_cpu_baseline(port):
    """cpu_baseline of a bench line: the port on the bench's own box (same config) and, when it is built, the reference's own code on all
    cores; the larger of the two is the baseline (the GPU is compared with the best CPU number this box gives), the other is kept beside it."""
    ref = reference_code_rate()
    if ref is None or "error" in ref or ref["value"] <= port["value"]:
        out = dict(port)
        if ref is not None:
            out["reference_code"] = ref
        return out
    out = dict(ref)
    out["port_same_config"] = port
    return out


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores
    cells = (args.cells,) * 3  # the bench arm's own box (same_config): ~1.2 s per step at 64^3 cells on 16 cores
    W = max(1, args.warmup)
    n, per_step = cpu_port_rate(cells, args.ppc, max(1, args.steps), threads, warmup=W)
    dt = float(np.sum(per_step))
    val = n * len(per_step) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(per_step), "warmup": W,
        "ms_per_step": 1e3 * dt / len(per_step), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"ECSIM uniform periodic box {args.cells}^3 cells, {args.ppc} ppc/species e+p, 8^3-cell blocks, Maxwellian, dt=1",
                   "sample": f"the whole {cells[0]}^3-cell box ({n} particles) per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{cells[0]}^3 cells x {args.ppc} ppc x 2 species = {n} particles, {len(per_step)} steps, OpenMP by blocks/cells"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    best = best_cpu_baseline(line["cpu_baseline"])
    if best["kind"] == "reference":  # the reference's own code on all cores beats the port on this box: it is the arm's number
        line["cpu_baseline"] = best
        line["value"] = best["value"]
        line["e2e"]["value"] = best["value"]
        line["ms_per_step"] = 1e3 * n / best["value"]  # what one step of the configured box costs at this rate
        line["config"]["sample"] = best["sample"]
    else:
        line["cpu_baseline"] = best
    _emit(line)


def bench_test_particle_movers(torch, n_particles=2_000_000, presort=True, only=None):
    """BASELINE configs[0] / [4] (scaled): pushes/s of the test-particle movers on one GPU, extra information next to the
    headline line.  Relativistic Boris: protons traced backward in a dipole on a 3-level AMR mesh with 4^3-cell blocks, sub-cycled
    by the local gyro period (pushes counted as mover calls, sub-cycles not counted).  Relativistic GCA: the MoverTest field
    (dipole + E = -v x B) on 5^3-cell blocks with 2 ghost layers."""
    from amps_b200 import _capi, api, workload as wl

    out = {}
    cases = {
        "relativistic_boris_dipole_amr": dict(mover=_capi.MOVER_RELATIVISTIC_BORIS, kw=dict(amr_levels=2, n_blocks=4), dt=0.05, backward=1,
                                              boundary=_capi.BOUNDARY_USER_FUNCTION),
        "relativistic_gca_dipole": dict(mover=_capi.MOVER_RELATIVISTIC_GCA, kw=dict(n_blocks=8, block_cells=(5, 5, 5), ghost_cells=(2, 2, 2),
                                                                                    rigidity_gv=(0.001, 0.05)), dt=0.05, backward=0,
                                        boundary=_capi.BOUNDARY_DELETE),
    }
    for name, c in cases.items():
        if only is not None and name != only:
            continue
        m, parts = wl.dipole_test_particles(n_particles, **c["kw"])
        bc = c["kw"].get("block_cells", (4, 4, 4))
        gc = c["kw"].get("ghost_cells", (1, 1, 1))
        cfg = api.make_config(bc, gc, (wl.QP,), (wl.MP,), (1.0,), c["dt"], periodic=False, capacity=n_particles + 16, boundary_mode=c["boundary"])
        cfg.time_step_mode = _capi.DT_SPECIES_GLOBAL
        cfg.coupler_interpolation = _capi.CPLR_LINEAR
        cfg.backward_time_integration = c["backward"]
        cfg.speed_of_light = wl.CLIGHT
        cfg.internal_sphere_radius = wl.RE
        cfg.exit_record_capacity = n_particles
        gca = c["mover"] == _capi.MOVER_RELATIVISTIC_GCA
        cfg.carry_magnetic_moment = 1 if gca else 0
        E, B = wl.background_analytic(m.center_x)
        ctx = api.Context(cfg, m)
        ctx.background_upload(E, B)
        if gca:
            ctx.background_upload_gca(wl.gca_var15(m.center_x, 1.0e3))
        best = None
        for rep in range(3):
            ctx.particles_upload(*parts)
            if presort:  # the mover's input comes from the reference's per-cell lists: cell order, not the generator's random order
                ctx.sort()
            if gca:
                ctx.InitiateMagneticMoment(_capi.MOVER_RELATIVISTIC_GCA)
            ctx.profile(True)
            st = ctx.MoveParticles(c["mover"], raise_on_particle_error=False)
            ctx.sort()
            ph = ctx.profile(False)
            ms = ph["move"][0]
            best = ms if best is None else min(best, ms)
        out[name] = {"pushes_per_s": n_particles / (best * 1e-3), "ms_per_move": best, "particles": n_particles,
                     "alg_bytes_per_push": 113.0 if gca else 105.0,
                     "achieved_gbs": (113.0 if gca else 105.0) * n_particles / (best * 1e-3) / 1e9,
                     "input_order": "by cell (sorted after the upload)" if presort else "random",
                     "left_domain": st["n_left_domain"], "errors": st["n_error"], "sub_steps": st["n_sub_steps"],
                     "sub_steps_per_s": (st["n_sub_steps"] / (best * 1e-3)) if st["n_sub_steps"] else None,
                     "mesh": f"{m.c.n_leaves} blocks of {bc[0]}^3 cells, levels {sorted(set(int(v) for v in m.leaf_level()))}"}
        ctx.close()
    return out


def ncu_traffic_for(phase):
    """(bytes, source) for the kernel of a phase from profiles/ncu_traffic.json, or (None, reason) when the kernel's source
    changed since the capture"""
    import hashlib

    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tab = json.load(f)
        e = tab[phase]
        with open(os.path.join(ROOT, e["source_file"]), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        if sha != e["source_sha256"]:
            return None, f"{e['source_file']} changed since {e['capture']}"
        return float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]), e["capture"]
    except Exception as exc:  # noqa: BLE001
        return None, repr(exc)[:120]


def measure_fp64_peak():
    """tools/fp64_peak (built by __graft_entry__.build): DFMA / DMMA peak of this GPU, ~1 s"""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=120)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        return None


def bench_gca_large(torch, dist, rank, world, local, n_per_gpu, K=5, W=2, chunk=20_000_000):
    """BASELINE configs[4] (input/gca_mover.input): the relativistic guiding-centre mover on n_per_gpu protons per GPU in the
    MoverTest dipole + E = -v x B field tabulated on 512 blocks of 5^3 cells with 2 ghost layers.  Test particles do not interact:
    every rank holds the whole (small) field table and its own share of the particles, so the path shards without any
    data-path collective ("weak": n_per_gpu is fixed as the GPUs grow).  One step = MoveParticles + the list hand-off (sort)."""
    from amps_b200 import _capi, api, workload as wl

    kw = dict(n_blocks=8, block_cells=(5, 5, 5), ghost_cells=(2, 2, 2), rigidity_gv=(0.001, 0.05))
    t0 = time.time()
    m, parts = wl.dipole_test_particles(min(chunk, n_per_gpu), seed=11 + 1000 * rank, **kw)
    cfg = api.make_config((5, 5, 5), (2, 2, 2), (wl.QP,), (wl.MP,), (1.0,), 0.05, periodic=False, capacity=n_per_gpu + 16, boundary_mode=_capi.BOUNDARY_DELETE)
    cfg.time_step_mode = _capi.DT_SPECIES_GLOBAL
    cfg.coupler_interpolation = _capi.CPLR_LINEAR
    cfg.speed_of_light = wl.CLIGHT
    cfg.internal_sphere_radius = wl.RE
    cfg.exit_record_capacity = 1024
    cfg.carry_magnetic_moment = 1
    cfg.device = local
    E, B = wl.background_analytic(m.center_x)
    ctx = api.Context(cfg, m)
    ctx.background_upload(E, B)
    ctx.background_upload_gca(wl.gca_var15(m.center_x, 1.0e3))
    ctx.particles_upload(*parts)
    n = parts[0].shape[1]
    k = 1
    while n < n_per_gpu:
        _, parts = wl.dipole_test_particles(min(chunk, n_per_gpu - n), seed=11 + 1000 * rank + k, **kw)
        ctx.particles_append(*parts)
        n += parts[0].shape[1]
        k += 1
    del parts
    ctx.InitiateMagneticMoment(_capi.MOVER_RELATIVISTIC_GCA)
    setup_s = time.time() - t0
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one():
        ctx.MoveParticles(_capi.MOVER_RELATIVISTIC_GCA, stats=False)
        ctx.sort()

    for _ in range(W):
        one()
    barrier()
    n_start = ctx.particle_count()
    ctx.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(K):
        one()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    ph = ctx.profile(False)
    n_end = ctx.particle_count()
    ctx.close()
    tot = float(n_start)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        nn = torch.tensor([tot], dtype=torch.float64, device="cuda")
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
        tot = float(nn.item())
    return {"workload": f"input/gca_mover.input shape: relativistic GCA (first order), protons 1-50 MV in the MoverTest dipole + E = -v x B, 512 blocks of "
                        f"5^3 cells, 2 ghost layers, dt = 0.05, {n_per_gpu} particles per GPU (field table replicated, particles sharded)",
            "value": tot * K / (ms * 1e-3), "unit": "pushes/s", "ms_per_step": ms / K, "steps": K, "warmup": W, "particles_total": tot,
            "particles_left_rank0": int(n_end), "move_ms": ph["move"][0] / K, "sort_ms": ph["sort"][0] / K, "alg_bytes_per_push": 113.0,
            "achieved_gbs_per_gpu": 113.0 * (tot / world) * K / (ms * 1e-3) / 1e9, "setup_s": round(setup_s, 1)}


def bench_amr_box(torch, local, base_blocks=16, ppc_by_level=(64, 16, 8), K=5, W=2, parity_sample=100_000):
    """BASELINE configs[3]: 3-level AMR box, (8 base_blocks)^3 base cells (128^3), one more level inside each of two spheres (r < 24 and
    r < 12 base cells) about the centre, 64 / 16 / 8 particles per cell and species on the levels ("mixed ppc", weights keep the
    density uniform), drifting Maxwellian, open (DELETE) outer boundary, corner-based B (the ECSIM mode that is defined on a refined
    mesh).  One GPU.  Before the timing, one step of a random sample of the particles is compared with the CPU oracle on the same
    mesh (block hand-off across levels included): per-particle results do not depend on the other particles."""
    from amps_b200 import _capi, api, workload as wl

    t0 = time.time()
    m = wl.amr_sphere_box((base_blocks,) * 3, (8, 8, 8), (1, 1, 1), radii=(24.0 * base_blocks / 16.0, 12.0 * base_blocks / 16.0))
    lev = m.leaf_level()
    n_est = sum(int((lev == l).sum()) * m.cells_per_block * 2 * p for l, p in enumerate(ppc_by_level))
    charge, mass, wgt = wl.species_tables(ppc_by_level[0], 1.0)
    cfg = api.make_config((8, 8, 8), (1, 1, 1), charge, mass, wgt, 1.0, periodic=False, capacity=int(n_est * 1.02) + 1024, boundary_mode=_capi.BOUNDARY_DELETE)
    cfg.b_mode = _capi.B_CORNER_BASED
    cfg.device = local
    E, B = wl.box_fields(m, E_amp=0.0, b_on_corners=True)
    ctx = api.Context(cfg, m)
    ctx.fields_upload(E, B, B.copy())
    n, first, sample = 0, True, None
    rng = np.random.default_rng(17)
    for l, ppc in enumerate(ppc_by_level):
        leaves = np.nonzero(lev == l)[0]
        for k, l0 in enumerate(range(0, len(leaves), 512)):
            x, v, w, sp, cells = wl.maxwellian_box(m, ppc, seed=900 + 7919 * l + 131 * k, drift=(0.02, 0.0, 0.0), leaves=leaves[l0:l0 + 512])
            w *= (ppc_by_level[0] / ppc) / 8.0 ** l
            if sample is None or l > 0:  # keep a few particles of every level for the parity check
                take = rng.choice(x.shape[1], size=min(x.shape[1], parity_sample // (2 * len(ppc_by_level))), replace=False)
                part = (x[:, take].copy(), v[:, take].copy(), w[take].copy(), sp[take].copy(), cells[take].copy())
                sample = part if sample is None else tuple(np.concatenate([a, b], axis=-1) for a, b in zip(sample, part))
            (ctx.particles_upload if first else ctx.particles_append)(x, v, w, sp, cells)
            first = False
            n += x.shape[1]
    setup_s = time.time() - t0
    # ---- parity of the mover on the sample: a second, small context on the same mesh against the oracle ----
    parity = None
    try:
        from oracle.oracle_py import Oracle

        cfg2 = api.make_config((8, 8, 8), (1, 1, 1), charge, mass, wgt, 1.0, periodic=False, capacity=sample[0].shape[1] + 16, boundary_mode=_capi.BOUNDARY_DELETE)
        cfg2.b_mode, cfg2.device = _capi.B_CORNER_BASED, local
        g2 = api.Context(cfg2, m)
        g2.fields_upload(E, B, B.copy())
        g2.particles_upload(*sample)
        st = g2.MoveParticles()
        mv = g2.particles_download()
        g2.close()
        o = Oracle(cfg2, m, "parity")
        o.set_fields(E, B, B.copy())
        o.add_particles(*sample)
        rc, st_o, ret, fc = o.move(0, os.cpu_count() or 1)
        pp = o.particles()
        o.close()
        ns = sample[0].shape[1]
        gx, gv, gc = np.empty((3, ns)), np.empty((3, ns)), np.empty(ns, dtype=np.int64)
        gx[:, mv["ptrs"]], gv[:, mv["ptrs"]], gc[mv["ptrs"]] = mv["x"], mv["v"], mv["cells"]
        alive = fc >= 0
        nv = np.sqrt((pp["v"][:, alive] ** 2).sum(axis=0))
        parity = {"sample": int(ns), "cells_equal": bool((gc == fc).all()), "stats_equal": all(st[k] == st_o[k] for k in st_o),
                  "max_rel_x": float((np.abs(gx[:, alive] - pp["x"][:, alive]).max(axis=0) / np.sqrt((pp["x"][:, alive] ** 2).sum(axis=0))).max()),
                  "max_rel_v": float((np.abs(gv[:, alive] - pp["v"][:, alive]).max(axis=0) / nv).max()),
                  "n_cross_block": int(st["n_cross_block"]), "n_left_domain": int(st["n_left_domain"])}
        parity["ok"] = bool(parity["cells_equal"] and parity["stats_equal"] and parity["max_rel_x"] <= 1e-10 and parity["max_rel_v"] <= 1e-10)
    except Exception as exc:  # noqa: BLE001
        parity = {"ok": False, "error": repr(exc)[:200]}
    ms, phases = timed_steps(ctx, torch, None, 1, local, K, W)
    n_after = ctx.particle_count()
    ctx.close()
    return {"workload": f"ECSIM 3-level AMR box: {8 * base_blocks}^3 base cells, +1 level inside r < {24.0 * base_blocks / 16.0:g} and r < {12.0 * base_blocks / 16.0:g} base cells, "
                        f"ppc/species {ppc_by_level} by level, drift 0.02, open boundary, corner-based B; {m.n_leaves} blocks of 8^3 cells "
                        f"({[int((lev == l).sum()) for l in range(len(ppc_by_level))]} per level)",
            "value": n * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "steps": K, "warmup": W, "particles": n, "particles_after": n_after,
            "phases_ms_per_step": phases, "mover_parity_on_sample": parity, "setup_s": round(setup_s, 1)}


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version banner) must not add lines to stdout: everything written to fd 1 while the
    benchmark runs goes to stderr; the one JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cells", type=int, default=64, help="box edge in cells (per GPU for N>1)")
    ap.add_argument("--ppc", type=int, default=64, help="particles per cell per species")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tp", action="store_true", help="skip the test-particle mover side measurements")
    ap.add_argument("--no-large", action="store_true", help="skip the 128^3-cells-per-GPU series (BASELINE configs[2])")
    ap.add_argument("--large-cells", type=int, default=128, help="cells per GPU edge of the large-box series")
    ap.add_argument("--gca-particles", type=int, default=100_000_000, help="particles per GPU of the guiding-centre series (BASELINE configs[4]); 0 = skip")
    ap.add_argument("--no-amr", action="store_true", help="skip the 3-level AMR box (BASELINE configs[3], one GPU)")
    ap.add_argument("--amr-base-blocks", type=int, default=16, help="base blocks per edge of the AMR box (16 = 128^3 base cells)")
    ap.add_argument("--no-mp-parity", action="store_true", help="world > 1: skip the sharded-step parity check before the timing")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; amps_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from amps_b200 import api

    # world > 1: the sharded step is checked against the single-domain CPU oracle on a small box BEFORE anything is timed
    # (tests/mp_parity.py; the oracle is the checker here, never the thing measured)
    mp_parity = None
    if world > 1 and not args.no_mp_parity:
        from tests import mp_parity as mpp

        try:
            ok_mp, rep_mp = mpp.run(dist, rank, world, local)
            if rank == 0:
                keep = ("ok", "cells_equal", "x_bit_equal", "v_bit_equal", "max_rel_J", "max_rel_M", "stats_equal", "sent_total", "recv_total",
                        "books_ok", "fused_step_equal", "fused_max_rel_M", "long_steps", "long_ok", "long_err", "peer_memory", "field_ok", "field_rel_E", "field_rel_B", "field_iterations", "n_total", "n_expected")
                mp_parity = {k: rep_mp.get(k) for k in keep}
        except Exception as exc:  # the verdict must reach the line either way
            mp_parity = {"ok": False, "error": repr(exc)[:300]}

    W = max(3, args.warmup)
    K = max(1, args.steps)
    P = 2 * args.ppc

    t_gen = time.time()
    # weak scaling: every GPU owns an args.cells^3 sub-box of one periodic box (Cartesian block decomposition)
    dec = DECOMP[world]
    n_cells = tuple(args.cells * dec[d] for d in range(3))
    m, cfg, parts, fields = build_box(n_cells, args.ppc, seed=100 + rank, rank=rank, world=world)
    cfg.device = local
    n_part = parts[0].shape[1]
    t_gen = time.time() - t_gen
    ctx = api.Context(cfg, m)
    if world > 1:
        ctx.comm_init(dist)
    exchange_kind = None if world == 1 else ("peer memory (CUDA IPC over NVLink): leavers written into the owner's buffer, counts stay on the device"
                                             if ctx.comm_uses_peer_memory() else "NCCL send/recv with a count round trip through the host")
    ctx.fields_upload(*fields)
    ctx.particles_upload(*parts)
    del parts
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(W):
        ctx.step()
    barrier()
    ctx.profile(True)
    launches0 = ctx.launch_count()

    # ---- timed region: exactly K steps, inputs resident in HBM ----
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    ev0.record(stream)
    for _ in range(K):
        ctx.step()
    ev1.record(stream)
    barrier()
    t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t0, t1)
    phases = ctx.profile(False)
    launches = ctx.launch_count() - launches0
    n_now = ctx.particle_count()

    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        nn = torch.tensor([n_part], dtype=torch.float64, device="cuda")
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
        n_total = float(nn.item())
    else:
        n_total = float(n_part)
    value = n_total * K / (ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers: fields in (H2D), step, J+M out (D2H) ----
    Eh = torch.from_numpy(fields[0]).pin_memory().numpy()
    Bp = torch.from_numpy(fields[1]).pin_memory().numpy()
    Bc = torch.from_numpy(fields[2]).pin_memory().numpy()
    Jh = torch.empty((m.n_corners, 3), dtype=torch.float64).pin_memory().numpy()
    Mh = torch.empty((m.n_corners, 243), dtype=torch.float64).pin_memory().numpy()
    h2d = Eh.nbytes + Bp.nbytes + Bc.nbytes
    d2h = Jh.nbytes + Mh.nbytes
    KE = max(1, args.e2e_steps)
    ctx.fields_upload(Eh, Bp, Bc)
    ctx.step_JM(Jh, Mh)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    te0 = time.perf_counter()
    e0.record(stream)
    for _ in range(KE):
        ctx.fields_upload(Eh, Bp, Bc)
        ctx.step_JM(Jh, Mh)  # == step() + JM_download(), the download pipelined behind the deposit
    e1.record(stream)
    barrier()
    te = time.perf_counter() - te0
    e2e_ms = max(e0.elapsed_time(e1), te * 1e3)  # the D2H read is synchronous: host wall time covers it
    if world > 1:
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = n_total * KE / (e2e_ms * 1e-3)

    # ---- end to end with the field solve on the device (row f1): the host keeps the master copy of E and B like AMPS's node buffers
    # do; per step it sends E^n, B^n (H2D), the device runs ECSIM::TimeStep's field half (J, M never leave HBM) and the particle
    # phase, and E^{n+1}, B^{n+1} come back (D2H).  GMRES tolerance 1e-8 = the reference's own ECSIM test (test/srcFastWave/main.cpp).
    e2e_dev = None
    if True:
        try:
            ctx.field_solver_init(dist if world > 1 else None)
            Ecur = torch.zeros((m.n_corners, 3), dtype=torch.float64).pin_memory().numpy()
            Bcur = torch.from_numpy(fields[2].copy()).pin_memory().numpy()
            outp = {"E": Ecur, "B": Bcur}  # the host's node buffers: read back in place, sent again with the next step
            ctx.fields_upload(Eh, Bp, Bc)
            ctx.step()  # J, M of the resident plasma for the first solve
            its_log = []

            def cycle():
                ctx.E_upload(Ecur)
                ctx.fields_upload(None, None, Bcur)
                its_log.append(ctx.field_step(theta=0.5, tol=1e-8, max_iter=200, restart=30))
                ctx.step()
                ctx.fields_download(E=True, E_half=False, B=True, out=outp)

            for _ in range(2):
                cycle()
            barrier()
            its_log.clear()
            td0 = time.perf_counter()
            e0.record(stream)
            for _ in range(KE):
                cycle()
            e1.record(stream)
            barrier()
            td = time.perf_counter() - td0
            d_ms = max(e0.elapsed_time(e1), td * 1e3)
            if world > 1:
                tt = torch.tensor([d_ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                d_ms = float(tt.item())
            e2e_dev = {"value": n_total * KE / (d_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(Ecur.nbytes + Bcur.nbytes),
                       "d2h_bytes_per_step": int(outp["E"].nbytes + outp["B"].nbytes), "steps": KE, "ms_per_step": d_ms / KE,
                       "gmres_iterations": [i for i, _ in its_log], "gmres_rel_residual": max(r for _, r in its_log), "gmres_tol": 1e-8, "gmres_start": "x0 = 0 (the reference's SetInitialGuess)",
                       "path": "amps_gpu_E_upload + amps_gpu_fields_upload(B^n) -> amps_gpu_field_step (UpdateRhs, GMRES, UpdateB, UpdateE on the "
                               "device; J and M stay in HBM) -> amps_gpu_step -> amps_gpu_fields_download(E^{n+1}, B^{n+1})"}
            # leave the frozen benchmark fields behind for what follows
            ctx.fields_upload(Eh, Bp, Bc)
        except Exception as exc:  # never lose the headline line
            e2e_dev = {"error": repr(exc)[:300]}

    # ---- the same loop with the packed rows (J + the 14 independent neighbour blocks of the symmetric mass matrix) ----
    e2e_packed = None
    if world == 1:
        Ph = torch.empty((m.n_corners, 129), dtype=torch.float64).pin_memory().numpy()
        ctx.fields_upload(Eh, Bp, Bc)
        ctx.step_JM_packed(Ph)
        tp0 = time.perf_counter()
        e0.record(stream)
        for _ in range(KE):
            ctx.fields_upload(Eh, Bp, Bc)
            ctx.step_JM_packed(Ph)
        e1.record(stream)
        torch.cuda.synchronize()
        tp = time.perf_counter() - tp0
        p_ms = max(e0.elapsed_time(e1), tp * 1e3)
        e2e_packed = {"value": n_total * KE / (p_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(Ph.nbytes),
                      "steps": KE, "ms_per_step": p_ms / KE,
                      "note": "amps_gpu_step_JM_packed: M[c][d] == M[c+d][-d] (ProcessCell adds the same block to both corners), so 129 of the "
                              "246 doubles per corner cross PCIe and the host rebuilds the rest while scattering into the corner buffers"}

    ctx.close()
    del ctx

    # ---- BASELINE configs[2] series: 128^3 cells per GPU (256^3 at 8 GPUs), its own short timed block ----
    large = None
    if not args.no_large:
        try:
            large = bench_large_box(torch, dist if world > 1 else None, rank, world, local, args.large_cells, args.ppc)
        except Exception as exc:  # extra block: never lose the headline line
            large = {"error": repr(exc)[:300]}

    # ---- BASELINE configs[4] series: relativistic GCA, 1e8 particles per GPU ----
    gca = None
    if args.gca_particles > 0:
        try:
            gca = bench_gca_large(torch, dist if world > 1 else None, rank, world, local, args.gca_particles)
        except Exception as exc:
            gca = {"error": repr(exc)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA-event time of its launches inside the timed region) ----
    peak, peak_src = load_peaks()
    dom = max(("move", "sort", "deposit"), key=lambda p: phases[p][0])
    dom_ms = phases[dom][0] / max(1, K)
    dom_alg = "deposit" if dom == "sort" else dom
    alg_bytes = KERNEL_ALG_BYTES[dom_alg](P) * n_part
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    step_alg = (ALG_BYTES_FIXED + ALG_BYTES_PER_CELL / P)
    # DRAM bytes of one launch from the `ncu --set full` capture of this very workload: profiles/ncu_traffic.json records, per
    # kernel, dram__bytes_read.sum + dram__bytes_write.sum together with the sha256 of the kernel's source file at capture time;
    # a kernel whose source changed since has no traffic figure (null) until it is profiled again
    traffic, traffic_src = ncu_traffic_for(dom) if (args.cells == 64 and args.ppc == 64 and world == 1) else (None, None)
    fp64 = measure_fp64_peak()
    roofline = {"bound": "hbm", "kernel": {"move": "move_lapenta_fast_kernel", "sort": "perm_kernel", "deposit": "deposit_kernel"}[dom],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "ms_per_launch": dom_ms, "alg_bytes_per_update": KERNEL_ALG_BYTES[dom_alg](P),
                "note": "the kernel is bound by the fp64 pipe (DMMA + DFMA share it: ncu sm__pipe_shared_cycles_active 63 %) and the shared-memory "
                        "wavefronts of the MMA operand staging (l1tex 72 %), not by HBM; inside amps_gpu_step it also writes the sorted "
                        "particle copy (65 B/particle), which the algorithmic bytes do not count"}
    if fp64:
        # the fp64 roofline of the same kernel: executed fp64 work of the deposit per update (DMMA tiles padded 27x12 -> 32x16:
        # 512 FMA, + ~95 DFMA/DMUL of phase 1) against the measured DFMA peak of this GPU (tools/fp64_peak.cu)
        roofline["fp64_peak_tflops"] = fp64.get("dfma_tflops")
        roofline["fp64_dmma_peak_tflops"] = fp64.get("dmma_tflops")
    # the headline end-to-end number: the device-resident cycle where it exists (one rank), else the host-solver path
    e2e_host = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": KE,
                "ms_per_step": e2e_ms / KE,
                "path": "amps_gpu_fields_upload -> amps_gpu_step_JM (J + the full mass matrix to the host's field solver every step)"}
    e2e_main = e2e_dev if (e2e_dev is not None and "value" in e2e_dev) else e2e_host
    step_gbs = step_alg * (n_part * K / (ms * 1e-3)) / 1e9 if world == 1 else step_alg * (value / world) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"ECSIM uniform periodic box {n_cells[0]}x{n_cells[1]}x{n_cells[2]} cells ({args.cells}^3 per GPU, block decomposition "
                               f"{dec[0]}x{dec[1]}x{dec[2]}), {args.ppc} ppc/species e+p, 8^3-cell blocks, single AMR level, Maxwellian "
                               f"v_th,e=0.05, dt=1 (BASELINE configs[1]" + ("" if world == 1 else "; NCCL particle migration + corner J/M exchange each step") + ")",
                   "particles_per_gpu": n_part, "particles_after": n_now, "particle_exchange": exchange_kind, "l2": "inputs (2.2 GB particle SoA) larger than L2, no flush",
                   "step": "amps_gpu_step: move(Lapenta2017; contracted arithmetic + exact pass near cell faces) + permutation sort + "
                           "UpdateJMassMatrix (gathers through the permutation, writes the sorted copy)", "gen_s": round(t_gen, 1)},
        "clocks": clocks,
        "e2e": e2e_main,
        "e2e_host_solver": e2e_host,
        "e2e_packed": e2e_packed,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "phases_ms_per_step": {p: (phases[p][0] / max(1, K)) for p in ("move", "sort", "deposit", "exchange")},
        "roofline_step": {"alg_bytes_per_update": step_alg, "achieved_gbs_per_gpu": step_gbs, "frac_hbm": step_gbs / peak,
                          "fp64_tflops_per_gpu": ALG_FLOP_PER_UPDATE * (value / world) / 1e12,
                          "fp64_frac": (ALG_FLOP_PER_UPDATE * (value / world) / 1e12 / fp64["dfma_tflops"]) if fp64 else None},
    }
    if mp_parity is not None:
        line["mp_parity"] = mp_parity
    if large is not None:
        line["u256_series"] = large
    if gca is not None:
        line["gca_series"] = gca
    if world == 1 and not args.no_amr:
        try:
            line["amr_box"] = bench_amr_box(torch, local, base_blocks=args.amr_base_blocks)
        except Exception as exc:
            line["amr_box"] = {"error": repr(exc)[:300]}
    if world == 1 and not args.no_tp:
        try:
            line["test_particle_movers"] = bench_test_particle_movers(torch)
        except Exception as exc:  # extra information only: never lose the headline line
            line["test_particle_movers"] = {"error": repr(exc)}
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n_cpu, per = cpu_port_rate((args.cells,) * 3, args.ppc, 2, cores, min_seconds=12.0)  # a bounded sample: >= 12 s of CPU work
        v = n_cpu * len(per) / float(np.sum(per))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"the bench box itself: {args.cells}^3 cells x {args.ppc} ppc x 2 species = {n_cpu} particles, {len(per)} steps, "
                                          "oracle -O3 OpenMP"}
        line["cpu_baseline"] = best_cpu_baseline(line["cpu_baseline"])
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
