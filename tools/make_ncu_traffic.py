#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` report: per bench phase the DRAM bytes of one launch of its kernel and the
sha256 of the kernel's source file at capture time (bench.py quotes the traffic only while the source is unchanged).
Usage: tools/make_ncu_traffic.py rep.ncu-rep label"""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, label = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
KERNELS = {"move": ("move_lapenta_fast_kernel", "amps_b200/csrc/mover_fast.cu"), "deposit": ("deposit_kernel", "amps_b200/csrc/deposit.cu"),
           "sort": ("perm_kernel", "amps_b200/csrc/sort.cu"), "field_operator": ("ecsim_operator_kernel", "amps_b200/csrc/field_solver.cu")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tab = {}
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
if os.path.exists(path):
    tab = json.load(open(path))
for phase, (kname, src) in KERNELS.items():
    for r in rows[2:]:
        if kname in r[ix["Kernel Name"]]:
            rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) * scale[units[ix["dram__bytes_read.sum"]]]
            wr = float(r[ix["dram__bytes_write.sum"]].replace(",", "")) * scale[units[ix["dram__bytes_write.sum"]]]
            sha = hashlib.sha256(open(os.path.join(ROOT, src), "rb").read()).hexdigest()
            tab[phase] = {"kernel": kname, "dram_bytes_read": rd, "dram_bytes_write": wr, "source_file": src, "source_sha256": sha, "capture": label,
                          "gpu_time_us": float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[ix["gpu__time_duration.sum"]]]}
            break
json.dump(tab, open(path, "w"), indent=1)
print(json.dumps(tab, indent=1))
