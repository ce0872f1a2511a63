#!/usr/bin/env python
"""Where a kernel's warps spend their time: the SASS of `ncu --page source --csv` cut into contiguous regions of equal execution
count (per launch / per cell / per chunk / per loop trip), each with its share of the warp stall samples, plus the hottest single
instructions.  Usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; tools/stall_regions.py src.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = rows[1]
isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
recs = []
for n, r in enumerate(rows[2:]):
    if r and r[0] == "Kernel Name":
        break  # several launches in one export: the first one is enough
    if len(r) > ist and r[0] != "Address":
        recs.append((n, int(r[iex]), int(r[ist]), r[isrc].strip()))
tot_st, tot_ex = sum(r[2] for r in recs), sum(r[1] for r in recs)
print(rows[0][1][:100])
print(f"executed warp-instructions {tot_ex}, stall samples {tot_st}")
regions, cur = [], None
for n, ex, st, s in recs:
    if cur and abs(ex - cur[2]) <= 0.02 * max(ex, 1) + 1:
        cur[1], cur[3], cur[4], cur[5] = n, cur[3] + st, cur[4] + 1, cur[5] + ex
    else:
        cur = [n, n, ex, st, 1, ex, s]
        regions.append(cur)
print("rows        executions  instrs  exec%  stall%  first instruction")
for a, b, ex, st, cnt, exsum, s in regions:
    if st / tot_st >= min_share:
        print(f"{a:5d}-{b:<5d} {ex:10d} {cnt:6d} {100.0 * exsum / tot_ex:6.2f} {100.0 * st / tot_st:6.2f}  {s[:60]}")
print("hottest instructions:")
for n, ex, st, s in sorted(recs, key=lambda r: -r[2])[:15]:
    print(f"{n:5d} {ex:10d} {100.0 * st / tot_st:6.2f}%  {s[:80]}")
