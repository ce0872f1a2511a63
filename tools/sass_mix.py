#!/usr/bin/env python
"""Opcode mix / hottest SASS of one kernel from `ncu --page source --csv`. Usage: tools_sass_mix.py src.csv [nparticles]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
npart = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ia, isrc, iex, ith, ist = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
mix = collections.Counter(); stall = collections.Counter(); tot = 0; totst = 0
recs = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break  # several launches in one export: the first one is enough
    if len(r) <= ist or r[0] == "Address": continue
    op = r[isrc].strip().split()
    if not op: continue
    o = op[1] if op[0].startswith('@') else op[0]
    o = o.split('.')[0]
    ex = int(r[iex]); st = int(r[ist])
    mix[o] += ex; stall[o] += st; tot += ex; totst += st
    recs.append((ex, st, r[ia], r[isrc].strip()))
print("static SASS instrs:", len(recs), " executed warp-instrs:", tot, " per particle (thread-instr):", tot * 32 / npart)
for o, c in mix.most_common(25):
    print(f"{o:10s} {c:12d} {100.0*c/tot:6.2f}%  stall {100.0*stall[o]/max(1,totst):6.2f}%  per-particle {c*32/npart:8.1f}")
