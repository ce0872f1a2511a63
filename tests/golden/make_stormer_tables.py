"""Golden vectors of the reference's known-answer tests for the test-particle path, read from the reference tree in this container and
stored as one small JSON fixture (the GPU box has no /root/reference):
  srcEarth/test/C1/reference_C1_stormer.csv      vertical Stoermer cutoff per (altitude, latitude)
  srcEarth/test/C4/reference_C4_invariants.csv   outer-box exit expected / rigidity drift limit per (point, factor)
  srcEarth/test/C2/reference_C2_stormer_symmetry.csv   the cutoff does not depend on the longitude (12 longitudes x 5 latitudes)
  srcEarth/test/C12/reference_C12_stormer_movers.csv   cutoff per mover on two shells (the BORIS rows: the full-orbit mover of row a7)
Usage: python tests/golden/make_stormer_tables.py [/root/reference]"""
import csv
import json
import os
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = {"source": "SWMFsoftware/AMPS srcEarth/test/{C1/reference_C1_stormer,C4/reference_C4_invariants,C2/reference_C2_stormer_symmetry,"
                 "C12/reference_C12_stormer_movers}.csv"}
with open(os.path.join(ref, "srcEarth/test/C1/reference_C1_stormer.csv")) as f:
    out["C1"] = [{"alt_km": float(r["alt_km"]), "lat_deg": float(r["lat_deg"]), "Rc_stormer_GV": float(r["Rc_stormer_GV"])} for r in csv.DictReader(f)]
with open(os.path.join(ref, "srcEarth/test/C4/reference_C4_invariants.csv")) as f:
    out["C4"] = [{"alt_km": float(r["alt_km"]), "lat_deg": float(r["lat_deg"]), "factor": float(r["factor"]), "R_GV": float(r["R_GV"]),
                  "Rc_stormer_GV": float(r["Rc_stormer_GV"]), "expected_allowed": int(r["expected_allowed"]), "rel_dR_limit": float(r["rel_dR_limit"])}
                 for r in csv.DictReader(f)]
with open(os.path.join(ref, "srcEarth/test/C2/reference_C2_stormer_symmetry.csv")) as f:
    out["C2"] = [{"alt_km": float(r["alt_km"]), "lon_deg": float(r["lon_deg"]), "lat_deg": float(r["lat_deg"]), "Rc_stormer_GV": float(r["Rc_stormer_GV"])}
                 for r in csv.DictReader(f)]
with open(os.path.join(ref, "srcEarth/test/C12/reference_C12_stormer_movers.csv")) as f:
    out["C12"] = [{"mover": r["mover"], "alt_km": float(r["alt_km"]), "lat_deg": float(r["lat_deg"]), "Rc_stormer_GV": float(r["Rc_stormer_GV"]),
                   "rel_tol": float(r["rel_tol"])} for r in csv.DictReader(f) if r["mover"] == "BORIS"]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stormer_tables.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["C1"]), "+", len(out["C4"]), "+", len(out["C2"]), "+", len(out["C12"]), "rows")
