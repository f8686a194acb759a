#!/usr/bin/env python
"""Per-reason and per-instruction warp-stall samples of one kernel in an .ncu-rep (source page).
    python scripts/ncu_stalls.py gpurun_out/prof.ncu-rep k_rows_product [ntop]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h0], [r for r in rows[h0 + 1:] if len(r) == len(rows[h0])]
idx = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[idx["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
print("total samples", tot, " instructions", len(data), " warp-inst executed",
      sum(int(r[idx["Instructions Executed"]]) for r in data))
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]:
    print(f"  {s:26s}{v:8d} {100 * v / tot:5.1f}%")
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:ntop]:
    st = sorted(((s, int(r[idx[s]])) for s in stalls if int(r[idx[s]]) > 0), key=lambda kv: -kv[1])[:2]
    print(r[idx["# Samples"]].rjust(7), r[idx["Instructions Executed"]].rjust(10),
          r[idx["Source"]].strip()[:58].ljust(58), " ".join(f"{a[6:]}={b}" for a, b in st))
