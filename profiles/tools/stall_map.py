#!/usr/bin/env python
"""Where do the stall samples of a kernel fall? Groups the SASS rows of an ncu report (--page source) into address ranges
(traversal loop vs the rest) and prints per-range instruction counts and stall reasons.
usage: stall_map.py report.ncu-rep [bucket_instrs]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 256
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
col = {k: hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples", "stall_no_inst",
                                  "stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_math", "stall_not_selected", "stall_selected")}
base = int(rows[2][col["Address"]], 16)
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[2:]:
    off = (int(r[col["Address"]], 16) - base) // 16
    b = off // bucket
    for k, c in col.items():
        if k != "Address":
            v = int(r[c] or 0)
            agg[b][k] += v
            tot[k] += v
print("bucket(instr range)   %inst lanes | %samples: no_inst long_sb wait short_sb branch math not_sel selected")
for b in sorted(agg):
    a = agg[b]
    if a["Instructions Executed"] == 0:
        continue
    s = max(tot["# Samples"], 1)
    print("%5d-%5d  %6.2f %5.1f | %6.2f: %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f" % (
        b * bucket, (b + 1) * bucket - 1, 100 * a["Instructions Executed"] / tot["Instructions Executed"],
        a["Thread Instructions Executed"] / max(a["Instructions Executed"], 1), 100 * a["# Samples"] / s,
        *(100 * a[k] / s for k in ("stall_no_inst", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving",
                                   "stall_math", "stall_not_selected", "stall_selected"))))
print("total samples", tot["# Samples"], {k: round(100 * v / max(tot["# Samples"], 1), 1) for k, v in tot.items() if k.startswith("stall")})
