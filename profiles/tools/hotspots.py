#!/usr/bin/env python
"""Per-source-line hot spots of a kernel from an ncu report (SASS page) + nvdisasm line info.
usage: hotspots.py report.ncu-rep libdtof_b200.so <mangled-kernel-substring> [top_n]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, so, kern = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "bvh" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l][0]
cur, off2line = None, {}
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text") and off2line:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, ii, it, isamp = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
base = int(rows[2][ia], 16)
agg, tot = collections.defaultdict(lambda: [0, 0, 0]), [0, 0, 0]
for r in rows[2:]:
    key = off2line.get(int(r[ia], 16) - base) or ("?", 0)
    for j, col in enumerate((ii, it, isamp)):
        agg[key][j] += int(r[col])
        tot[j] += int(r[col])
src = {}
print("kernel %s: warp-inst %.3e  thread-inst %.3e  avg active lanes %.2f" % (kern, tot[0], tot[1], tot[1] / tot[0]))
for (f, ln), (wi, ti, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ""
    print("%5.2f%% inst  lanes %4.1f  stall-samples %5.2f%%  %s:%d  %s" % (100 * wi / tot[0], ti / max(wi, 1), 100 * sm / tot[2], f, ln, text))
