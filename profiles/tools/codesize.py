#!/usr/bin/env python
"""Static SASS size of a kernel broken down by source line (nvdisasm -g line info).
usage: codesize.py lib.so <mangled-kernel-substring> [top_n]"""
import collections, os, re, subprocess, sys, tempfile
so, kern = sys.argv[1:3]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l][0]
cur, cnt, n = None, collections.Counter(), 0
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text") and n:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", l):
        cnt[cur] += 1
        n += 1
print(f"{kern}: {n} instructions = {n * 16} bytes")
byfile = collections.Counter()
for (f, ln), c in cnt.items():
    byfile[f] += c
print(dict(byfile))
src = {}
for (f, ln), c in cnt.most_common(top_n):
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print(f"{c:6d}  {f}:{ln}  {text}")
