#!/usr/bin/env python
"""Static record of the shipped library: per kernel the `cuobjdump -res-usage` line (registers, stack, shared) and a SASS
opcode histogram (cuobjdump -sass), with the mnemonics that matter for this path called out: 128-bit loads (LDG.E.128 /
LDS.128), local-memory traffic (LDL / STL = stack + spills), MUFU, FMNMX3, atomics / reductions, and the Blackwell
data-movement / tensor instructions this path does not use (UTMALDG / UTMASTG / UTC*MMA: none expected, no contraction).
usage: sass_report.py lib.so [kernel-substring ...]   (default: the headline kernels)"""
import collections, re, subprocess, sys
so = sys.argv[1]
want = sys.argv[2:] or ["render_kernelILi1ELb0ELb0ELi0ELb0", "render_kernelILi0ELb0ELb0ELi0ELb0", "wf_trace_kernelILi0ELb0",
                        "wf_trace_kernelILi0ELb1", "wf_trace_kernelILi1ELb0", "wf_shade_kernelILb1ELb0", "wf_generate_kernelILi0",
                        "wf_splat_kernel", "develop_kernel", "peer_reduce_kernel", "rays_kernelILi0"]
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout.split("\n")
usage = {}
for i, l in enumerate(res):
    m = re.match(r"\s*Function (\S+):", l)
    if m:
        usage[m.group(1)] = res[i + 1].strip()
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
cur, hist = None, {}
arch = [l.strip() for l in sass if l.startswith("arch =")]
for l in sass:
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur:
        hist[cur][m.group(1)] += 1
print(f"library {so}: {sorted(set(arch))}")
watch = ["LDG.E.128", "LDS.128", "LDL", "STL", "MUFU", "FMNMX3", "ATOM", "RED", "SHFL", "VOTE", "UTMALDG", "UTMASTG", "UTCMMA", "UTCHMMA", "HMMA"]
for k in want:
    for fn in hist:
        if k in fn:
            h = hist[fn]
            n = sum(h.values())
            print(f"\n== {fn}\n   {usage.get(fn, '?')}\n   {n} SASS instructions = {n * 16} bytes")
            base = collections.Counter()
            for op, c in h.items():
                base[op.split(".")[0]] += c
            print("   by opcode: " + ", ".join(f"{op} {c}" for op, c in base.most_common(24)))
            called = {w: sum(c for op, c in h.items() if op.startswith(w)) for w in watch}
            print("   watched:   " + ", ".join(f"{w} {c}" for w, c in called.items()))
