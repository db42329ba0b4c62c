#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel in libagx.so:  scripts/sass_stats.py ILi0ELi3ELi128 [top]"""
import collections, re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else "ILi0ELi3ELi128"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["cuobjdump", "-sass", "airgym_b200/libagx.so"], capture_output=True, text=True).stdout
cur, hist, total = None, collections.Counter(), 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur and "agx_step_kernel" in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            hist[m.group(2)] += 1
            total += 1
print("total", total)
for k, v in hist.most_common(top):
    print(f"{v:6d} {k}")
