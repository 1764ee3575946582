#!/usr/bin/env python
"""SASS instruction count per source-line bucket of one kernel: sass_lines.py file.o <kernel substring> [bucket]"""
import sys, re, collections, subprocess, tempfile, os, glob
obj, pat = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
cub = glob.glob(d + "/*.cubin")[0]
out = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
cur = None; fn = None; cnt = collections.Counter()
for l in out.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m: fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if fn and pat in fn and re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+\S', l): cnt[cur] += 1
print("total", sum(cnt.values()))
b = collections.Counter()
for k, n in cnt.items():
    b[("none", 0) if k is None else (k[0], k[1] // bucket * bucket)] += n
for k, n in sorted(b.items(), key=lambda x: -x[1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]: print(k, n)
