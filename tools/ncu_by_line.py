#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

    python tools/ncu_by_line.py <report.ncu-rep> <kernel regex> <object.o> <mangled-name substring> [top]

ncu's CSV source page is SASS-only; nvdisasm --print-line-info on the same cubin gives the SASS->line
map (objects are compiled with -lineinfo).  The two listings are matched by instruction offset."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kregex, obj, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
seq, cur, infn = [], None, False
for ln in dis:
    m = re.match(r"\.text\.(\S+):", ln)
    if m:
        infn = mangled in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq.append((int(m.group(1), 16), m.group(2), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):     # next launch's section
        break
    if len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][0], 16)
byoff = {int(r[0], 16) - base: r for r in data}
agg, samp, stall = collections.Counter(), collections.Counter(), collections.Counter()
for off, ins, c in seq:
    r = byoff.get(off)
    if r is None:
        continue
    key = c or ("?", 0)
    agg[key] += int(r[ix["Instructions Executed"]] or 0)
    samp[key] += int(r[ix["# Samples"]] or 0)
tot, stot = sum(agg.values()) or 1, sum(samp.values()) or 1
print(f"kernel {kregex}: {tot} warp instructions, {len(seq)} SASS instructions matched {len(byoff)}")
srcdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spectrogram_enhancement_b200", "csrc")
cache = {}
for (f, l), c in sorted(agg.items(), key=lambda kv: -max(kv[1] / tot, samp[kv[0]] / stot))[:top]:
    if f not in cache:
        p = os.path.join(srcdir, f)
        cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
    txt = cache[f][l - 1].strip()[:88] if 0 < l <= len(cache[f]) else ""
    print(f"instr {100 * c / tot:5.1f}%  stall-samples {100 * samp[(f, l)] / stot:5.1f}%  {f}:{l}  {txt}")
