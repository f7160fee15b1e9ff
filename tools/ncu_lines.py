#!/usr/bin/env python
"""Per-source-line warp-stall samples of the kernels in an .ncu-rep (needs -lineinfo and `--import-source on`):

    python tools/ncu_lines.py <file.ncu-rep> [kernel regex] [top N lines]

For every captured launch matching the regex: duration, and the source lines with the most stall samples (share of the
kernel's samples).  Text output small enough to travel back from the GPU box."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "."
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
names = [(r[h.index("ID")], r[h.index("Kernel Name")], r[h.index("gpu__time_duration.sum")], r[h.index("Grid Size")]) for r in rows[2:]]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
blocks, cur = [], None
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "Function Name":
        cur = {"name": r[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r[0] not in ("File Path",):
        cur["rows"].append(r)
for k, b in enumerate(blocks):
    if not re.search(pat, b["name"]):
        continue
    hd = b["hdr"]
    i_line, i_src, i_samp = 0, 1, hd.index("# Samples")
    agg, text = defaultdict(int), {}
    for r in b["rows"]:
        try:
            s = int(r[i_samp])
        except (ValueError, IndexError):
            continue
        if r[i_line] not in ("", "-"):
            last = r[i_line]
            text[last] = r[i_src]
        agg[last] += s
    tot = sum(agg.values()) or 1
    meta = names[k] if k < len(names) else ("?", b["name"], "?", "?")
    print(f"== launch {meta[0]}: {re.sub(r'[(].*', '', b['name'])} grid {meta[3]} duration {meta[2]} samples {tot}")
    for line, s in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
        print(f"   {100 * s / tot:5.1f}%  L{line:>5}: {text.get(line, '')[:110].strip()}")
