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
launches, cur, fpath = [], None, None
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
        continue
    if r[0] == "Function Name":
        # one section per (launch, source file): a file seen before in the current group starts the next launch
        if cur is None or cur["name"] != r[1] or fpath in cur["files"]:
            cur = {"name": r[1], "files": set(), "rows": []}
            launches.append(cur)
        cur["files"].add(fpath)
        cur["file"], cur["hdr"] = fpath.split("/")[-1], None
    elif r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur.get("hdr") is not None:
        cur["rows"].append((cur["file"], cur["hdr"], r))
for k, b in enumerate(launches):
    if not re.search(pat, b["name"]):
        continue
    agg, text, last = defaultdict(int), {}, None
    for fname, hd, r in b["rows"]:
        try:
            s = int(r[hd.index("# Samples")])
        except (ValueError, IndexError):
            continue
        if r[0] not in ("", "-"):
            last = (fname, r[0])
            text[last] = r[1]
        if last is not None:
            agg[last] += s
    tot = sum(agg.values()) or 1
    meta = names[k] if k < len(names) else ("?", b["name"], "?", "?")
    print(f"== launch {meta[0]}: {re.sub(r'[(].*', '', b['name'])} grid {meta[3]} duration {meta[2]} samples {tot}")
    for (fname, line), s in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
        print(f"   {100 * s / tot:5.1f}%  {fname}:{line:>5}: {text.get((fname, line), '')[:100].strip()}")
