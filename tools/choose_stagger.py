#!/usr/bin/env python
"""Pick the worker stagger of bench.py's two timed legs from A/B runs (gpurun_out/<tag>_ab_*.json): prints a table
(markdown) and writes `export GMRFB_BENCH_STAGGER_MS=..` / `export GMRFB_BENCH_E2E_STAGGER_MS=..` lines to <out.env>.
A stagger is only taken when its mean beats the unstaggered mean by more than 1 %."""
import glob
import json
import sys
from collections import defaultdict

tag, out_env, out_md = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
for f in sorted(glob.glob(f"gpurun_out/{tag}_ab_*.json")):
    try:
        d = json.load(open(f))
    except Exception:  # noqa: BLE001
        continue
    rows.append(dict(file=f, inflight=d["detail"]["problems_per_gpu_in_flight"], steps=d["steps"],
                     dev=d["detail"].get("worker_stagger_ms", 0.0), e2e=d["e2e"].get("worker_stagger_ms", 0.0),
                     value=d["value"], e2e_value=d["e2e"]["value"], single_ms=d["detail"]["single_solve_latency_ms"]))
md = ["# Worker stagger of the timed legs (bench.py --stagger-ms / --e2e-stagger-ms), A/B on one B200", "",
      "| run | in flight | steps | device-leg stagger ms | solves/s (device resident) | e2e stagger ms | solves/s (end to end) |",
      "|---|---|---|---|---|---|---|"]
for r in rows:
    md.append(f"| `{r['file'].split('/')[-1]}` | {r['inflight']} | {r['steps']} | {r['dev']} | {r['value']:.2f} | {r['e2e']} | {r['e2e_value']:.2f} |")
base = [r for r in rows if r["inflight"] == 4]
dv, ev = defaultdict(list), defaultdict(list)
for r in base:
    dv[r["dev"]].append(r["value"])
    ev[r["e2e"]].append(r["e2e_value"])
mean = lambda v: sum(v) / len(v)  # noqa: E731
best_dev = max(dv, key=lambda k: mean(dv[k])) if dv else 0.0
best_e2e = max(ev, key=lambda k: mean(ev[k])) if ev else 0.0
if 0.0 in dv and mean(dv[best_dev]) < 1.01 * mean(dv[0.0]):
    best_dev = 0.0
if 0.0 in ev and mean(ev[best_e2e]) < 1.01 * mean(ev[0.0]):
    best_e2e = 0.0
md += ["", "means (4 in flight): device leg " + ", ".join(f"{k} ms: {mean(v):.2f}" for k, v in sorted(dv.items())) +
       "; end to end " + ", ".join(f"{k} ms: {mean(v):.2f}" for k, v in sorted(ev.items())),
       "", f"chosen: device leg {best_dev} ms, end to end {best_e2e} ms"]
open(out_md, "w").write("\n".join(md) + "\n")
open(out_env, "w").write(f"export GMRFB_BENCH_STAGGER_MS={best_dev}\nexport GMRFB_BENCH_E2E_STAGGER_MS={best_e2e}\n")
print("\n".join(md[-3:]))
