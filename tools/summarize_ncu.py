#!/usr/bin/env python
"""Summarise ncu artefacts from gpurun_out/ into tracked files under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/X_launches.csv profiles/rNN_launch_list.md
    python tools/summarize_ncu.py full gpurun_out/X_prof.ncu-rep profiles/rNN_ncu_top_kernel  [kernel regex]
"""
import collections
import csv
import json
import re
import subprocess
import sys


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]
    ik, iv, iu, ig = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Grid Size")
    tot = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("gmrfb::", "")
        v = float(r[iv].replace(",", ""))
        v = v / 1000 if r[iu] in ("ns", "nsecond") else v * 1000 if r[iu] in ("ms", "msecond") else v
        tot[name][0] += 1
        tot[name][1] += v
        n += 1
    T = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list of one timed bench step ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` on `python bench.py --steps 1 --warmup 3 "
                "--no-cpu-baseline` (per-launch times are cold-cache and serialised: compare SHARES with the in-run "
                "`kernel_profile` of the bench line, not absolutes).\n\n")
        f.write(f"launches: {n}, summed kernel time: {T / 1000:.2f} ms\n\n| kernel | launches | ms | share | avg us |\n|---|---|---|---|---|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1000:.3f} | {100 * v[1] / T:.1f}% | {v[1] / v[0]:.1f} |\n")
    print(open(dst).read())


def full(src, dst, pattern=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]

    def col(name):
        return h.index(name) if name in h else None

    want = {
        "duration_us": "gpu__time_duration.sum", "dram_read_MB": "dram__bytes_read.sum", "dram_write_MB": "dram__bytes_write.sum",
        "dmma_pipe_pct": "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "registers": "launch__registers_per_thread", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    }
    recs = []
    for r in rows[2:]:
        name = r[col("Kernel Name")]
        if pattern and not re.search(pattern, name):
            continue
        rec = {"kernel": re.sub(r"\(.*", "", name).replace("void ", "").replace("gmrfb::", ""), "grid": r[col("Grid Size")],
               "block": r[col("Block Size")]}
        for k, m in want.items():
            c = col(m)
            if c is not None:
                try:
                    v = float(r[c].replace(",", ""))
                except ValueError:
                    continue
                u = rows[1][c] if len(rows[1]) > c else ""  # ncu picks one unit per column
                if k == "duration_us":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                elif k.endswith("_MB"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                rec[k] = v
        recs.append(rec)
    # the capture of the launch with the longest duration represents the kernel's steady state
    top = max(recs, key=lambda q: q.get("duration_us", 0))
    summary = {"source": src, "kernel": top["kernel"], "captures": len(recs), "largest_launch": top,
               "dram_bytes_per_launch": (top.get("dram_read_MB", 0) + top.get("dram_write_MB", 0)) * 1e6,
               "all_captures": recs}
    json.dump(summary, open(dst + ".json", "w"), indent=1)
    with open(dst + ".md", "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n| kernel | grid | us | dram MB r/w | DMMA pipe % | SM thr % | DRAM thr % | warps act % | regs |\n|---|---|---|---|---|---|---|---|---|\n")
        for q in recs:
            f.write(f"| `{q['kernel']}` | {q['grid']} | {q.get('duration_us', 0):.1f} | {q.get('dram_read_MB', 0):.1f}/{q.get('dram_write_MB', 0):.1f} | "
                    f"{q.get('dmma_pipe_pct', 0):.1f} | {q.get('sm_throughput_pct', 0):.1f} | {q.get('dram_throughput_pct', 0):.1f} | "
                    f"{q.get('warps_active_pct', 0):.1f} | {q.get('registers', 0):.0f} |\n")
    print(open(dst + ".md").read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
