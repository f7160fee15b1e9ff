"""Summarise `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` of the GEMM launches
of one bench step into profiles/<name>.json: per kernel instantiation the launches, DRAM bytes and time; bench.py
reads `bytes_per_launch` of the dominant kernel as `roofline.traffic`.
Usage: python tools/summarize_traffic.py gpurun_out/<tag>_gemm_traffic.csv profiles/r02_gemm_traffic.json [nx ordering nd_cover]\n(the configuration of the capture is recorded; bench.py reports the traffic only for a run with the same configuration and launch count)"""
import csv, json, re, sys


def main(src, dst, nx=1001, ordering="nd", nd_cover="1"):
    lines = open(src, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    per = {}
    for r in csv.DictReader(lines[start:]):
        name, metric = r["Kernel Name"], r["Metric Name"]
        val, unit = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        m = re.search(r"k_gemm2<\(?bool\)?(\d), \(?bool\)?(\d)", name) or re.search(r"k_gemm2<(\d), (\d)", name)
        key = "k_gemm<%s%s> (DMMA)" % ("T" if m and m.group(1) == "1" else "N", "N" if m and m.group(2) == "1" else "T") if m else name
        e = per.setdefault(key, {"launches": set(), "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_ms": 0.0})
        e["launches"].add(r["ID"])
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
        if metric == "dram__bytes_read.sum":
            e["dram_read_bytes"] += val * scale
        elif metric == "dram__bytes_write.sum":
            e["dram_write_bytes"] += val * scale
        elif metric == "gpu__time_duration.sum":
            e["time_ms"] += val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    out = {"source": src, "config": {"nx": int(nx), "ordering": ordering, "nd_cover": str(nd_cover),
                                    "command": "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-extras --inflight 1 "
                                               "(profiled replay of one posterior solve, cudaProfilerStart/Stop window)"},
           "kernels": {}}
    for k, e in per.items():
        n = len(e["launches"])
        tot = e["dram_read_bytes"] + e["dram_write_bytes"]
        out["kernels"][k] = {"launches": n, "dram_read_bytes": e["dram_read_bytes"], "dram_write_bytes": e["dram_write_bytes"],
                             "bytes_per_launch": tot / max(n, 1), "time_ms_under_ncu": e["time_ms"]}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
