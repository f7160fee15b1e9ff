#!/bin/bash
# Quick GPU check: parity tests, short bench with per-launch dump, block-tridiagonal timing.
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.txt
rm -f $OUT/${TAG}_dump.csv
GMRFB_PROFILE_DUMP=$OUT/${TAG}_dump.csv timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --inflight 1 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print("ms_per_step",d["ms_per_step"],"e2e",d["e2e"]["ms_per_step"],"launches",d["gpu_launches"],"roofline",d["roofline"]["kernel"],round(d["roofline"]["frac"],3))
for r in d["kernel_profile"]: print(r)
PY
timeout 300 python tools/bench_btd.py --b 1024 --N 16 > $OUT/${TAG}_btd_1024.txt 2>&1; cat $OUT/${TAG}_btd_1024.txt
timeout 300 python tools/bench_btd.py --b 4096 --N 3 > $OUT/${TAG}_btd_4096.txt 2>&1; cat $OUT/${TAG}_btd_4096.txt
