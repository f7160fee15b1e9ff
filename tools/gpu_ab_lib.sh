#!/bin/bash
# A/B of two builds of libgmrfb.so (tools/ab/libgmrfb_<name>.so) and of the GEMM engine's switches on the bench step,
# plus the batched-GEMM probe.  Usage: bash tools/gpu_ab_lib.sh <tag> "<lib>[:ENV=val ...]" ...
set -u
TAG=${1:-ab}; shift
OUT=gpurun_out; mkdir -p $OUT
for p in ${PROBES:-}; do
  [ -x tools/probe/$p ] && timeout 300 tools/probe/$p > $OUT/${TAG}_$p.txt 2>&1; echo "$p rc=$?"
done



cp diffeqgmrfs.jl_b200/libgmrfb.so /tmp/libgmrfb_tree.so
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.txt
i=0
for spec in "$@"; do
  i=$((i+1))
  lib=${spec%%:*}; envs=""; [ "$spec" != "$lib" ] && envs=${spec#*:}
  cp tools/ab/libgmrfb_$lib.so diffeqgmrfs.jl_b200/libgmrfb.so
  env $envs GMRFB_PROFILE_DUMP=$OUT/${TAG}_${i}_dump.csv timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --inflight ${INFLIGHT:-1} > $OUT/${TAG}_${i}.json 2> $OUT/${TAG}_${i}.err
  python - <<PY
import json
d=json.load(open("$OUT/${TAG}_${i}.json"))
kp={r["name"]:r for r in d["kernel_profile"]}
print("[$spec] ms_per_step %.2f value %.2f"%(d["ms_per_step"],d["value"]), " ".join("%s %.2fms %.1fTF"%(k.split()[0],kp[k]["ms"],kp[k].get("tflops",0)) for k in kp if "gemm" in k), "ok", d["parity_check"]["ok"])
PY
done
cp /tmp/libgmrfb_tree.so diffeqgmrfs.jl_b200/libgmrfb.so
