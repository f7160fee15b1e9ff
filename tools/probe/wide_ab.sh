#!/bin/bash
# A/B of the wide-supernode threshold on the bench (single GPU): value, e2e, latency and the solve-phase kernels
for WM in 128 65; do
  GMRFB_WIDE_MIN=$WM GMRFB_WIDE_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --skip-extras --no-cpu-baseline > gpurun_out/wide_ab_$WM.json 2> gpurun_out/wide_ab_$WM.err
  grep -m1 "wide supernodes" gpurun_out/wide_ab_$WM.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/wide_ab_$WM.json").read().strip().splitlines()[-1])
print("WIDE_MIN=$WM value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "latency", round(d["detail"]["single_solve_latency_ms"],2), "ok", d["parity_check"]["ok"])
for r in d["kernel_profile"]:
    if any(k in r["name"] for k in ("wide","fwd_step","bwd_step","gemm<NN>","potrf")): print("   ", r["name"], r["launches"], r["ms"], r.get("gbs"))
PY
done
