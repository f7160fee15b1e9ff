#!/bin/bash
# Final round-2 record on the shipped code: GPU test suite, the bench line, the ncu launch list / DRAM-traffic pass /
# --set full capture of the GEMM engine on the profiled step of the same command, and configurations 1-3.
# Usage (repo root, GPU box): bash tools/gpu_final_r02.sh [tag]
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $OUT/${TAG}_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-extras --inflight 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/summarize_ncu.py launches $OUT/${TAG}_launches.csv $OUT/${TAG}_launch_list.md > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
    -k regex:k_gemm2 --csv --log-file $OUT/${TAG}_gemm_traffic.csv $CMD > $OUT/${TAG}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
python tools/summarize_traffic.py $OUT/${TAG}_gemm_traffic.csv $OUT/${TAG}_gemm_traffic.json 1001 nd 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"k_wide_|k_fwd_step|k_bwd_step|k_apply_inv" -c 40 \
    -f -o $OUT/${TAG}_solve_prof $CMD > $OUT/${TAG}_ncu_solve.log 2>&1
echo "ncu solve-phase rc=$?"
python tools/summarize_ncu.py full $OUT/${TAG}_solve_prof.ncu-rep $OUT/${TAG}_ncu_solve_phase > /dev/null 2>&1
rm -f $OUT/${TAG}_solve_prof.ncu-rep
timeout 900 python tools/bench_configs.py --configs 1,2,3 > $OUT/${TAG}_configs.json 2> $OUT/${TAG}_configs.err; echo "configs rc=$?"
du -sh $OUT; ls -la $OUT | grep ${TAG}
