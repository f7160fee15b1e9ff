#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch list and one --set full capture of the GEMM kernel.
# Usage (from the repo root on the GPU box): bash tools/gpu_evidence.sh <tag>
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.txt
tail -3 $OUT/${TAG}_pytest_gpu.txt
( time python bench.py --steps 5 --warmup 3 ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
( time python bench.py --impl reference --steps 2 --warmup 1 ${REF_ARGS:-} ) > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --inflight 1"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# DRAM traffic of every GEMM launch of the same step (one pass, three metrics): roofline.traffic of the bench line
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
    -k regex:k_gemm2 --csv --log-file $OUT/${TAG}_gemm_traffic.csv $CMD > $OUT/${TAG}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${NCU_KERNEL:-k_gemm2} -s ${NCU_KSKIP:-193} -c ${NCU_KCOUNT:-15} \
    -f -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -20
