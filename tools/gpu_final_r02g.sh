#!/bin/bash
# Final round-2 record, second edition: (1) A/B of the worker stagger of the two timed legs at the driver's step counts,
# the winner exported for everything that follows (and then made the default in bench.py); (2) GPU test suite, the bench
# line, the ncu launch list / DRAM-traffic pass of the profiled step of the same command, configurations 1-3 and the P2
# Darcy configuration.
set -u
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
AB="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-extras"
i=0
for CFG in "0 0 4" "25 25 4" "31 31 4" "16 16 4" "25 25 6" "25 25 4" "0 0 4"; do
  set -- $CFG
  i=$((i+1))
  timeout 200 $AB --stagger-ms $1 --e2e-stagger-ms $2 --inflight $3 > $OUT/${TAG}_ab_$i.json 2> $OUT/${TAG}_ab_$i.err; echo "ab $i ($CFG) rc=$?"
done
python tools/choose_stagger.py $TAG $OUT/${TAG}_stagger.env $OUT/${TAG}_stagger_ab.md
source $OUT/${TAG}_stagger.env
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $OUT/${TAG}_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-extras --inflight 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/summarize_ncu.py launches $OUT/${TAG}_launches.csv $OUT/${TAG}_launch_list.md > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
    -k regex:k_gemm2 --csv --log-file $OUT/${TAG}_gemm_traffic.csv $CMD > $OUT/${TAG}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
python tools/summarize_traffic.py $OUT/${TAG}_gemm_traffic.csv $OUT/${TAG}_gemm_traffic.json 1001 nd 1 > /dev/null 2>&1
timeout 600 python tools/bench_configs.py --configs 1,2,3 > $OUT/${TAG}_configs.json 2> $OUT/${TAG}_configs.err; echo "configs rc=$?"
timeout 300 python tools/bench_fem2d.py --nel 300 --out $OUT/${TAG}_fem2d_darcy_p2.json > $OUT/${TAG}_fem2d.log 2>&1; echo "fem2d rc=$?"
du -sh $OUT; ls -la $OUT | grep ${TAG}
