#!/bin/bash
# Round-2 evidence in one gpurun call: default bench, in-flight variant, ncu launch list / DRAM-traffic pass / --set full
# capture of the GEMM engine on the profiled step of the SHIPPED default configuration, and a --set full capture of the
# panel-solve kernels.  The .ncu-rep files are summarised ON THE BOX (gpurun copies back at most 64 MiB) and removed.
# Usage (repo root, GPU box): bash tools/gpu_evidence_r02.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
( time python bench.py --steps 5 --warmup 3 ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-extras --inflight 6 > $OUT/${TAG}_bench_inflight6.json 2> /dev/null
echo "inflight 6 rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-extras --inflight 1"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/summarize_ncu.py launches $OUT/${TAG}_launches.csv $OUT/${TAG}_launch_list.md > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
    -k regex:k_gemm2 --csv --log-file $OUT/${TAG}_gemm_traffic.csv $CMD > $OUT/${TAG}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
python tools/summarize_traffic.py $OUT/${TAG}_gemm_traffic.csv $OUT/${TAG}_gemm_traffic.json 1001 nd 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm2 -s ${NCU_KSKIP:-170} -c ${NCU_KCOUNT:-14} \
    -f -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
python tools/summarize_ncu.py full $OUT/${TAG}_prof.ncu-rep $OUT/${TAG}_ncu_top_kernel > /dev/null 2>&1
ls -la $OUT/${TAG}_prof.ncu-rep
[ $(stat -c %s $OUT/${TAG}_prof.ncu-rep) -gt 30000000 ] && rm -f $OUT/${TAG}_prof.ncu-rep
PCMD="python tools/bench_panel.py --nx 601 --nrhs 50 --reps 2"
$PCMD > $OUT/${TAG}_panel_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_mr_|k_gemm2|k_apply_inv" -s 430 -c 24 \
    -f -o $OUT/${TAG}_panel_prof $PCMD > $OUT/${TAG}_ncu_panel.log 2>&1
echo "ncu panel rc=$?"
python tools/summarize_ncu.py full $OUT/${TAG}_panel_prof.ncu-rep $OUT/${TAG}_ncu_panel > /dev/null 2>&1
rm -f $OUT/${TAG}_panel_prof.ncu-rep
du -sh $OUT; ls -la $OUT | tail -16
