CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-extras --inflight 1"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_potrf64" -s 0 -c 3 -f -o gpurun_out/r2_potrfb $CMD > gpurun_out/r2_potrfb.log 2>&1
python tools/ncu_lines.py gpurun_out/r2_potrfb.ncu-rep k_potrf64 45 > gpurun_out/r2_potrfb_lines.txt 2>&1
rm -f gpurun_out/r2_potrfb.ncu-rep
head -60 gpurun_out/r2_potrfb_lines.txt
