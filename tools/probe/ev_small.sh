CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-extras --inflight 1"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_front_factor_small" -s 2 -c 1 -f -o gpurun_out/r2_ffs $CMD > gpurun_out/r2_ffs.log 2>&1
python tools/ncu_lines.py gpurun_out/r2_ffs.ncu-rep k_front 40 > gpurun_out/r2_ffs_lines.txt 2>&1
ncu -i gpurun_out/r2_ffs.ncu-rep --page details 2>/dev/null | grep -E "Duration|Registers|Theoretical Occupancy|Achieved Occupancy|Shared Memory Config|Dynamic Shared|Stall|stall|Issue Slots|No Eligible|Eligible Warps|L1/TEX Hit|Mem Busy|Max Bandwidth|Block Limit" > gpurun_out/r2_ffs_details.txt
rm -f gpurun_out/r2_ffs.ncu-rep
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_front_selinv_small" -s 12 -c 1 -f -o gpurun_out/r2_fss $CMD > gpurun_out/r2_fss.log 2>&1
python tools/ncu_lines.py gpurun_out/r2_fss.ncu-rep k_front 40 > gpurun_out/r2_fss_lines.txt 2>&1
ncu -i gpurun_out/r2_fss.ncu-rep --page details 2>/dev/null | grep -E "Duration|Registers|Theoretical Occupancy|Achieved Occupancy|Dynamic Shared|No Eligible|Eligible Warps|Block Limit" > gpurun_out/r2_fss_details.txt
rm -f gpurun_out/r2_fss.ncu-rep
head -50 gpurun_out/r2_ffs_lines.txt
