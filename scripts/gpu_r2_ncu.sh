#!/bin/bash
# ncu --set full captures of the round-2 kernels (one GPU): conv fwd / dgrad / wgrad 96->96 at 1 M voxels, BN, maps
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_umma_kernel|conv_wgrad_umma_kernel" -s 3 -c 3 -f -o gpurun_out/r2_prof_conv96 python scripts/microbench_conv.py 1000000 96 96 --reps 1 --prec bf16 > gpurun_out/r2_ncu_conv96.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"col_reduce_kernel|bn_apply_kernel|bn_bwd_apply_kernel|kernel_map_sym_kernel|insert_kernel|assign_rows_kernel|plenoxel_decode" -s 9 -c 9 -f -o gpurun_out/r2_prof_rows python scripts/bench_rows.py 1000000 96 > gpurun_out/r2_ncu_rows.log 2>&1
ncu -i gpurun_out/r2_prof_conv96.ncu-rep --page raw --csv > gpurun_out/r2_prof_conv96_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_rows.ncu-rep --page raw --csv > gpurun_out/r2_prof_rows_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_conv96.ncu-rep --page source --csv > gpurun_out/r2_prof_conv96_src.csv 2>/dev/null
python scripts/ncu_raw_summary.py gpurun_out/r2_prof_conv96_raw.csv --json gpurun_out/r2_traffic_conv96.json > gpurun_out/r2_ncu_full_conv96_bf16.txt 2>&1
python scripts/ncu_raw_summary.py gpurun_out/r2_prof_rows_raw.csv --json gpurun_out/r2_traffic_rows.json > gpurun_out/r2_ncu_full_rows.txt 2>&1
tail -12 gpurun_out/r2_ncu_full_conv96_bf16.txt; tail -12 gpurun_out/r2_ncu_full_rows.txt
