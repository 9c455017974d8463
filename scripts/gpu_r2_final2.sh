#!/bin/bash
# Round-2 final evidence pass, second edition (CTA-pair forward kernel, 2-kernel coordinate insert; one GPU): tests, bench lines (default / variants / CPU arm), ResNet14 configs, map and
# layer sweeps, ncu --set full captures of the dominant kernels, launch list of two bench steps.
mkdir -p gpurun_out
P=gpurun_out/r2h
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 > ${P}_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > ${P}_bench.log 2> ${P}_bench.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --detail > ${P}_bench_detail.log 2> ${P}_bench_per_layer.txt
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --scenes 1 > ${P}_bench_1scene.log 2>/dev/null
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --fused-head > ${P}_bench_fused_head.log 2>/dev/null
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --geometry faithful --scene-scale 0.34 > ${P}_bench_2B_034.log 2>/dev/null
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --geometry faithful --scene-scale 0.56 > ${P}_bench_2B_056.log 2>/dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_cpu_default.log 2>/dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --cpu-voxels 1000000 --cpu-forward-only > ${P}_cpu_same_scene_fwd.log 2>/dev/null
timeout 300 python scripts/bench_resnet14.py --batch 16 --steps 30 > ${P}_resnet14.log 2>&1
timeout 300 python scripts/bench_resnet14.py --batch 4 --steps 30 >> ${P}_resnet14.log 2>&1
timeout 300 python scripts/bench_resnet14.py --batch 64 --steps 20 >> ${P}_resnet14.log 2>&1
timeout 300 python scripts/sweep_maps.py 1000000 > ${P}_maps_1m.log 2>&1
for shape in "32 32" "64 64" "96 96" "128 96" "128 128" "256 256"; do timeout 120 python scripts/microbench_conv.py 1000000 $shape --prec bf16 --reps 7; done > ${P}_layers_bf16.log 2>&1
timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec tf32 --reps 7 >> ${P}_layers_bf16.log 2>&1
timeout 300 python scripts/sweep_maps.py 10000000 2>&1 | grep -v "BN\|ReLU" > ${P}_maps_10m.log
# ncu: full captures + launch list (numbers printed under ncu are never bench values)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_umma_kernel|conv_umma_pair_kernel|conv_wgrad_umma_kernel" -s 1 -c 5 -f -o ${P}_prof_conv96 python scripts/microbench_conv.py 1000000 96 96 --reps 1 --prec bf16 > ${P}_ncu_conv96.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"col_reduce_kernel|bn_apply_kernel|bn_bwd_apply_kernel|kernel_map_sym_kernel|insert_kernel|assign_rows_kernel|plenoxel_decode" -s 9 -c 9 -f -o ${P}_prof_rows python scripts/bench_rows.py 1000000 96 > ${P}_ncu_rows.log 2>&1
ncu -i ${P}_prof_conv96.ncu-rep --page raw --csv > ${P}_prof_conv96_raw.csv 2>/dev/null
ncu -i ${P}_prof_rows.ncu-rep --page raw --csv > ${P}_prof_rows_raw.csv 2>/dev/null
ncu -i ${P}_prof_conv96.ncu-rep --page source --csv > ${P}_prof_conv96_src.csv 2>/dev/null
python scripts/ncu_raw_summary.py ${P}_prof_conv96_raw.csv --json ${P}_traffic_conv96.json > ${P}_ncu_full_conv96_bf16.txt 2>&1
python scripts/ncu_raw_summary.py ${P}_prof_rows_raw.csv --json ${P}_traffic_rows.json > ${P}_ncu_full_rows.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt-precision > ${P}_launches_run.log 2>&1
python scripts/ncu_launch_summary.py ${P}_launches.csv > ${P}_launches.txt 2>&1
gzip -f ${P}_launches.csv
rm -f ${P}_prof_conv96.ncu-rep ${P}_prof_rows.ncu-rep   # (the CSV pages are kept; the reports are large)
tail -3 ${P}_tests.log; tail -c 600 ${P}_bench.log; cat ${P}_resnet14.log | tail -6; head -12 ${P}_launches.txt
