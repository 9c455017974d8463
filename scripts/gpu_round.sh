#!/bin/bash
# One GPU call: parity tests, smoke, bench (bf16 + tf32), ncu launch list of a bench run and one
# `ncu --set full` capture of the dominant convolution kernels.  Logs land in gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-r1c}
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-1500; }
if [ -z "$SKIP_TESTS" ]; then
TMO=1200 run ${TAG}_tests python -m pytest tests -m gpu -x -q --timeout 400
TMO=300 run ${TAG}_smoke python -c "import __graft_entry__ as g; g.smoke()"
fi
TMO=600 TAILN=2 run ${TAG}_bench_bf16 python bench.py --steps 10 --warmup 3 --precision bf16 --detail
TMO=600 TAILN=2 run ${TAG}_bench_tf32 python bench.py --steps 10 --warmup 3 --precision tf32 --detail --no-cpu-baseline
if [ -n "$NCU" ]; then
# launch list: skip the 3 warm-up steps' launches roughly, record ~2 steps
TMO=900 TAILN=2 run ${TAG}_ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2200 --csv \
   --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --precision ${NCU_PREC:-bf16} --no-cpu-baseline
TMO=600 TAILN=3 run ${TAG}_ncu_full ncu --set full --clock-control none --import-source on -k regex:'conv_umma_kernel|conv_wgrad_umma_kernel' -s 1 -c 5 \
   -o gpurun_out/prof_${TAG}_conv96 -f python scripts/microbench_conv.py 1000000 96 96 --reps 1 --prec ${NCU_PREC:-bf16}
fi
