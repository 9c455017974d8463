#!/bin/bash
# Staged GPU run: parity tests (each stage under its own timeout so a hung kernel cannot take the
# box down), smoke, diagnostics and a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TMO=900 run t1_parity_nontf32 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "not tf32" --timeout 300
TMO=600 run t2_tf32 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tf32" --timeout 120
TMO=900 run t3_models python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 400
TMO=300 run t4_smoke python -c "import __graft_entry__ as g; g.smoke()"
if [ -n "$DIAG" ]; then TMO=800 TAILN=3 run t6_diag python scripts/diag_ops_in_model.py 12000 fp32; fi
TMO=900 TAILN=3 run t5_bench python bench.py --steps 5 --warmup 3 --voxels ${BENCH_VOXELS:-300000} ${BENCH_ARGS}
if [ -n "$BENCH2" ]; then TMO=900 TAILN=3 run t7_bench2 python bench.py --steps 5 --warmup 3 --voxels $BENCH2 --no-cpu-baseline; fi
