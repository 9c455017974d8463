#!/bin/bash
# quick GPU iteration: conv parity tests + layer microbenchmarks
# env: TESTK (pytest -k), CFGS ("vox cin cout;vox cin cout"), PRECS, ONLY
mkdir -p gpurun_out
if [ "${TESTK:-x}" != "none" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "${TESTK:-tf32 or bf16}" 2>&1 | tail -4
fi
IFS=';' read -ra LIST <<< "${CFGS:-1000000 96 96;1000000 32 32;200000 128 128;8000 256 256}"
for cfg in "${LIST[@]}"; do
  for prec in ${PRECS:-bf16 tf32}; do
    timeout 120 python scripts/microbench_conv.py $cfg --prec $prec --only ${ONLY:-fwd,dgrad,wgrad} 2>&1 | grep -v "^CUDA kernel\|^For debugging\|^Compile with\|^$" | tail -5
  done
done
