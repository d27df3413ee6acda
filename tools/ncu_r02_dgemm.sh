#!/bin/bash
# --set full capture of the 13 large-tile GEMM launches of the second C2 evaluation (B200_PROFILING.md recipe)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/one_eval.py 8192 1 > gpurun_out/r02_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dgemm_kernel<\(int\)128, \(int\)64' -s 13 -c 13 \
    -f -o gpurun_out/r02_dgemm_full python tools/one_eval.py 8192 1 > gpurun_out/r02_ncu_dgemm.log 2>&1
echo "dgemm full rc $?"; tail -3 gpurun_out/r02_ncu_dgemm.log
