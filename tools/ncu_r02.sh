#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe): plain run first, then the launch list of one C2 evaluation,
# then --set full captures of the 13 large-tile GEMM launches and of two leaves of the second evaluation.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/one_eval.py 8192 1 > gpurun_out/r02_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_c2_lml_grad.csv \
    python tools/one_eval.py 8192 1 > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc $?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dgemm_kernel<128, 64, ' -s 13 -c 13 \
    -f -o gpurun_out/r02_dgemm_full python tools/one_eval.py 8192 1 > gpurun_out/r02_ncu_dgemm.log 2>&1
echo "dgemm full rc $?"
ncu --set full --clock-control none --import-source on -k regex:leaf_potrf_inv -s 70 -c 2 \
    -f -o gpurun_out/r02_leaf_full python tools/one_eval.py 8192 1 > gpurun_out/r02_ncu_leaf.log 2>&1
echo "leaf full rc $?"
ls -la gpurun_out/r02_* | head
