#!/bin/bash
# Leaf evidence of round 2 (B200_PROFILING.md recipe): phase stamps (tools/leaf_prof), plain run, launch list of one C2
# evaluation, then a --set full capture of two leaf launches of the second evaluation.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./tools/leaf_prof > gpurun_out/r02_leaf_phases.txt 2>&1
python tools/one_eval.py 8192 1 > gpurun_out/r02_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_c2_lml_grad.csv python tools/one_eval.py 8192 1 > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc $?"
ncu --set full --clock-control none --import-source on -k regex:leaf_potrf_inv -s 70 -c 2 -f -o gpurun_out/r02_leaf_full python tools/one_eval.py 8192 1 > gpurun_out/r02_ncu_leaf.log 2>&1
echo "leaf full rc $?"
ls -la gpurun_out/r02_leaf_full* gpurun_out/r02_launches_c2_lml_grad.csv
