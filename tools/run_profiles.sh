#!/bin/bash
# ncu evidence for the bench command (run under gpurun, one GPU).
# 1) launch list with device time per launch; 2) one full capture of the dense pass kernel.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-solve --no-cpu"
$CMD > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_dense_apply -s 3 -c 2 -o gpurun_out/prof_dense -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full.log
