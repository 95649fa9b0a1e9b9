#!/bin/bash
# ncu evidence (run under gpurun, one GPU).  Each ncu run follows a plain run of the same command.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-solve --no-cpu"
$CMD > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_dense_apply -s 3 -c 2 -o gpurun_out/prof_dense -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "dense capture rc=$?"
CMD2="python tools/sweep_step.py 4096"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sweep_gemm -s 1 -c 1 -o gpurun_out/prof_sweep -f $CMD2 > gpurun_out/ncu_sweep.log 2>&1
echo "sweep capture rc=$?"
cat gpurun_out/plain3.log | tail -1
