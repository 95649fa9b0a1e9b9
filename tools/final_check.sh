#!/bin/bash
# Round-end evidence on one GPU (run under gpurun): tests, smoke, the two bench arms, the ncu launch list of the
# bench command (after the same command exited 0 without ncu).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/final_pytest.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/final_smoke.log)"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "reference arm rc=$?"
python bench.py --steps 20 --warmup 5 --storage kron --shapes 56,56,56,56 > gpurun_out/r02_bench_kron56_n1.json 2> gpurun_out/r02_bench_kron56_n1.err; echo "kron bench rc=$?"
CMD="python bench.py --steps 5 --warmup 3 --no-solve --no-cpu --no-configs"
$CMD > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
