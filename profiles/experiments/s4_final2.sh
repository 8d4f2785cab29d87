#!/bin/bash
# final state of round 2: GPU test suite, smoke, default bench line, ncu launch list of the same command
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest_gpu.log 2>&1; tail -2 gpurun_out/s4_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/s4_bench_1gpu_c.json 2> gpurun_out/s4_bench_1gpu_c.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s4_ref_1gpu.json 2>/dev/null; echo "ref rc=$?"; cut -c1-160 gpurun_out/s4_ref_1gpu.json
python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > /dev/null 2>&1; echo "short bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
grep -c chrom_kernel gpurun_out/r2_launches_bench.csv
