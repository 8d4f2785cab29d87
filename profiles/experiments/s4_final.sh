#!/bin/bash
# last session of round 2: GPU test suite, smoke, default bench line; launch list of ONE host call selected by its
# NVTX range (the C ABI's tracing); full capture of the algebraic-contact kernel
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest_gpu.log 2>&1; tail -2 gpurun_out/s4_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/s4_bench_1gpu.json 2> gpurun_out/s4_bench_1gpu.err; echo "bench rc=$?"
ncu --nvtx --nvtx-include "binfb_hmc_run_host/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_nvtx_hmc_run_host.csv python -c "import __graft_entry__ as g; g.smoke()" \
    > gpurun_out/ncu_nvtx.log 2>&1; echo "nvtx launch list rc=$?"; tail -4 gpurun_out/r2_launches_nvtx_hmc_run_host.csv | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:chrom_kernel -s 3 -c 1 -o gpurun_out/prof_r2_chrom_alg \
    python bench.py --no-cpu --no-e2e --no-extra --contact algebraic --steps 3 --warmup 3 > gpurun_out/ncu_alg.log 2>&1; echo "alg rc=$?"
python profiles/ncu_summary.py gpurun_out/prof_r2_chrom_alg.ncu-rep 30 > gpurun_out/r2_chrom_alg.ncu_summary.txt 2>&1
head -24 gpurun_out/r2_chrom_alg.ncu_summary.txt
