#!/bin/bash
# round 2: launch list of the default bench line, then one `ncu --set full` capture per dominant kernel
# (each only after the same command has run without ncu).  Summaries -> gpurun_out/, copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
prof() {  # name, kernel regex, skip, bench args...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/prof_$name \
      python bench.py --no-cpu --no-e2e --no-extra "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name rc=$?"
  python profiles/ncu_summary.py gpurun_out/prof_$name.ncu-rep 30 > gpurun_out/$name.ncu_summary.txt 2>&1
  head -22 gpurun_out/$name.ncu_summary.txt
}
prof r2_chrom chrom_kernel 3 --steps 3 --warmup 3
prof r2_chrom5k chrom_kernel 2 --workload chromatin5k --steps 2 --warmup 1
prof r2_poly poly_hmc_kernel 3 --workload poly --steps 3 --warmup 3
prof r2_sink sink_push_kernel 3 --workload sink --steps 3 --warmup 3
prof r2_generic_mid gen_traj_mid 3 --workload generic --steps 3 --warmup 3
