#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_chromatin.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -4
for a in "" "--chrom-sets 0" "" "--chrom-sets 0"; do
  python bench.py --no-cpu --no-extra --steps 6 --warmup 3 $a 2>/dev/null | python profiles/experiments/bench_brief.py "sets[$a]"
done
for a in "" "--chrom-sets 0"; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:chrom_kernel -s 3 -c 1 python bench.py --no-cpu --no-e2e --no-extra --steps 3 --warmup 3 $a 2>&1 | grep -E "dram__|gpu__time|hit_rate" | sed "s/^/[$a] /"
done
