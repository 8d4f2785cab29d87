#!/bin/bash
set -u
for e in 0.003 0.004 0.005 0.006; do python bench.py --no-cpu --no-extra --no-e2e --steps 10 --warmup 3 --eps $e 2>/dev/null | python profiles/experiments/bench_brief.py "chrom eps=$e"; done
for e in 0.012 0.015 0.018 0.021; do python bench.py --workload poly --no-cpu --no-e2e --steps 10 --warmup 3 --eps $e 2>/dev/null | python profiles/experiments/bench_brief.py "poly eps=$e"; done
for e in 0.001 0.0015 0.002 0.0025; do python bench.py --workload chromatin5k --no-cpu --no-e2e --steps 3 --warmup 3 --eps $e 2>/dev/null | python profiles/experiments/bench_brief.py "5k eps=$e"; done
