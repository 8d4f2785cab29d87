#!/bin/bash
# two half-size persistent CTAs per SM (4 chains x 2 roles each, two-stage rings) against one CTA of 8 chains:
#   python profiles/experiments/build_variants.py base= ns2=-DBINFB_CHROM_NS=2 two=-DBINFB_CHROM_NS=2,-DBINFB_CTAS_PER_SM=2
set -u
rm -f gpurun_out/variants.txt
bash profiles/experiments/run_variants.sh "base ns2" > /dev/null 2>&1
bash profiles/experiments/run_variants.sh "two" --chrom-warps 4 > /dev/null 2>&1
bash profiles/experiments/run_variants.sh "base" > /dev/null 2>&1
cat gpurun_out/variants.txt
