#!/bin/bash
# LOCKSTEP kernels with progress flags instead of the per-step chain barrier (-DBINFB_LOCKFLAGS=1): parity suite, then
# A/B timing at 512 chains.  Build the variants first:
#   python profiles/experiments/build_variants.py base= lf=-DBINFB_LOCKFLAGS=1
set -u
cp binf_b200/libbinf_b200.so /tmp/lib_orig.so
cp build/variants/lib_lf.so binf_b200/libbinf_b200.so
python -m pytest tests/test_gpu_chromatin.py -x -q -m gpu 2>&1 | tail -3
cp /tmp/lib_orig.so binf_b200/libbinf_b200.so
rm -f gpurun_out/variants.txt
bash profiles/experiments/run_variants.sh "base lf base lf" --chains 512 2>&1 | tail -9
