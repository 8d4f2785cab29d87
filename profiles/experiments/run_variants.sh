#!/bin/bash
# A/B timing of experiment variants (built by build_variants.py) on one GPU box.
#   bash profiles/experiments/run_variants.sh "v1 v2 ..." [bench args]
# Each variant replaces binf_b200/libbinf_b200.so for its run; the original is restored at the end.
set -u
VARS="$1"; shift
mkdir -p gpurun_out
cp binf_b200/libbinf_b200.so /tmp/lib_orig.so
for v in $VARS; do
  cp build/variants/lib_$v.so binf_b200/libbinf_b200.so
  echo "=== $v" >> gpurun_out/variants.txt
  timeout 300 python bench.py --no-cpu --no-e2e --no-extra --no-equilibrate --steps 6 --warmup 3 "$@" 2> gpurun_out/variant_$v.err | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('%s ms_per_step=%.3f value=%.4g frac=%.4f acc=%.4f' % (d['config']['workload'], d['ms_per_step'], d['value'], d['roofline']['frac'], d.get('acceptance_rate', -1)))
" >> gpurun_out/variants.txt
done
cp /tmp/lib_orig.so binf_b200/libbinf_b200.so
cat gpurun_out/variants.txt
