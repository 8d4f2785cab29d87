#!/bin/bash
# 512 chains per GPU (the replica-exchange shard, BASELINE.json configs[4]): chains per CTA in the small-batch plan
# (4 warps per chain): 4 (default: 128 CTAs on 148 SMs), 3 (171 groups, 12 warps per CTA), 2 (256 groups)
set -x
python -m pytest tests/test_gpu_chromatin.py -x -q -m gpu -k full_size > gpurun_out/s4_fullsize_test.log 2>&1
tail -3 gpurun_out/s4_fullsize_test.log
for w in 0 3 2; do
  python bench.py --workload chromatin --chains 512 --chrom-warps $w --no-extra --no-cpu --no-e2e --steps 10 \
     > gpurun_out/s4_c512_w$w.json 2> gpurun_out/s4_c512_w$w.err
  python -c "import json;d=json.loads(open('gpurun_out/s4_c512_w$w.json').read().strip().splitlines()[-1]);print('warps',$w,d['ms_per_step'],d['roofline']['frac'])"
done
