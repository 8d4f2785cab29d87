#!/bin/bash
# the algebraic contact function (BINFB_FLAG_CONTACT_ALGEBRAIC: 2 MUFU + 21 FMA-pipe ops per pair) against the
# logistic one (3 MUFU + 18) at the headline shape: 4096 chains x 1000 beads, L = 20, fused Gibbs sweep
for c in logistic algebraic; do
  python bench.py --contact $c --no-extra --no-cpu --no-e2e --steps 10 > gpurun_out/s4_contact_$c.json 2> gpurun_out/s4_contact_$c.err
  python -c "import json;d=json.loads(open('gpurun_out/s4_contact_$c.json').read().strip().splitlines()[-1]);print('$c',d['ms_per_step'],d['roofline']['frac'],d['roofline']['sfu']['frac'],d['acceptance_rate'])"
done
