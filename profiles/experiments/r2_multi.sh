#!/bin/bash
# N-GPU checks: NCCL label-swap self-test, then the default bench line under torchrun
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/rex_nccl_selftest.py > gpurun_out/r2_rex_selftest_$N.log 2>&1
echo "selftest rc=$?"; tail -4 gpurun_out/r2_rex_selftest_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_$N.json 2> gpurun_out/r2_bench_$N.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_$N.json").read().strip().split("\n") if l.startswith("{")][-1])
except Exception as e:
    print("no bench line:", e); raise SystemExit
def show(tag, x):
    if not x: print(tag, x); return
    if "error" in x: print(tag, "ERROR", x["error"]); return
    r = x.get("roofline") or {}
    print(tag, "ms/step %.4f" % x["ms_per_step"], "value %.4g" % x["value"], "frac %.4f" % r.get("frac", float("nan")),
          "acc", x.get("acceptance_rate"), "e2e", (x.get("e2e") or {}).get("value"), "swap", x.get("swap_rates"), x.get("swap_overhead_frac"))
show("HEAD", d)
for k, v in (d.get("extra") or {}).items():
    show(k, v)
if (d.get("extra") or {}).get("rex") and "config" in d["extra"]["rex"]:
    print("ladder", d["extra"]["rex"]["config"]["ladder"], d["extra"]["rex"]["config"]["collective"])
PY
