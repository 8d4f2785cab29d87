#!/bin/bash
# round 2: GPU tests, then the default bench line (headline + extra legs), brief summary on stdout
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; tail -${TAILN:-12} gpurun_out/r2_pytest.log
python bench.py ${BENCH_ARGS:---no-cpu --steps 6 --warmup 3} > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/r2_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench.json").read().strip().split("\n")[-1])
except Exception as e:
    print("no bench line:", e); raise SystemExit
def show(tag, x):
    if not x: print(tag, x); return
    if "error" in x: print(tag, "ERROR", x["error"]); return
    r = x.get("roofline") or {}
    print(tag, "ms/step %.4f" % x["ms_per_step"], "value %.4g" % x["value"], "frac %.4f" % r.get("frac", float("nan")),
          "acc", x.get("acceptance_rate"), "e2e", (x.get("e2e") or {}).get("value"), "clk", (x.get("clocks") or {}).get("sm_mhz"),
          "swap", x.get("swap_rates"), x.get("swap_overhead_frac"))
show("HEAD", d)
for k, v in (d.get("extra") or {}).items():
    show(k, v)
if (d.get("extra") or {}).get("rex") and "config" in d["extra"]["rex"]:
    print("ladder", d["extra"]["rex"]["config"]["ladder"])
PY
