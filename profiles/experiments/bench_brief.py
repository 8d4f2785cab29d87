"""print ms_per_step / roofline fraction / acceptance of a bench.py JSON line read from stdin"""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        print(tag, d["config"]["workload"], "ms_per_step=%.3f" % d["ms_per_step"], "frac=%.4f" % d["roofline"]["frac"],
              "acc=%s" % d.get("acceptance_rate"))
