"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the product arm's config of
the workload, the keys the driver reads, and a cpu_baseline describing the bounded sample it timed."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_workload_configs_and_l2_rule():
    b = _bench_module()
    for name in ("chromatin", "poly", "chromatin5k"):
        cfg = b.workload_config(name, 1)
        assert set(cfg) == {"workload", "chains_per_gpu", "leapfrog_steps", "timestep", "parallelism", "l2", "gibbs",
                            "equilibration_sweeps"}
        assert cfg["workload"] == b.WORKLOADS[name]["name"] and cfg["leapfrog_steps"] == 20
        # timing rule: inputs larger than L2, or an L2 flush between timed iterations -- and the line says which
        big = b.working_set_bytes(name, cfg["chains_per_gpu"]) >= 130e6
        assert b.needs_l2_flush(name, cfg["chains_per_gpu"]) == (not big)
        assert ("flushed" in cfg["l2"]) == (not big)
    assert not b.needs_l2_flush("chromatin", 4096) and b.needs_l2_flush("chromatin5k", 592)
    assert b.workload_config("chromatin", 8)["parallelism"].startswith("chain-sharded x8")
    assert b.FLOP_PER_PAIR == 31.0 and b.FLOP_PER_DATUM == 14.0


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "poly",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().split("\n") if l.strip()]
    assert len(lines) == 1, r.stdout          # stdout carries exactly the JSON line
    d = json.loads(lines[0])
    b = _bench_module()
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "leapfrog steps/s"
    assert d["config"] == b.workload_config("poly", 1)        # the product arm's config of this workload
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and "single-chain" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 exit without work and without output
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "poly",
                         "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600,
                        env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=ROOT)
    assert r2.returncode == 0 and r2.stdout.strip() == ""
