"""Summarise an .ncu-rep (run where ncu is installed): key raw metrics + stall breakdown.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.avg"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print("%-70s %-14s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, ns = {s: 0 for s in stalls}, 0
for r in data:
    try:
        ns += int(r[ix["# Samples"]])
    except Exception:
        continue
    for s in stalls:
        try:
            tot[s] += int(r[ix[s]])
        except Exception:
            pass
print("samples", ns)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
    print("  %-26s %9d %5.1f%%" % (s, v, 100.0 * v / ns))
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:ntop]
for r in top:
    st = {s: int(r[ix[s]] or 0) for s in stalls}
    big = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(r[ix["Address"]][-5:], r[ix["Source"]][:60].ljust(60), r[ix["# Samples"]].rjust(7),
          r[ix["Instructions Executed"]].rjust(11), big)
