"""Aggregate an ncu report per CUDA source line (needs -lineinfo and --import-source on).
    python profiles/ncu_lines.py gpurun_out/x.ncu-rep [top_n] [kernel-id]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kid = sys.argv[3] if len(sys.argv) > 3 else ":::1"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
print("kernel:", r[hdr.index("Kernel Name")])
for k in keys:
    if k in hdr:
        print(f"  {k:70s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-id", kid],
                     capture_output=True, text=True).stdout
fname = None
h = None
agg = {}
for row in csv.reader(io.StringIO(src)):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        h = row
        continue
    if h is None or not row[0].strip().isdigit():
        continue
    d = dict(zip(h[4:], row[4:]))

    def f(k):
        try:
            return float(d.get(k, "0") or 0)
        except ValueError:
            return 0.0
    key = (fname, int(row[0]))
    old = agg.get(key, (0, 0, 0, 0, ""))
    agg[key] = (old[0] + f("Instructions Executed"), old[1] + f("# Samples"), old[2] + f("L1 Wavefronts Shared"),
                old[3] + f("L1 Wavefronts Shared Excessive"), row[1].strip()[:95])
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
tw = sum(v[2] for v in agg.values()) or 1
print(f"total: {ti:.3e} warp-instructions, {ts:.0f} samples, {tw:.3e} shared wavefronts")
print(" inst%  smpl%   wf%  excessM  file:line  source")
by = 1 if len(sys.argv) > 4 and sys.argv[4] == "samples" else 0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
    print(f"{v[0] / ti * 100:5.1f}  {v[1] / ts * 100:5.1f}  {v[2] / tw * 100:5.1f}  {v[3] / 1e6:6.1f}  {k[0]}:{k[1]:<4d} {v[4]}")
