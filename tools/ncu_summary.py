"""Summarises an .ncu-rep (read here, without a GPU): key raw metrics, warp-stall breakdown and the per-source-line
executed-instruction counts of one kernel.   python tools/ncu_summary.py gpurun_out/x.ncu-rep [source.cu] [top N] > profiles/x.txt"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
for v in vals:
    d = dict(zip(hdr, v))
    u = dict(zip(hdr, units))
    print("Kernel Name:", d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print(f"{k}: {d[k]} {u.get(k, '')}")
    stalls = sorted(((float(d[k].replace(",", "")), k) for k in hdr if "issue_stalled" in k and "per_issue_active" in k and d[k]), reverse=True)
    print("\nwarp stall reasons (warps stalled per issue-active cycle):")
    for x, k in stalls[:9]:
        print(f"   {x:5.2f}  {k}")
    break
if src:
    # per CUDA source line: the "cuda,sass" view carries, on every source-line row, the totals of the SASS rows under it
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    agg = {}
    fname, head = "", None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            head = r
        elif head and len(r) == len(head) and r[0].isdigit():
            d = dict(zip(head, r))
            try:
                n = int(d["Instructions Executed"].replace(",", "") or 0)
                t = int(d["Thread Instructions Executed"].replace(",", "") or 0)
                sm = int(d["# Samples"].replace(",", "") or 0)
                wf = int((d.get("L1 Wavefronts Shared") or "0").replace(",", "") or 0)
                wi = int((d.get("L1 Wavefronts Shared Ideal") or "0").replace(",", "") or 0)
            except (ValueError, KeyError):
                continue
            k = (fname, int(r[0]), r[1])
            v = agg.setdefault(k, [0, 0, 0, 0, 0])
            v[0] += n; v[1] += t; v[2] += sm; v[3] += wf; v[4] += wi
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[2] for v in agg.values()) or 1
    tw = sum(v[3] for v in agg.values()) or 1
    print(f"\nper-source-line executed warp instructions (top {top} of {tot}); stall = share of warp-stall samples; smem = share of shared-memory wavefronts (x = wavefronts / ideal)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        x = f"{v[3] / v[4]:.1f}x" if v[4] else "   -"
        print(f"{100 * v[0] / tot:5.1f}%  eff {v[1] / (32 * v[0]) if v[0] else 0:4.2f}  stall {100 * v[2] / ts:4.1f}%  smem {100 * v[3] / tw:4.1f}% {x:>5}  {k[0]}:{k[1]}  {k[2].strip()[:110]}")
