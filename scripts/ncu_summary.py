"""Summarise an .ncu-rep: key metrics, stall reasons, instruction/sample share per code region."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.avg.per_second", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
for k in keys:
    if k in M:
        print(f"{k:75s} {M[k]:>18s} {units[hdr.index(k)]}")
st = sorted(((float(v), h) for h, v in M.items() if "pcsamp_warps_issue_stalled" in h and not h.endswith("not_issued")), reverse=True)
tot = sum(v for v, _ in st)
print("stall samples:", ", ".join(f"{h.split('stalled_')[1]} {100*v/tot:.1f}%" for v, h in st[:10]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
ia, isrc, isamp, iex = h2.index("Address"), h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
data = []
for r in rows[2:]:
    try:
        data.append((r[isrc], int(r[isamp]), int(r[iex])))
    except Exception:
        pass
ts, ti = sum(d[1] for d in data), sum(d[2] for d in data)
print("total warp-instructions", ti, "samples", ts)
chunk = 50
for k in range(0, len(data), chunk):
    seg = data[k:k + chunk]
    s_, i_ = sum(d[1] for d in seg), sum(d[2] for d in seg)
    if s_ > ts * 0.01 or i_ > ti * 0.01:
        print(f"  instr {k:5d}-{k+chunk:5d}: samples {100*s_/ts:5.1f}%  inst {100*i_/ti:5.1f}%   {seg[0][0].strip()[:50]}")
