"""Per-CUDA-source-line share of executed warp instructions and stall samples from an .ncu-rep.

    python scripts/ncu_lines.py <rep> [top_n] [samples]     # third argument: rank by stall samples instead of instructions
"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = {}
cur = None
cur_file = ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = (r[1] if len(r) > 1 else "").split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r
        isamp, iex, ithr = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":          # a CUDA source line header row (aggregated)
        cur = (int(r[0]), (cur_file + ": " if not cur_file.endswith("mcl_philox.cu") else "") + r[1].strip())
        try:
            lines[cur] = [int(r[isamp] or 0), int(r[iex] or 0), int(r[ithr] or 0)]
        except ValueError:
            pass
ts = sum(v[0] for v in lines.values()); ti = sum(v[1] for v in lines.values())
print(f"total samples {ts}  warp-instructions {ti}")
for (ln, src), (s, i, t) in sorted(lines.items(), key=lambda kv: -kv[1][0 if len(sys.argv) > 3 else 1])[:top]:
    print(f"L{ln:4d} inst {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%  thr/inst {t/max(i,1):4.1f}  {src[:90]}")
