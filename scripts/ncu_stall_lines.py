"""Top CUDA source lines for one stall reason of an .ncu-rep (e.g. stall_long_sb, stall_barrier, stall_wait).

    python scripts/ncu_stall_lines.py <rep> <stall column> [top_n]
"""
import csv, io, subprocess, sys
rep, col = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
hdr = None
lines = {}
tot_all = 0
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Line No":
        hdr = r; ic = hdr.index(col); isamp = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        v = int(r[ic] or 0); s = int(r[isamp] or 0)
    except ValueError:
        continue
    lines[(int(r[0]), r[1].strip())] = lines.get((int(r[0]), r[1].strip()), 0) + v
    tot_all += s
tot = sum(lines.values())
print(f"{col}: {tot} samples = {100 * tot / max(tot_all, 1):.1f}% of all")
for (ln, src), v in sorted(lines.items(), key=lambda kv: -kv[1])[:top]:
    print(f"L{ln:4d} {100 * v / max(tot, 1):5.1f}%  {src[:100]}")
