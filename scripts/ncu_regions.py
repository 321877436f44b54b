"""Share of executed warp instructions / stall samples per marked code region of mcl_philox.cu."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
src = open("mcluminescence_b200/csrc/mcl_philox.cu").read().split("\n")
# region boundaries: lines that start a region (first match wins, in file order)
marks = [("warp searches", r"^struct Holes"),
         ("CTA barrier (cta_sync)", r"^__device__ __forceinline__ void cta_sync"), ("kernel setup", r"^template <int NT, int MINB, typename NearT, int PPC, bool SLAB_SMEM = false, bool REGRID = true>"), ("seed holes", r"Box.seed \(engine.py:124-129\): holes"),
         ("seed electrons + sort", r"electrons, stored in grid-cell order"), ("K-nearest init", r"^// Seeding, part 3 \(Box._rebuild"), ("K-nearest init (call site)", r"Box._rebuild \(engine.py:113-119\): the KC nearest holes of every electron \(kept"),
         ("leg setup", r"per-replica constants of the rate law"), ("step top", r"// ---------------- loop condition"),
         ("sweep", r"per-electron clocks \+ running argmin"), ("reduce+B1", r"// warp argmin -> one row per warp"),
         ("scalar/dt", r"filling clock \(tl_trap_lab.py:53-60\) and dt"), ("histogram", r"fused occupancy histogram"),
         ("event: remove", r"Box.remove_pair \(engine.py:154-175\)"), ("scan+retarget", r"Which of MY other electrons"),
         ("compaction", r"compaction: keep tombstones"), ("fill", r"Box.add_electron \(engine.py:133-152\)"),
         ("record/tail", r"record \(simulate.py:64,85-89\)"),
         ("pipeline: clocks (sweep team + re-evaluations)", r"The specialised loop, PIPELINED"), ("pipeline: sweep team", r"===== the sweep team"),
         ("pipeline: decision warp", r"===== the decision warp"), ("pipeline: way out (flush, masks)", r"on the way out: every electron"),
         ("leg driver", r"const bool fast_ok = ")]
starts = []
for name, pat in marks:
    for i, l in enumerate(src):
        if re.search(pat, l):
            starts.append((i + 1, name)); break
starts.sort()
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; agg = {}; cur_file = ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1] if len(r) > 1 else ""; continue
    if r and r[0] == "Line No":
        hdr = r; isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        ln, s_, i_ = int(r[0]), int(r[isamp] or 0), int(r[iex] or 0)
    except ValueError:
        continue
    if cur_file.endswith("mcl_rng.cuh"):
        a = agg.setdefault("helpers/philox", [0, 0]); a[0] += s_; a[1] += i_
        continue
    if not cur_file.endswith("mcl_philox.cu"):
        a = agg.setdefault("(cuda headers: sync/shuffle/atomics)", [0, 0]); a[0] += s_; a[1] += i_
        continue
    name = "before"
    for st, nm in starts:
        if ln >= st: name = nm
    a = agg.setdefault(name, [0, 0]); a[0] += s_; a[1] += i_
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
for st, nm in [(0, "before"), (0, "helpers/philox")] + starts + [(0, "(cuda headers: sync/shuffle/atomics)")]:
    if nm in agg: print(f"{nm:50s} inst {100*agg[nm][1]/ti:5.1f}%   samples {100*agg[nm][0]/ts:5.1f}%")
