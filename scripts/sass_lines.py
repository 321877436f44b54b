"""Static SASS statistics per CUDA source line of one kernel: instructions, local-memory (spill) loads / stores.
usage: nvdisasm -g -c <cubin> > x.sass; python scripts/sass_lines.py x.sass <kernel name substring> [min line] [max line]"""
import re, sys, collections
path, key = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
inst = collections.Counter(); spill = collections.Counter(); cur = None; infn = False
for ln in open(path):
    if ln.startswith("\t.section") or ln.startswith(".section"):
        infn = False
    m = re.match(r"\s*\.global\s+(\S+)", ln) or re.match(r"\s*\.type\s+(\S+),@function", ln)
    if m:
        infn = key in m.group(1) and "philox_kernel" in m.group(1)
    if not infn:
        continue
    m = re.search(r'//## File "[^"]*mcl_philox\.cu", line (\d+)', ln)
    if m:
        cur = int(m.group(1)); continue
    if re.search(r'//## File "', ln):
        cur = -1; continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        inst[cur] += 1
        if re.search(r"\b(STL|LDL)\b", ln):
            spill[cur] += 1
tot = sum(inst.values())
print("total instructions", tot, "spill instrs", sum(spill.values()))
for l in sorted(inst):
    if l is not None and lo <= l <= hi:
        print(f"L{l:5d} inst {inst[l]:5d} spill {spill[l]:3d}")
