"""The SASS of the per-ray DFS loop (Walker::step, csrc/hdt_device.cuh) of a traversal kernel, with its size.

    python scripts/sass_loop.py [lib.so] [kernel-name-substring ...] > profiles/r2_sass_walker_step.txt

For every kernel whose mangled name contains the substring: the innermost backward branch that encloses an intersection
mask (>= 70 instructions) is taken as the loop of the resumed, tame walk -- the one nearly every warp runs -- and printed
with per-class instruction counts.  cuobjdump must be on PATH."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "hashdag_b200/libhashdag_b200.so"
pats = sys.argv[2:] or ["trace_paths_kernelIN3hdt19HashDagResolvedDevTILb1", "trace_shadows_kernelIN3hdt19HashDagResolvedDevTILb1"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if not any(p in name for p in pats):
        continue
    ins = []
    for l in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?)\s*;?\s*/\*", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    loops = []
    for a, t in ins:
        m = re.match(r"(?:@!?U?P\d+\s+)?BRA 0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            n = (a - int(m.group(1), 16)) // 16 + 1
            if n >= 70:
                loops.append((n, int(m.group(1), 16), a))
    if not loops:
        continue
    n, lo, hi = min(loops)   # the resumed tame walk is the first (and smallest) of the three inlined walks
    body = [(a, t) for a, t in ins if lo <= a <= hi]
    cls = collections.Counter()
    for _, t in body:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
        cls[op] += 1
    print(f"== {name}\n== DFS loop {lo:#06x}..{hi:#06x}: {n} instructions per iteration (all blocks: ascent, fetch incl. leaf / in-leaf branches, mask)")
    print("== by opcode: " + ", ".join(f"{k} {v}" for k, v in cls.most_common()))
    for a, t in body:
        print(f"{a:04x}  {t}")
    print()
