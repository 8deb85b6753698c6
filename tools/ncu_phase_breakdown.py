#!/usr/bin/env python3
"""Executed warp instructions of one kernel by call path (who was inlined where).

    ncu -i REP --page source --csv > sass.csv                 (per-SASS-instruction counts of the captured kernel)
    cuobjdump -xelf all librt_b200.so; nvdisasm -gi -c rt_api.sm_100a.cubin > all.sass   (same build!)
    ncu_phase_breakdown.py sass.csv all.sass <mangled kernel name substring> <csrc dir> [depth]

The per-line source view charges an inlined helper's instructions (dot3, operator*, ...) to the helper's own line; this
tool walks nvdisasm's "inlined at" chains instead and charges every instruction to the chain of functions it was inlined
through, cut at `depth` functions below the kernel."""
import csv
import re
import sys
from collections import defaultdict
from pathlib import Path

sass_csv, disasm, kernel, csrc = sys.argv[1:5]
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 2

# ---- function line ranges of the sources
starts = {}
sig = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:RT_DEV|__device__|__global__|static|inline|__host__)[^;{]*?\b([A-Za-z_]\w*)\s*\(")
for f in list(Path(csrc).glob("*.cuh")) + list(Path(csrc).glob("*.cu")):
    lst, pending_template = [], False
    for i, l in enumerate(f.read_text().split("\n"), 1):
        m = sig.match(l)
        if m and not l.rstrip().endswith(";"):
            lst.append((i, m.group(1)))
    starts[f.name] = lst


def func_of(fname, line):
    best = "?"
    for s, n in starts.get(fname, []):
        if s <= line + 1:  # (a template header line sits one above the signature)
            best = n
        else:
            break
    return best


# ---- disassembly: offset -> chain of functions (outermost first)
lines = open(disasm).read().split("\n")
i0 = next(i for i, l in enumerate(lines) if l.startswith("//-") and kernel in l)
i1 = next((i for i in range(i0 + 1, len(lines)) if lines[i].startswith("//-") and ".text." in lines[i]), len(lines))
chain_at, group, in_group = {}, [], False
cur_chain = ()
for l in lines[i0:i1]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if not in_group:
            group, in_group = [], True
        group.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if in_group:
        cur_chain = tuple(func_of(f, n) for f, n in reversed(group))  # outermost first
        in_group = False
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        chain_at[int(m.group(1), 16)] = (cur_chain, m.group(2))

# ---- executed counts
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ii = hdr.index("Instructions Executed")
base = int(rows[2][0], 16)
tot = 0
by = defaultdict(int)
local = defaultdict(int)
for r in rows[2:]:
    if len(r) <= ii or not r[0].startswith("0x"):
        continue
    n = int(r[ii] or 0)
    off = int(r[0], 16) - base
    chain, ins = chain_at.get(off, (("?",), r[1]))
    key = " > ".join(chain[:depth + 1][1:]) or "(kernel body)"
    by[key] += n
    tot += n
    if "LDL" in ins or "STL" in ins:
        local[key] += n
print(f"# {tot} warp instructions; share by call path below the kernel (depth {depth}); [..] = of which local loads/stores")
for k, n in sorted(by.items(), key=lambda x: -x[1]):
    if n * 1000 >= tot:
        print(f"{100 * n / tot:6.2f} %  [{100 * local[k] / tot:5.2f} %]  {k}")
