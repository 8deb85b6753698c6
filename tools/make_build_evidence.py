#!/usr/bin/env python3
"""Regenerates profiles/r2_ptxas_v.txt and profiles/r2_sass_excerpt.txt from the current sources / the built library
(no GPU needed): `-Xptxas -v` per entry function, and per kernel the counts of the SASS instructions the design
claims rest on."""
import collections
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "hslu_i" / "ba_raytracing" / "f2501_raytracer_b200"


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def ptxas_v():
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, RT_LIB_PATH=str(Path(td) / "lib.so"))
        r = subprocess.run([sys.executable, "-m", "hslu_i.ba_raytracing.f2501_raytracer_b200.build", "--force", "--verbose"], cwd=ROOT, env=env,
                           capture_output=True, text=True)
    text = r.stdout + r.stderr
    entries, cur = [], None
    for l in text.split("\n"):
        m = re.search(r"Compiling entry function '([^']+)'", l)
        if m:
            cur = [m.group(1), []]
            entries.append(cur)
        elif cur and ("bytes stack frame" in l or "Used " in l):
            cur[1].append(l.replace("ptxas info    : ", "").strip())
    dm = demangle([e[0] for e in entries])
    lines = ["# nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xptxas -v",
             "# (tools/make_build_evidence.py), entry functions only, names demangled with c++filt", ""]
    for name, info in entries:
        lines.append(re.sub(r"\(.*", "", dm[name]))
        lines += ["    " + i for i in info]
    (ROOT / "profiles" / "r2_ptxas_v.txt").write_text("\n".join(lines) + "\n")


def sass():
    r = subprocess.run(["cuobjdump", "-sass", str(PKG / "librt_b200.so")], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for l in r.split("\n"):
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if cur and m:
            op = m.group(1)
            kernels[cur]["n"] += 1
            for key in ("LDG.E.128.CONSTANT", "LDG.E.128", "LDG.E.CONSTANT", "LDG.E", "LDL", "STL", "LDS", "STS", "STG.E.128", "ATOM", "RED", "MEMBAR", "MUFU",
                        "FFMA", "FMUL", "FADD", "HMMA", "UTCHMMA", "UTMALDG"):
                if op == key or (key in ("LDL", "STL", "LDS", "STS", "ATOM", "RED", "MEMBAR", "MUFU", "FFMA", "FMUL", "FADD", "HMMA", "UTCHMMA", "UTMALDG") and
                                 (op.split(".")[0] == key or (key == "ATOM" and op.startswith("ATOM")))) or (key.startswith("LDG") and op.startswith(key) and
                                                               not any(op.startswith(k2) for k2 in ("LDG.E.128.CONSTANT", "LDG.E.128", "LDG.E.CONSTANT")
                                                                       if len(k2) > len(key))):
                    kernels[cur][key] += 1
                    break
    dm = demangle(list(kernels))
    lines = ["# cuobjdump -sass librt_b200.so (sm_100a), round 2 (tools/make_build_evidence.py): per kernel, the memory instructions that matter for the design claims.",
             "# LDG.E.128.CONSTANT = the 128-bit read-only fetches of BVH nodes / primitive records / materials / cloud points (DESIGN.md section 3);",
             "# LDL / STL = local memory (register spills, cand[] / stack[] arrays); no HMMA / UTC*MMA / UTMA anywhere: no tensor-core or TMA work on this path.", ""]
    for k, c in kernels.items():
        name = re.sub(r"\(.*", "", dm[k])
        always = ("HMMA", "UTCHMMA", "UTMALDG")
        body = ", ".join(f"{key} {c[key]}" for key in ("LDG.E.128.CONSTANT", "LDG.E.128", "LDG.E.CONSTANT", "LDG.E", "LDL", "STL", "LDS", "STS", "STG.E.128", "ATOM", "RED",
                                                        "MEMBAR", "MUFU", "FFMA", "FMUL", "FADD", "HMMA", "UTCHMMA", "UTMALDG") if c[key] or key in always)
        lines.append(f"{name}: {c['n']} instructions; {body}")
    (ROOT / "profiles" / "r2_sass_excerpt.txt").write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    ptxas_v()
    sass()
