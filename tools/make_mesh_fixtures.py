#!/usr/bin/env python3
"""Convert the reference's text meshes (data/obj/text/*.obj) into the binary `.rtmesh` fixtures
under assets/meshes/.  Run in the build container, where /root/reference exists; the GPU box has
no /root/reference, so tests, bench.py and smoke() read the committed fixtures instead.

Format (little endian): 8-byte magic b"RTMESH1\\0", u32 n_positions, u32 n_normals, u32 n_faces,
f32 positions[3*n_positions], f32 normals[3*n_normals], i32 face_v[3*n_faces], i32 face_vn[3*n_faces]
(0-based indices; corner order as in the OBJ `f` lines, which are all triangles).
Floats are parsed exactly like tobj does (`str::parse::<f32>`, correctly rounded) via numpy.float32.
"""
import struct
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/data/obj/text")
OUT = Path(__file__).resolve().parent.parent / "assets" / "meshes"


def convert(name: str) -> None:
    pos, nrm, fv, fn = [], [], [], []
    for line in (REF / f"{name}.obj").read_text().splitlines():
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            pos.append([np.float32(x) for x in p[1:4]])
        elif p[0] == "vn":
            nrm.append([np.float32(x) for x in p[1:4]])
        elif p[0] == "f":
            corners = [c.split("/") for c in p[1:]]
            assert len(corners) == 3, "fixtures expect triangulated meshes"
            fv.append([int(c[0]) - 1 for c in corners])
            fn.append([int(c[2]) - 1 for c in corners])
    pos = np.asarray(pos, np.float32)
    nrm = np.asarray(nrm, np.float32)
    fv = np.asarray(fv, np.int32)
    fn = np.asarray(fn, np.int32)
    OUT.mkdir(parents=True, exist_ok=True)
    with open(OUT / f"{name}.rtmesh", "wb") as f:
        f.write(b"RTMESH1\0")
        f.write(struct.pack("<III", len(pos), len(nrm), len(fv)))
        f.write(pos.tobytes())
        f.write(nrm.tobytes())
        f.write(fv.tobytes())
        f.write(fn.tobytes())
    print(f"{name}: {len(pos)} positions, {len(nrm)} normals, {len(fv)} faces")


if __name__ == "__main__":
    if not REF.exists():
        sys.exit("reference not mounted; fixtures are already committed")
    for n in ("text", "text_lowres"):
        convert(n)
