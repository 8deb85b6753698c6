"""Synthetic large-mesh workload (BASELINE.json configs[4]): the semesterbild spheres, planes and
lights with the text mesh replaced by a procedurally generated multi-million-triangle relief.

The reference has no such scene; it is the scaling workload the task defines ("extreme_quality at
high_resolution on a synthetic multi-million-triangle OBJ mesh").  The relief is a displaced
height field facing the camera in front of the back wall.  Its cells are sized so that
2 * area * cos(theta) stays above the reference's `|det| > f32::EPSILON` acceptance threshold
(src/geometry/basic/triangle.rs:194) for near-frontal rays -- a finer mesh would simply be
invisible to the reference algorithm.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .scene import Mesh, Scene, example_scene, material


def relief_mesh(nx: int = 2048, ny: int = 1024, seed: int = 42, amplitude: float = 0.01) -> Mesh:
    """(nx x ny) cells -> 2*nx*ny triangles on the unit square [0,1]^2, height = value noise * amplitude
    along -z (towards the camera); per-vertex normals from the height gradient."""
    rng = np.random.default_rng(seed)
    gx, gy = nx + 1, ny + 1
    # smooth value noise: bilinear upsampling of coarse random lattices at three octaves
    h = np.zeros((gy, gx), np.float64)
    u = np.linspace(0.0, 1.0, gx)
    v = np.linspace(0.0, 1.0, gy)
    for octave, weight in ((8, 0.6), (32, 0.3), (128, 0.1)):
        lat = rng.random((octave + 1, octave + 1))
        fx, fy = u * octave, v * octave
        ix, iy = np.minimum(fx.astype(int), octave - 1), np.minimum(fy.astype(int), octave - 1)
        tx, ty = fx - ix, fy - iy
        tx, ty = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty)
        a = lat[iy][:, ix] * (1 - tx) + lat[iy][:, ix + 1] * tx
        b = lat[iy + 1][:, ix] * (1 - tx) + lat[iy + 1][:, ix + 1] * tx
        h += weight * (a * (1 - ty)[:, None] + b * ty[:, None])
    z = -(h * amplitude)
    X, Y = np.meshgrid(u, v)
    pos = np.stack([X, Y, z], -1).reshape(-1, 3).astype(np.float32)
    dzdx = np.gradient(z, u, axis=1)
    dzdy = np.gradient(z, v, axis=0)
    nrm = np.stack([dzdx, dzdy, -np.ones_like(z)], -1)  # faces -z (the camera)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    nrm = nrm.reshape(-1, 3).astype(np.float32)
    j, i = np.mgrid[0:ny, 0:nx]
    v00 = (j * gx + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + gx, v00 + gx + 1
    # winding chosen so that e1 x e2 also points towards -z
    faces = np.concatenate([np.stack([v00, v01, v10], 1), np.stack([v10, v01, v11], 1)]).astype(np.int32)
    return Mesh(pos, nrm, np.ascontiguousarray(faces), np.ascontiguousarray(faces))


def large_mesh_scene(features: str, nx: int = 2048, ny: int = 1024, seed: int = 42) -> Scene:
    """semesterbild without the text + the relief mesh (2*nx*ny triangles) in front of the back wall."""
    lib = L.load()
    import ctypes as C

    # semesterbild with an EMPTY text mesh = its 9 spheres, 4 bounded planes (48 triangles), 5 lights
    h = C.c_void_p()
    empty = Mesh.empty()
    L.check(lib.rth_example_scene(b"semesterbild", features.encode(), L.ptr(empty.positions), 0, L.ptr(empty.normals), 0,
                                  L.ptr(empty.face_v), L.ptr(empty.face_vn), 0, C.byref(h)))
    scene = Scene(h)
    cfg = L.Config()
    L.check(lib.rth_config_from_features(features.encode(), C.byref(cfg), None, None, None))
    sh = np.float32(cfg.height) / np.float32(cfg.width)
    sd = (1.0 + float(sh)) / 2.0
    mesh = relief_mesh(nx, ny, seed)
    # Lay the relief 0.02 in front of the back wall, in the wall's own frame (src/main.rs:150-177:
    # rotor = from_euler_angles(-0.04, 0.125, 0.51), isometry translation (0.25 SW, 0.002 SH, 0.037 SD),
    # wall centre (0.5 SW, 0.55 SH, SD), size SW x 1.1 SH), then transform it like the wall.
    mesh.positions[:, 0] = 0.01 + 0.98 * mesh.positions[:, 0]
    mesh.positions[:, 1] = (0.01 + 1.08 * mesh.positions[:, 1]) * float(sh)
    mesh.positions[:, 2] = (sd - 0.02) + mesh.positions[:, 2]
    mesh.positions = np.ascontiguousarray(mesh.positions, np.float32)
    scene.add_mesh(mesh, material([0.123533, 0.719221, 0.826410], 0.0, 0.592264),  # text.mtl's Material.001
                   translation=(0.25, 0.002 * float(sh), 0.037 * sd), euler_rpy=(-0.04, 0.125, 0.51), scale=1.0)
    return scene
