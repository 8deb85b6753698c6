"""Host-side scene API: thin Python wrappers over the C++ mirror of the reference's `Scene`
(csrc/host/rt_host.hpp, exported through include/rt_host.h).

Reference interfaces mirrored (paths relative to the reference repo):
  Scene::{new, add_sphere, add_triangle, add_light, from_obj, backface_culling}  src/scene/scene.rs:29-178
  BoundedPlane::with_material(..).to_basic_geometries()                         src/geometry/composite/bounded_plane.rs:49-216
  Material::new / TransmissionProperties::{new, new_with_boost, none}           src/raytracing/material.rs:21-74,142-154
  the entry-point scenes of src/main.rs, examples/test_scene.rs, examples/test_text.rs
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from . import _lib as L

MESH_DIR = L.REPO_ROOT / "assets" / "meshes"


def material(color, metallic=0.0, shininess=0.0, *, opacity=None, ior=1.0, boost=0.0) -> np.ndarray:
    """Material::new(color, metallic, shininess, transmission).

    `opacity=None` with the default ior 1.0 is TransmissionProperties::default(); pass `ior=0.0`
    for TransmissionProperties::none(); a float opacity is TransmissionProperties::new[_with_boost].
    """
    m = np.zeros((), L.MATERIAL_DTYPE)
    m["color"] = color
    m["metallic"] = metallic
    m["shininess"] = shininess
    m["ior"] = ior
    m["opacity"] = 0.0 if opacity is None else opacity
    m["has_opacity"] = 0 if opacity is None else 1
    m["boost"] = boost
    return m


@dataclass
class Mesh:
    """A triangulated mesh with per-corner position and normal indices (what tobj hands over)."""

    positions: np.ndarray  # (n,3) f32
    normals: np.ndarray  # (m,3) f32
    face_v: np.ndarray  # (f,3) i32
    face_vn: np.ndarray  # (f,3) i32

    @staticmethod
    def load_rtmesh(path: Path) -> "Mesh":
        raw = Path(path).read_bytes()
        if raw[:8] != b"RTMESH1\0":
            raise ValueError(f"{path}: not an rtmesh file")
        n_pos, n_nrm, n_f = struct.unpack_from("<III", raw, 8)
        off = 20
        pos = np.frombuffer(raw, "<f4", 3 * n_pos, off).reshape(n_pos, 3).copy()
        off += 12 * n_pos
        nrm = np.frombuffer(raw, "<f4", 3 * n_nrm, off).reshape(n_nrm, 3).copy()
        off += 12 * n_nrm
        fv = np.frombuffer(raw, "<i4", 3 * n_f, off).reshape(n_f, 3).copy()
        off += 12 * n_f
        fn = np.frombuffer(raw, "<i4", 3 * n_f, off).reshape(n_f, 3).copy()
        return Mesh(pos, nrm, fv, fn)

    @staticmethod
    def text(high_quality_model: bool) -> "Mesh":
        """The reference's data/obj/text/text.obj (or text_lowres.obj) as committed fixtures."""
        return Mesh.load_rtmesh(MESH_DIR / ("text.rtmesh" if high_quality_model else "text_lowres.rtmesh"))

    @staticmethod
    def empty() -> "Mesh":
        return Mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), np.zeros((0, 3), np.int32))


@dataclass
class FlatScene:
    """The arrays that cross the C ABI (rt_scene_desc). Keeps them alive for the ctypes struct."""

    triangles: np.ndarray
    spheres: np.ndarray
    materials: np.ndarray
    lights: np.ndarray

    def desc(self) -> L.SceneDesc:
        d = L.SceneDesc()
        d.triangles = self.triangles.ctypes.data if len(self.triangles) else None
        d.n_triangles = len(self.triangles)
        d.spheres = self.spheres.ctypes.data if len(self.spheres) else None
        d.n_spheres = len(self.spheres)
        d.materials = self.materials.ctypes.data if len(self.materials) else None
        d.n_materials = len(self.materials)
        d.lights = self.lights.ctypes.data if len(self.lights) else None
        d.n_lights = len(self.lights)
        return d

    @property
    def n_primitives(self) -> int:
        return len(self.triangles) + len(self.spheres)


def _f3(v) -> np.ndarray:
    return np.ascontiguousarray(v, np.float32).reshape(3)


class Scene:
    """Scene<Vec3> (src/scene/scene.rs:24-27)."""

    def __init__(self, handle=None):
        self._lib = L.load()
        self._h = handle if handle is not None else self._lib.rth_scene_new()
        if not self._h:
            raise MemoryError("rth_scene_new failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.rth_scene_free(h)

    # -- construction -------------------------------------------------------------------------
    def add_sphere(self, center, radius: float, mat: np.ndarray) -> None:
        L.check(self._lib.rth_scene_add_sphere(self._h, L.ptr(_f3(center)), radius, L.ptr(np.ascontiguousarray(mat))))

    def add_triangle(self, v1, v2, v3, mat: np.ndarray, normal=None) -> None:
        m = np.ascontiguousarray(mat)
        if normal is None:
            L.check(self._lib.rth_scene_add_triangle(self._h, L.ptr(_f3(v1)), L.ptr(_f3(v2)), L.ptr(_f3(v3)), L.ptr(m)))
        else:
            L.check(self._lib.rth_scene_add_triangle_with_normal(
                self._h, L.ptr(_f3(v1)), L.ptr(_f3(v2)), L.ptr(_f3(v3)), L.ptr(_f3(normal)), L.ptr(m)))

    def add_bounded_plane(self, normal, center, up, width: float, height: float, depth: float, mat: np.ndarray) -> None:
        L.check(self._lib.rth_scene_add_bounded_plane(
            self._h, L.ptr(_f3(normal)), L.ptr(_f3(center)), L.ptr(_f3(up)), width, height, depth, L.ptr(np.ascontiguousarray(mat))))

    def add_light(self, position, color, intensity: float) -> None:
        L.check(self._lib.rth_scene_add_light(self._h, L.ptr(_f3(position)), L.ptr(_f3(color)), intensity))

    def add_obj(self, path, translation=(0, 0, 0), euler_rpy=(0, 0, 0), scale=1.0) -> None:
        L.check(self._lib.rth_scene_add_obj(self._h, str(path).encode(), L.ptr(_f3(translation)), L.ptr(_f3(euler_rpy)), scale))

    def add_mesh(self, mesh: Mesh, mat: np.ndarray | None, translation=(0, 0, 0), euler_rpy=(0, 0, 0), scale=1.0) -> None:
        L.check(self._lib.rth_scene_add_mesh(
            self._h, L.ptr(mesh.positions), len(mesh.positions), L.ptr(mesh.normals) if len(mesh.normals) else None,
            len(mesh.normals), L.ptr(mesh.face_v), L.ptr(mesh.face_vn) if len(mesh.normals) else None, len(mesh.face_v),
            L.ptr(np.ascontiguousarray(mat)) if mat is not None else None, L.ptr(_f3(translation)), L.ptr(_f3(euler_rpy)), scale))

    def backface_culling(self, view_direction=(0, 0, 1)) -> None:
        L.check(self._lib.rth_scene_backface_culling(self._h, L.ptr(_f3(view_direction))))

    # -- flattening ---------------------------------------------------------------------------
    def flatten(self) -> FlatScene:
        nt, ns, nm, nl = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        L.check(self._lib.rth_scene_counts(self._h, C.byref(nt), C.byref(ns), C.byref(nm), C.byref(nl)))
        fs = FlatScene(np.zeros(nt.value, L.TRIANGLE_DTYPE), np.zeros(ns.value, L.SPHERE_DTYPE),
                       np.zeros(nm.value, L.MATERIAL_DTYPE), np.zeros(nl.value, L.LIGHT_DTYPE))
        L.check(self._lib.rth_scene_flatten(self._h, L.ptr(fs.triangles), L.ptr(fs.spheres), L.ptr(fs.materials), L.ptr(fs.lights)))
        return fs


def example_scene(name: str, features: str = "") -> Scene:
    """The scene of a reference entry point: "semesterbild"/"main", "test_scene", "test_text"."""
    lib = L.load()
    cfg = L.Config()
    hq, sbc, rnd = C.c_int(), C.c_int(), C.c_int()
    L.check(lib.rth_config_from_features(features.encode(), C.byref(cfg), C.byref(hq), C.byref(sbc), C.byref(rnd)))
    if name == "test_scene":
        mesh = Mesh.empty()
    elif name == "test_text":
        mesh = Mesh.text(True)
    else:
        mesh = Mesh.text(bool(hq.value))
    h = C.c_void_p()
    L.check(lib.rth_example_scene(name.encode(), features.encode(), L.ptr(mesh.positions), len(mesh.positions),
                                  L.ptr(mesh.normals), len(mesh.normals), L.ptr(mesh.face_v), L.ptr(mesh.face_vn),
                                  len(mesh.face_v), C.byref(h)))
    return Scene(h)
