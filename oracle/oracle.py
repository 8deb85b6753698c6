"""ctypes wrapper of the CPU oracle (oracle/rt_oracle.cpp).  TEST INFRASTRUCTURE ONLY: imported
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
the product package."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle_rt.so"
_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "rt_oracle.cpp"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-s", "-B"], check=True)
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB))
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
        lib.oracle_render.restype = i32
        lib.oracle_render.argtypes = [vp, vp, i32, i32, vp, u32, vp, vp, vp, vp]
        lib.oracle_trace_closest.restype = i32
        lib.oracle_trace_closest.argtypes = [vp, u32, i32, vp, u32, vp, vp]
        lib.oracle_trace_shadow.restype = i32
        lib.oracle_trace_shadow.argtypes = [vp, u32, i32, vp, vp, u32, vp]
        lib.oracle_cloud_index.restype = u32
        lib.oracle_cloud_index.argtypes = [C.c_uint64, u32, u32, u32, u32]
        lib.oracle_encode.restype = u32
        lib.oracle_encode.argtypes = [C.c_float, C.c_float, C.c_float]
        _lib = lib
    return _lib


BRUTE, BVH, PACKET8 = 0, 1, 2


def render(desc, cfg, mode: int = BVH, threads: int = 0, pixels: np.ndarray | None = None, want_rgb: bool = False):
    """desc / cfg: ctypes structs byte-compatible with rt_scene_desc / rt_config.
    Returns (argb[H*W] u32, mask[H*W] u8, rgb[H*W,3] f32 | None, counters dict)."""
    lib = load()
    n = cfg.width * cfg.height
    argb = np.zeros(n, np.uint32)
    mask = np.zeros(n, np.uint8)
    rgb = np.zeros((n, 3), np.float32) if want_rgb else None
    cnt = np.zeros(6, np.uint64)
    pix = None if pixels is None else np.ascontiguousarray(pixels, np.uint32)
    rc = lib.oracle_render(C.addressof(desc), C.addressof(cfg), mode, threads,
                           None if pix is None else pix.ctypes.data, 0 if pix is None else len(pix),
                           argb.ctypes.data, mask.ctypes.data, None if rgb is None else rgb.ctypes.data, cnt.ctypes.data)
    if rc:
        raise RuntimeError(f"oracle_render failed: {rc}")
    counters = dict(zip(["closest_rays", "shadow_rays", "triangle_tests", "sphere_tests", "box_tests", "shadow_rays_facing_away"], map(int, cnt)))
    return argb, mask, rgb, counters


def trace_closest(desc, rays: np.ndarray, flags: int = 0, mode: int = BRUTE):
    lib = load()
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    t = np.zeros(len(rays), np.float32)
    prim = np.zeros(len(rays), np.uint32)
    lib.oracle_trace_closest(C.addressof(desc), flags, mode, rays.ctypes.data, len(rays), t.ctypes.data, prim.ctypes.data)
    return t, prim


def trace_shadow(desc, rays: np.ndarray, max_dist: np.ndarray, flags: int = 0, mode: int = BRUTE):
    lib = load()
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    md = np.ascontiguousarray(max_dist, np.float32)
    out = np.zeros((len(rays), 5), np.float32)
    lib.oracle_trace_shadow(C.addressof(desc), flags, mode, rays.ctypes.data, md.ctypes.data, len(rays), out.ctypes.data)
    return out
