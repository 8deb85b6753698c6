"""ctypes bindings of librt_b200.so (include/rt_api.h, include/rt_host.h).

The library is built in-tree by `build.py` (nvcc, sm_100a).  There is NO fallback: if the shared
object is missing or a compute entry point reports no CUDA device, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parents[2]
LIB_PATH = Path(os.environ.get("RT_LIB_PATH", PKG_DIR / "librt_b200.so"))

RT_OK = 0
RT_ERR_INVALID_ARGUMENT = -1
RT_ERR_NO_DEVICE = -2
RT_ERR_CUDA = -3
RT_ERR_OUT_OF_MEMORY = -4
RT_ERR_IO = -5

FLAG_ANTI_ALIASING = 1 << 0
FLAG_AA_ROTATION_SCALE = 1 << 1
FLAG_REFLECTIONS = 1 << 2
FLAG_REFRACTIONS = 1 << 3
FLAG_BACKFACE_CULLING = 1 << 4
FLAG_PACKET_SEMANTICS = 1 << 5
FLAG_COUNT_WORK = 1 << 6
FLAG_ZERO_MISSES = 1 << 7
FLAG_PER_RAY_SHADOWS = 1 << 8

# numpy dtypes byte-compatible with the C structs
MATERIAL_DTYPE = np.dtype(
    [("color", "<f4", 3), ("metallic", "<f4"), ("shininess", "<f4"), ("ior", "<f4"), ("opacity", "<f4"),
     ("has_opacity", "<u4"), ("boost", "<f4")]
)
TRIANGLE_DTYPE = np.dtype([("v1", "<f4", 3), ("v2", "<f4", 3), ("v3", "<f4", 3), ("normal", "<f4", 3), ("material", "<u4")])
SPHERE_DTYPE = np.dtype([("center", "<f4", 3), ("radius", "<f4"), ("material", "<u4")])
LIGHT_DTYPE = np.dtype([("position", "<f4", 3), ("color", "<f4", 3), ("intensity", "<f4")])
assert MATERIAL_DTYPE.itemsize == 36 and TRIANGLE_DTYPE.itemsize == 52
assert SPHERE_DTYPE.itemsize == 20 and LIGHT_DTYPE.itemsize == 28


class SceneDesc(C.Structure):
    _fields_ = [
        ("triangles", C.c_void_p), ("n_triangles", C.c_uint32),
        ("spheres", C.c_void_p), ("n_spheres", C.c_uint32),
        ("materials", C.c_void_p), ("n_materials", C.c_uint32),
        ("lights", C.c_void_p), ("n_lights", C.c_uint32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("flags", C.c_uint32), ("aa_samples_per_pixel", C.c_uint32),
        ("aa_table", C.c_void_p), ("aa_table_len", C.c_uint32), ("light_multiplicator", C.c_uint32),
        ("light_clouds", C.c_void_p), ("n_light_clouds", C.c_uint32), ("reflection_max_depth", C.c_uint32),
        ("refraction_max_depth", C.c_uint32), ("seed", C.c_uint64),
        ("tile_rows", C.c_uint32), ("tile_rank", C.c_uint32), ("tile_world", C.c_uint32), ("tables_epoch", C.c_uint32),
    ]


class SceneInfo(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_slots", C.c_uint32), ("bvh_depth", C.c_uint32), ("builder", C.c_uint32),
                ("build_ms", C.c_double), ("upload_ms", C.c_double), ("device_bytes", C.c_uint64), ("scratch_bytes", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double), ("total_ms", C.c_double), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
        ("box_tests", C.c_uint64), ("triangle_tests", C.c_uint64), ("sphere_tests", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("beam_triangle_tests", C.c_uint64),
        ("prepass_box_tests", C.c_uint64), ("prepass_triangle_tests", C.c_uint64), ("prepass_sphere_tests", C.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


BAND_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32)  # rt_band_callback


class RtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"librt_b200 error {code}: {message}")
        self.code = code


_lib = None


def _declare(lib: C.CDLL) -> None:
    u32, u64, i32, vp, f32 = C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_float
    P = C.POINTER
    sig = {
        "rt_api_version": (i32, []),
        "rt_last_error": (C.c_char_p, []),
        "rt_device_count": (i32, []),
        "rt_measure_fp32_peak": (i32, [i32, P(C.c_double)]),
        "rt_scene_create": (i32, [P(SceneDesc), i32, P(vp)]),
        "rt_scene_destroy": (None, [vp]),
        "rt_scene_get_info": (i32, [vp, P(SceneInfo)]),
        "rt_check_frame": (i32, [vp]),
        "rt_last_kernel": (C.c_char_p, [vp]),
        "rt_render": (i32, [vp, P(Config), vp, vp, P(Stats)]),
        "rt_render_progressive": (i32, [vp, P(Config), vp, vp, u32, vp, vp, P(Stats)]),
        "rt_progressive_overlapped_bands": (u32, [vp]),
        "rt_render_device": (i32, [vp, P(Config), vp, vp, vp, P(Stats)]),
        "rt_frame_constants": (i32, [P(Config), vp, vp, P(u32)]),
        "rt_multi_create": (i32, [P(SceneDesc), vp, u32, P(vp)]),
        "rt_multi_destroy": (None, [vp]),
        "rt_multi_device_count": (u32, [vp]),
        "rt_multi_render": (i32, [vp, P(Config), vp, vp, P(Stats)]),
        "rt_multi_get_info": (i32, [vp, P(SceneInfo), P(u32)]),
        "rt_kernel_times": (i32, [vp, u32, vp, P(u32)]),
        "rt_tile_local_rows": (u32, [P(Config)]),
        "rt_tile_row": (u32, [P(Config), u32]),
        "rt_scatter_tile_rows": (i32, [P(Config), vp, vp, u32, vp]),
        "rt_place_tiles": (i32, [P(Config), vp, u64, vp, vp, vp]),
        "rt_bvh_digest": (i32, [P(SceneDesc), P(u64), P(u32), P(u32)]),
        "rt_trace_closest": (i32, [vp, u32, vp, u32, vp, vp]),
        "rt_trace_shadow": (i32, [vp, u32, vp, vp, u32, vp]),
        "rt_sampler_aa_table": (i32, [u64, u32, i32, vp, u32]),
        "rt_sampler_light_clouds": (i32, [u64, u32, u32, vp]),
        "rt_cloud_index": (u32, [u64, u32, u32, u32, u32]),
        # rt_host.h
        "rth_scene_new": (vp, []),
        "rth_scene_free": (None, [vp]),
        "rth_scene_add_sphere": (i32, [vp, vp, f32, vp]),
        "rth_scene_add_triangle": (i32, [vp, vp, vp, vp, vp]),
        "rth_scene_add_triangle_with_normal": (i32, [vp, vp, vp, vp, vp, vp]),
        "rth_scene_add_bounded_plane": (i32, [vp, vp, vp, vp, f32, f32, f32, vp]),
        "rth_scene_add_light": (i32, [vp, vp, vp, f32]),
        "rth_scene_add_obj": (i32, [vp, C.c_char_p, vp, vp, f32]),
        "rth_scene_add_mesh": (i32, [vp, vp, u32, vp, u32, vp, vp, u32, vp, vp, vp, f32]),
        "rth_scene_backface_culling": (i32, [vp, vp]),
        "rth_scene_counts": (i32, [vp, P(u32), P(u32), P(u32), P(u32)]),
        "rth_scene_flatten": (i32, [vp, vp, vp, vp, vp]),
        "rth_example_scene": (i32, [C.c_char_p, C.c_char_p, vp, u32, vp, u32, vp, vp, u32, P(vp)]),
        "rth_config_from_features": (i32, [C.c_char_p, P(Config), P(i32), P(i32), P(i32)]),
        "rth_feature_string": (i32, [C.c_char_p, i32, C.c_char_p, u32]),
        "rth_render": (i32, [vp, C.c_char_p, u64, i32, vp, vp, P(Stats)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = the .so does not export what include/*.h declares
        fn.restype = res
        fn.argtypes = args


EXPORTED_SYMBOLS = None  # filled by load(); tests compare it with the prototypes in include/*.h


def load(path: os.PathLike | None = None) -> C.CDLL:
    """Load librt_b200.so (once). Raises if it has not been built: there is no other code path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise ImportError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(str(p))
    _declare(lib)
    if path is None:
        _lib = lib
    return lib


def check(code: int) -> None:
    if code != RT_OK:
        msg = load().rt_last_error()
        raise RtError(code, msg.decode() if msg else "")


def ptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
