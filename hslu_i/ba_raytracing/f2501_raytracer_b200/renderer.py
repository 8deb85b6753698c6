"""Renderer side of the host API: feature list -> rt_config, device scene handle, render calls.

Mirrors `RaytracerRenderer::<C>::default().render(&buffer, &scene)`
(reference src/renderer/raytracer_renderer.rs:1360-1379, called at src/main.rs:332) and
`ImageBuffer<W,H>` (src/image_buffer.rs:9-46).  Everything here ends in a call through the C ABI
of librt_b200.so; there is no Python, PyTorch or CPU implementation of the path.
"""
from __future__ import annotations

import ctypes as C
import itertools
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .scene import FlatScene, Scene

_EPOCHS = itertools.count(1)  # rt_config.tables_epoch values: one per freshly generated set of sampler tables


@dataclass
class RenderConfig:
    """rt_config plus the arrays it points to (kept alive here)."""

    cfg: L.Config
    aa_table: np.ndarray
    light_clouds: np.ndarray | None
    features: str = ""
    high_quality_model: bool = False
    scene_backface_culling: bool = False

    @staticmethod
    def from_features(features: str = "", seed: int = 0, n_light_clouds: int = 4096, width: int | None = None,
                      height: int | None = None, extra_flags: int = 0) -> "RenderConfig":
        """Resolve a Cargo feature list (e.g. "high_quality,simd_render", "default") like the
        reference's cfg!() switches do (Cargo.toml:62-83, raytracer_renderer.rs:55-93) and build
        the seeded sampler tables."""
        lib = L.load()
        cfg = L.Config()
        hq, sbc, rnd = C.c_int(), C.c_int(), C.c_int()
        L.check(lib.rth_config_from_features(features.encode(), C.byref(cfg), C.byref(hq), C.byref(sbc), C.byref(rnd)))
        if width:
            cfg.width = width
        if height:
            cfg.height = height
        cfg.flags |= extra_flags
        cfg.seed = seed
        cfg.tables_epoch = next(_EPOCHS)  # the tables below are never modified in place
        aa = np.zeros((cfg.aa_table_len, 2), np.float32)
        L.check(lib.rt_sampler_aa_table(seed, cfg.aa_samples_per_pixel, rnd.value, L.ptr(aa), cfg.aa_table_len))
        cfg.aa_table = aa.ctypes.data
        clouds = None
        if cfg.light_multiplicator > 1:
            clouds = np.zeros((n_light_clouds, cfg.light_multiplicator, 3), np.float32)
            L.check(lib.rt_sampler_light_clouds(seed, cfg.light_multiplicator, n_light_clouds, L.ptr(clouds)))
            cfg.light_clouds = clouds.ctypes.data
            cfg.n_light_clouds = n_light_clouds
        return RenderConfig(cfg, aa, clouds, features, bool(hq.value), bool(sbc.value))

    def with_tiles(self, tile_rows: int, rank: int, world: int) -> "RenderConfig":
        c = L.Config()
        C.memmove(C.byref(c), C.byref(self.cfg), C.sizeof(L.Config))
        c.tile_rows, c.tile_rank, c.tile_world = tile_rows, rank, world
        return RenderConfig(c, self.aa_table, self.light_clouds, self.features, self.high_quality_model, self.scene_backface_culling)

    def uncached(self) -> "RenderConfig":
        """Same config with tables_epoch = 0: the sampler tables are re-uploaded on every call."""
        c = L.Config()
        C.memmove(C.byref(c), C.byref(self.cfg), C.sizeof(L.Config))
        c.tables_epoch = 0
        return RenderConfig(c, self.aa_table, self.light_clouds, self.features, self.high_quality_model, self.scene_backface_culling)

    @property
    def samples_per_pixel(self) -> int:
        """Primary rays per pixel (get_total_rays, raytracer_renderer.rs:1018-1020)."""
        if not self.cfg.flags & L.FLAG_ANTI_ALIASING:
            return 1
        return self.cfg.aa_table_len if self.cfg.flags & L.FLAG_PACKET_SEMANTICS else self.cfg.aa_samples_per_pixel

    def local_rows(self) -> int:
        return L.load().rt_tile_local_rows(C.byref(self.cfg))


class ImageBuffer:
    """ImageBuffer<W,H>: row-major u32 framebuffer, zero-initialised (src/image_buffer.rs:25-29)."""

    def __init__(self, width: int, height: int):
        self.width, self.height = width, height
        self.buffer = np.zeros(width * height, np.uint32)
        self.hit_mask = np.zeros(width * height, np.uint8)

    def get_u32_slice(self) -> np.ndarray:
        return self.buffer

    def rgb(self) -> np.ndarray:
        b = self.buffer.reshape(self.height, self.width)
        return np.stack([(b >> 16) & 255, (b >> 8) & 255, b & 255], -1).astype(np.uint8)


class DeviceScene:
    """rt_scene: geometry + BVH resident on one GPU (rt_scene_create / rt_scene_destroy)."""

    def __init__(self, flat: FlatScene | Scene, device: int = 0):
        self._lib = L.load()
        self.flat = flat.flatten() if isinstance(flat, Scene) else flat
        self.device = device
        h = C.c_void_p()
        desc = self.flat.desc()
        L.check(self._lib.rt_scene_create(C.byref(desc), device, C.byref(h)))
        self._h = h

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.rt_scene_destroy(h)

    __del__ = close

    def render(self, rc: RenderConfig, buffer: ImageBuffer | None = None) -> tuple[ImageBuffer, dict]:
        """rt_render with HOST buffers (copies inside the call)."""
        buf = buffer or ImageBuffer(rc.cfg.width, rc.cfg.height)
        st = L.Stats()
        L.check(self._lib.rt_render(self._h, C.byref(rc.cfg), L.ptr(buf.buffer), L.ptr(buf.hit_mask), C.byref(st)))
        return buf, st.as_dict()

    def render_progressive(self, rc: RenderConfig, buffer: ImageBuffer | None = None, band_rows: int = 64, on_band=None) -> tuple[ImageBuffer, dict]:
        """rt_render_progressive: the frame fills in band by band; `on_band(first_row, n_rows)` is called after
        each band is in `buffer` (return True to stop), like the reference's window watching the render thread."""
        buf = buffer or ImageBuffer(rc.cfg.width, rc.cfg.height)
        st = L.Stats()
        cb = L.BAND_CALLBACK(lambda _user, first, n: 1 if (on_band and on_band(first, n)) else 0)
        L.check(self._lib.rt_render_progressive(self._h, C.byref(rc.cfg), L.ptr(buf.buffer), L.ptr(buf.hit_mask), band_rows, cb, None, C.byref(st)))
        d = st.as_dict()
        d["overlapped_bands"] = self._lib.rt_progressive_overlapped_bands(self._h)
        return buf, d

    def render_device(self, rc: RenderConfig, d_argb: int, d_mask: int | None, stream: int | None = None,
                      want_stats: bool = False) -> dict | None:
        """rt_render_device: outputs are device pointers (e.g. torch tensors' data_ptr())."""
        st = L.Stats()
        L.check(self._lib.rt_render_device(self._h, C.byref(rc.cfg), d_argb, d_mask, stream, C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def check(self) -> None:
        """rt_check_frame: waits for the frames enqueued with render_device() and raises if one of them is invalid."""
        L.check(self._lib.rt_check_frame(self._h))

    def dominant_kernel_name(self) -> str:
        """rt_last_kernel: the dominant kernel the last render call launched."""
        return (self._lib.rt_last_kernel(self._h) or b"").decode()

    def bvh_info(self) -> dict:
        """rt_scene_get_info: what rt_scene_create built and where its time went."""
        info = L.SceneInfo()
        L.check(self._lib.rt_scene_get_info(self._h, C.byref(info)))
        return {k: getattr(info, k) for k, _ in info._fields_}

    def kernel_times(self, n: int = 64) -> np.ndarray:
        """rt_kernel_times: device ms of the dominant kernel of the last (up to 64) render calls."""
        out = np.zeros(max(n, 1), np.float32)
        got = C.c_uint32(0)
        L.check(self._lib.rt_kernel_times(self._h, n, L.ptr(out), C.byref(got)))
        return out[: got.value]

    def trace_closest(self, rays: np.ndarray, flags: int = 0) -> tuple[np.ndarray, np.ndarray]:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        t = np.zeros(len(rays), np.float32)
        prim = np.zeros(len(rays), np.uint32)
        L.check(self._lib.rt_trace_closest(self._h, flags, L.ptr(rays), len(rays), L.ptr(t), L.ptr(prim)))
        return t, prim

    def trace_shadow(self, rays: np.ndarray, max_dist: np.ndarray, flags: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        md = np.ascontiguousarray(max_dist, np.float32)
        out = np.zeros((len(rays), 5), np.float32)
        L.check(self._lib.rt_trace_shadow(self._h, flags, L.ptr(rays), L.ptr(md), len(rays), L.ptr(out)))
        return out


class MultiDeviceScene:
    """rt_multi: the scene replicated on several GPUs of this process; render() splits the frame into interleaved
    row tiles and every device stores its rows into the first device's frame over NVLink peer memory."""

    def __init__(self, flat: FlatScene | Scene, devices: list[int] | None = None, n_devices: int = 0):
        self._lib = L.load()
        self.flat = flat.flatten() if isinstance(flat, Scene) else flat
        h = C.c_void_p()
        desc = self.flat.desc()
        arr = None
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            n_devices = len(devices)
        L.check(self._lib.rt_multi_create(C.byref(desc), arr, n_devices, C.byref(h)))
        self._h = h

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.rt_multi_destroy(h)

    __del__ = close

    @property
    def n_devices(self) -> int:
        return self._lib.rt_multi_device_count(self._h)

    def info(self) -> dict:
        info, peers = L.SceneInfo(), C.c_uint32(0)
        L.check(self._lib.rt_multi_get_info(self._h, C.byref(info), C.byref(peers)))
        d = {k: getattr(info, k) for k, _ in info._fields_}
        d["peer_devices"] = peers.value
        return d

    def render(self, rc: RenderConfig, buffer: ImageBuffer | None = None) -> tuple[ImageBuffer, dict]:
        buf = buffer or ImageBuffer(rc.cfg.width, rc.cfg.height)
        st = L.Stats()
        L.check(self._lib.rt_multi_render(self._h, C.byref(rc.cfg), L.ptr(buf.buffer), L.ptr(buf.hit_mask), C.byref(st)))
        return buf, st.as_dict()


class RaytracerRenderer:
    """RaytracerRenderer<C>: `render(buffer, scene)` for a Cargo feature list.  gpu_count > 1 (or 0 = every visible
    GPU) partitions the frame inside render(), like the reference's rayon pool does (src/image_buffer.rs:48-97)."""

    def __init__(self, features: str = "", seed: int = 0, device: int = 0, gpu_count: int = 1):
        self.features, self.seed, self.device, self.gpu_count = features, seed, device, gpu_count
        self.config = RenderConfig.from_features(features, seed)

    def render(self, buffer: ImageBuffer, scene: Scene | FlatScene) -> dict:
        if (buffer.width, buffer.height) != (self.config.cfg.width, self.config.cfg.height):
            raise ValueError("ImageBuffer size does not match the feature set's resolution")
        dev = DeviceScene(scene, self.device) if self.gpu_count == 1 else MultiDeviceScene(scene, n_devices=self.gpu_count)
        try:
            _, stats = dev.render(self.config, buffer)
        finally:
            dev.close()
        return stats
