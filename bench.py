#!/usr/bin/env python3
"""Benchmark of the render path: rays/s and ms/frame of one frame of a reference scene.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--scenes c1,c2,c4,c5|none] [--impl ours|reference]

A "step" is one frame of the workload: primary rays, BVH traversal, shadow/transmittance rays,
bounces, AA accumulation and the ARGB framebuffer write.  Default workload = BASELINE.json's
target config: semesterbild `high_quality` + `simd_render` (768x640, 16 AA samples, 5 lights x
19 soft-shadow copies).  With N > 1 (torchrun, one rank per GPU) the frame is split into
interleaved row tiles and gathered on rank 0 over NCCL; the frame (total work) is fixed, so the
scaling is "strong".  Rays per frame come from the counter build of the same kernels (untimed).

The metric is quoted "per scene": after the headline workload the same run measures the other BASELINE
configs (`--scenes`, default c1,c2,c4,c5) with fewer steps and reports them in the `scenes` table of the
same JSON line (ms/frame, rays/s, dominant kernel, roofline fraction, scene-create time).

`--impl reference` times the reference's CPU algorithm (the brute-force C++ oracle, all host
threads; the Rust crate cannot be built in this image) on a bounded pixel sample of the same
workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene, cargo features, description)
    "c3": ("semesterbild", "high_quality,simd_render", "semesterbild high_quality+simd_render 768x640 S=16 M=19 (BASELINE configs[2])"),
    "c3_scalar": ("semesterbild", "high_quality", "semesterbild high_quality scalar path 768x640 S=9 M=19"),
    "c1": ("main", "default,simd_render", "cargo run --release defaults + simd_render 1140x950 S=16 M=19 depth 13/18 (BASELINE configs[0])"),
    "c1lite": ("main", "", "main.rs --no-default-features 768x640"),
    "c2": ("test_scene", "simd_render,backface_culling,medium_resolution", "examples/test_scene 1140x950 (BASELINE configs[1])"),
    "c4": ("semesterbild", "realistic,high_resolution", "semesterbild realistic high_resolution 1620x1350 depth 9/8 (BASELINE configs[3])"),
    "c5": ("large_mesh", "extreme_quality,realistic,high_resolution,simd_render",
           "semesterbild objects + synthetic 4 194 304-triangle relief, extreme_quality+realistic 1620x1350 S=24 M=28 depth 21/21 (BASELINE configs[4])"),
}
# (steps, warmup) of the per-scene table entries: bounded so that the default run stays within minutes
SCENE_STEPS = {"c1": (3, 1), "c2": (10, 3), "c4": (10, 3), "c5": (2, 1), "c3": (10, 3), "c3_scalar": (10, 3), "c1lite": (10, 3)}
# canonical algorithmic costs
F_BOX, F_TRI, F_SPH = 22, 45, 22
B_BOX, B_TRI, B_SPH = 32, 52, 20
# beam-vs-triangle culling test (rt_device.cuh beam_filter): 3 sub, 2 cross (18), w = g - wu - wv (6), 4 dot (20),
# 4 abs-dot (20), margins (10); once per (shading point, light, candidate triangle), not per ray
F_BEAM_TRI = 77


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name: str, seed: int):
    from hslu_i.ba_raytracing.f2501_raytracer_b200 import RenderConfig, example_scene

    scene_name, features, desc = WORKLOADS[name]
    if scene_name == "large_mesh":
        from hslu_i.ba_raytracing.f2501_raytracer_b200.synthetic import large_mesh_scene

        flat = large_mesh_scene(features).flatten()
    else:
        flat = example_scene(scene_name, features).flatten()
    rc = RenderConfig.from_features(features, seed)
    return flat, rc, features, desc


def workload_config(flat, rc, features: str, desc: str, seed: int) -> dict:
    """Identity of the workload: the same dict in our arm and in the reference arm (no arm-specific keys)."""
    return {"workload": desc, "features": features, "width": rc.cfg.width, "height": rc.cfg.height,
            "primitives": flat.n_primitives, "lights": len(flat.lights), "samples_per_pixel": rc.samples_per_pixel,
            "light_copies": rc.cfg.light_multiplicator, "reflection_depth": rc.cfg.reflection_max_depth,
            "refraction_depth": rc.cfg.refraction_max_depth, "seed": seed}


def oracle_sample(flat, rc, n_pixels: int, threads: int, seed: int = 1234):
    """Brute-force oracle (the reference's linear scan, 8-ray AVX2 packets = its `simd_render` path
    where the config is a packet config, scalar otherwise) on a random pixel sample; returns rays, seconds, pixels."""
    from oracle import oracle

    W, H = rc.cfg.width, rc.cfg.height
    rng = np.random.default_rng(seed)
    pix = rng.choice(W * H, size=min(n_pixels, W * H), replace=False).astype(np.uint32)
    desc = flat.desc()
    t0 = time.perf_counter()
    _, _, _, cnt = oracle.render(desc, rc.cfg, oracle.PACKET8, threads, pix)
    dt = time.perf_counter() - t0
    oracle_sample.last_counters = cnt  # (cpu_baseline reports the share of shadow queries whose surface faces away from the light)
    return cnt["closest_rays"] + cnt["shadow_rays"], dt, len(pix)


def size_sample(flat, rc, cores: int, budget_s: float) -> int:
    """Pixels per sample for ~budget_s of CPU work: a short probe (one pixel per thread -- the oracle hands out
    chunks of n/(8 threads) pixels, so every thread is busy even then), a second probe ten times larger when
    the first one was over in a blink, then scaled to the budget.  Never below 4 pixels per thread."""
    frame = rc.cfg.width * rc.cfg.height
    n = min(frame, 2 * cores)
    _, dt, n = oracle_sample(flat, rc, n, cores, seed=7)
    if dt < 0.5 and n < frame:
        n = min(frame, 16 * n)
        _, dt, n = oracle_sample(flat, rc, n, cores, seed=8)
    per_pixel = max(dt / max(n, 1), 1e-7)
    return int(max(min(4 * cores, frame), min(frame, budget_s / per_pixel)))


def cpu_baseline(flat, rc, budget_s: float = 15.0) -> dict:
    cores = os.cpu_count() or 1
    n = size_sample(flat, rc, cores, budget_s)
    rays, dt, n = oracle_sample(flat, rc, n, cores)
    cnt = oracle_sample.last_counters
    return {"value": rays / dt, "unit": "rays/s", "cores": cores, "kind": "port",
            "shadow_queries_facing_away": {
                "fraction": cnt["shadow_rays_facing_away"] / max(cnt["shadow_rays"], 1),
                "note": "share of the sample's shadow queries whose hit faces away from the light (`light_valid` false, raytracer_renderer.rs:849): the "
                        "reference casts them and drops the result, the CUDA path decides `diffuse > 0` first and traces nothing; both arms count them as rays"},
            "sample": f"{n} random pixels of the workload ({rays} rays, {dt:.1f} s), brute-force C++ oracle, "
                      f"8-ray AVX2 packets (the reference's simd_render instantiation: linear scan over all {flat.n_primitives} "
                      f"primitives per packet), {cores} threads, pixel queue in chunks of {max(1, min(64, n // (8 * cores)))}"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    flat, rc, features, desc = build_workload(args.workload, args.seed)
    cores = os.cpu_count() or 1
    # one step = a bounded pixel sample: ~ (150 s / (K + W)) of CPU work on all host threads
    step_budget = max(2.0, 150.0 / max(args.steps + args.warmup, 1))
    if os.environ.get("RT_BENCH_REF_STEP_S"):  # tests: a much smaller sample per step
        step_budget = float(os.environ["RT_BENCH_REF_STEP_S"])
    n_step = size_sample(flat, rc, cores, step_budget)
    for i in range(args.warmup):
        oracle_sample(flat, rc, n_step, cores, seed=100 + i)
    tot_rays, tot_t = 0, 0.0
    for i in range(args.steps):
        rays, dt, n = oracle_sample(flat, rc, n_step, cores, seed=200 + i)
        tot_rays += rays
        tot_t += dt
    value = tot_rays / tot_t
    frame_pixels = rc.cfg.width * rc.cfg.height
    chunk = max(1, min(64, n_step // (8 * cores)))
    threads_busy = min(cores, -(-n_step // chunk))
    line = {
        "impl": "reference", "metric": "rays_per_second", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(flat, rc, features, desc, args.seed),
        "run": {"step": f"{n_step} random pixels of the frame per step ({n_step / frame_pixels:.5f} of a frame)", "threads": threads_busy,
                "pixel_queue_chunk": chunk,
                "native_note": "librt_b200.so is loaded by this arm for HOST code only (scene construction / flattening, seeded sampler tables); "
                               "no kernel is launched, the timed work is oracle/liboracle_rt.so"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads_busy, "kind": "port",
                         "sample": f"{n_step} random pixels per step, brute-force C++ oracle with 8-ray AVX2 packets (the reference's simd_render linear scan), "
                                   f"{threads_busy} threads busy of {cores} host threads; "
                                   "the Rust reference itself cannot be built in this image (no cargo/rustc, crates not vendored)"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_frame_extrapolated": 1e3 * (tot_t / max(args.steps, 1)) * frame_pixels / n_step,
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout() -> None:
    """Everything chatty (NCCL's version banner, library warnings) goes to stderr; stdout carries the one
    JSON line only."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


class Ctx:
    """Process-wide state of our arm: torch, the process group, helpers."""

    def __init__(self, args):
        import torch

        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: librt_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        self.fp32_peak = None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x: float, op) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def allmax(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist is not None else x

    def allsum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist is not None else x


def measure(ctx: Ctx, name: str, steps: int, warmup: int, headline: bool) -> dict:
    """One workload on this process group: counter passes (untimed), K timed device-resident steps, the
    end-to-end leg through the host API.  Every rank runs it; rank 0 uses the result."""
    import ctypes as C

    from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, ImageBuffer, RenderConfig
    from hslu_i.ba_raytracing.f2501_raytracer_b200 import _lib as L
    from hslu_i.ba_raytracing.f2501_raytracer_b200.distributed import TiledRenderer

    torch, args, rank, world = ctx.torch, ctx.args, ctx.rank, ctx.world
    lib = L.load()
    t0 = time.perf_counter()
    flat, rc, features, desc = build_workload(name, args.seed)
    host_scene_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    dev = DeviceScene(flat, device=ctx.local_rank)  # rt_scene_create: BVH build + upload
    torch.cuda.synchronize()
    scene_create_ms = ctx.allmax(1e3 * (time.perf_counter() - t0))
    bvh_info = dev.bvh_info()
    host_builder = None
    if bvh_info["builder"] == 1 and not headline:
        # the scene was built on the GPU: for comparison, the host builder's time and the frame time on ITS tree
        os.environ["RT_BVH_BUILDER"] = "host"
        try:
            t0 = time.perf_counter()
            dev_h = DeviceScene(flat, device=ctx.local_rank)
            torch.cuda.synchronize()
            create_h = ctx.allmax(1e3 * (time.perf_counter() - t0))
            th = TiledRenderer(dev_h, rc, rank, world, args.tile_rows)
            th.step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.barrier()
            e0.record()
            th.step()
            e1.record()
            torch.cuda.synchronize()
            th.check()
            host_builder = {"scene_create_ms": create_h, "ms_per_frame": ctx.allmax(e0.elapsed_time(e1)), "bvh": dev_h.bvh_info()}
            dev_h.close()
            del th
        finally:
            del os.environ["RT_BVH_BUILDER"]
    tiled = TiledRenderer(dev, rc, rank, world, args.tile_rows)

    # ---- untimed: work counters of this rank's tiles (counter build of the same kernels)
    def count_pass(extra_flags: int) -> dict:
        rcc = RenderConfig(L.Config(), rc.aa_table, rc.light_clouds)
        C.memmove(C.byref(rcc.cfg), C.byref(tiled.rc.cfg), C.sizeof(L.Config))
        rcc.cfg.flags |= L.FLAG_COUNT_WORK | extra_flags
        stc = dev.render_device(rcc, tiled.slab.data_ptr(), tiled.mask.data_ptr(), None, want_stats=True)
        out = {k: ctx.allsum(float(stc[k])) for k in ("closest_rays", "shadow_rays", "box_tests", "triangle_tests", "sphere_tests", "beam_triangle_tests",
                                                       "prepass_box_tests", "prepass_triangle_tests", "prepass_sphere_tests")}
        out["kernel_launches"] = int(stc["kernel_launches"])
        return out

    counts = count_pass(0)  # what the timed kernels execute (shadow beams)
    rays_per_frame = counts["closest_rays"] + counts["shadow_rays"]

    def alg(c):  # tests of the dominant kernel alone: totals minus the primary pre-pass kernel's share
        nb, nt, ns = c["box_tests"] - c["prepass_box_tests"], c["triangle_tests"] - c["prepass_triangle_tests"], c["sphere_tests"] - c["prepass_sphere_tests"]
        fl = nb * F_BOX + nt * F_TRI + ns * F_SPH + c["beam_triangle_tests"] * F_BEAM_TRI
        by = nb * B_BOX + (nt + c["beam_triangle_tests"]) * B_TRI + ns * B_SPH
        return fl, by

    flops_alg, bytes_alg = alg(counts)
    counts_ray = flops_ray = None
    if headline:  # the same frame with one BVH walk per shadow ray: the canonical per-ray test counts
        counts_ray = count_pass(L.FLAG_PER_RAY_SHADOWS)
        assert rays_per_frame == counts_ray["closest_rays"] + counts_ray["shadow_rays"]
        flops_ray, _ = alg(counts_ray)
    if ctx.fp32_peak is None:
        peak = C.c_double(0.0)
        L.check(lib.rt_measure_fp32_peak(ctx.local_rank, C.byref(peak)))
        ctx.fp32_peak = peak.value

    # ---- warm-up
    for _ in range(warmup):
        tiled.step()
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps; per-step CUDA events on the launching stream, L2 flushed between steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    ctx.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for i in range(steps):
        ctx.flush.zero_()
        ev[i][0].record()
        tiled.render_local()
        ev[i][1].record()
        tiled.gather()
        ev[i][2].record()
    torch.cuda.synchronize()
    ctx.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    tiled.check()  # the frames' device watchdog words (a wait loop that gave up = invalid frame)
    dom_ms = dev.kernel_times(min(steps, 64))  # dominant kernel alone, one CUDA-event pair per timed step
    dom_ms_max = ctx.allmax(float(dom_ms.mean()) if len(dom_ms) else 0.0)
    step_ms = ctx.allmax(sum(e[0].elapsed_time(e[2]) for e in ev))
    kern_ms_max = ctx.allmax(sum(e[0].elapsed_time(e[1]) for e in ev))
    ms_per_step = step_ms / steps

    # ---- end to end through the public host API: host buffers, copies inside the timed region
    h2d = (rc.light_clouds.nbytes if rc.light_clouds is not None else 0) + 4

    def pinned_tables(cfgd):
        """The step's inputs (the sampler tables) in PINNED host memory; tables_epoch = 0: copied host-to-device on every call."""
        out = cfgd.uncached()
        keep = []
        for name in ("aa_table", "light_clouds"):
            arr = getattr(out, name)
            if arr is not None:
                t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
                keep.append(t)
                view = t.numpy()
                setattr(out, name, view)
                setattr(out.cfg, name, view.ctypes.data)
        out._pinned = keep
        return out

    if world == 1:
        buf = ImageBuffer(rc.cfg.width, rc.cfg.height)
        rc_e2e = pinned_tables(rc)
        dev.render(rc_e2e, buf)
        t0 = time.perf_counter()
        for _ in range(steps):
            dev.render(rc_e2e, buf)  # H2D sampler tables, kernels, D2H framebuffer + mask, host merge
        e2e_s = time.perf_counter() - t0
        d2h = rc.cfg.width * rc.cfg.height * 4  # the frame; the hit mask is derived from it (pixel != 0)
    else:
        host = torch.empty(rc.cfg.width * rc.cfg.height, dtype=torch.int32).pin_memory() if rank == 0 else None
        host_m = torch.empty(rc.cfg.width * rc.cfg.height, dtype=torch.uint8).pin_memory() if rank == 0 else None
        tiled.rc = pinned_tables(tiled.rc)
        tiled.step()
        ctx.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            tiled.step()  # (the gather keeps the ranks in step: rank 0 cannot finish a frame before every rank has sent its tiles)
            if rank == 0:
                host.copy_(tiled.frame, non_blocking=True)
                host_m.copy_(tiled.frame_mask, non_blocking=True)
            torch.cuda.synchronize()
        ctx.barrier()
        e2e_s = ctx.allmax(time.perf_counter() - t0)
        d2h = rc.cfg.width * rc.cfg.height * 5
        tiled.check()
    e2e_ms = 1e3 * e2e_s / steps

    bounces = bool(rc.cfg.flags & (L.FLAG_REFLECTIONS | L.FLAG_REFRACTIONS))
    dominant_kernel = dev.dominant_kernel_name()
    res = dict(name=name, flat=flat, rc=rc, features=features, desc=desc, counts=counts, counts_ray=counts_ray, rays_per_frame=rays_per_frame,
               flops_alg=flops_alg, bytes_alg=bytes_alg, flops_ray=flops_ray, kernel_ms=dom_ms_max, render_ms=kern_ms_max / steps,
               ms_per_step=ms_per_step, e2e_ms=e2e_ms, h2d=h2d, d2h=d2h, clocks=clocks, wall=wall, scene_create_ms=scene_create_ms,
               host_scene_s=host_scene_s, dominant_kernel=dominant_kernel, steps=steps, warmup=warmup,
               launches_per_step=int(counts["kernel_launches"]) + (0 if world == 1 else 1), bvh=bvh_info, host_builder=host_builder)
    dev.close()
    del tiled
    torch.cuda.empty_cache()
    return res


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scenes", default=os.environ.get("RT_BENCH_SCENES", "c1,c2,c4,c5"),
                    help="other BASELINE configs measured after the headline workload (comma list, or 'none')")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--tile-rows", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return

    ctx = Ctx(args)
    rank, world = ctx.rank, ctx.world
    m = measure(ctx, args.workload, args.steps, args.warmup, headline=True)
    scenes = {}
    for name in [s for s in args.scenes.split(",") if s and s != "none" and s != args.workload]:
        if name not in WORKLOADS:
            raise SystemExit(f"unknown scene {name}")
        k, w = SCENE_STEPS.get(name, (5, 2))
        s = measure(ctx, name, k, w, headline=False)
        ach = s["flops_alg"] / world / (s["kernel_ms"] * 1e-3) / 1e12 if s["kernel_ms"] else None
        scenes[name] = {
            "workload": s["desc"], "ms_per_frame": s["ms_per_step"], "rays_per_s": s["rays_per_frame"] / (s["ms_per_step"] * 1e-3),
            "rays_per_frame": int(s["rays_per_frame"]), "e2e_ms_per_frame": s["e2e_ms"], "e2e_rays_per_s": s["rays_per_frame"] / (s["e2e_ms"] * 1e-3),
            "steps": k, "warmup": w, "dominant_kernel": s["dominant_kernel"], "kernel_ms": s["kernel_ms"],
            "roofline_frac_fp32": ach / ctx.fp32_peak if ach and ctx.fp32_peak else None, "achieved_tflops": ach,
            "algorithmic_gbs": s["bytes_alg"] / world / (s["kernel_ms"] * 1e-3) / 1e9 if s["kernel_ms"] else None,
            "scene_create_ms": s["scene_create_ms"], "primitives": s["flat"].n_primitives, "bvh": s["bvh"],
            "host_builder_for_comparison": s["host_builder"],
            "tests_per_ray": {k2: s["counts"][k2] / s["rays_per_frame"] for k2 in ("box_tests", "triangle_tests", "sphere_tests", "beam_triangle_tests")},
        }

    if rank == 0:
        flat, rc, counts, counts_ray = m["flat"], m["rc"], m["counts"], m["counts_ray"]
        rays_per_frame, kernel_ms = m["rays_per_frame"], m["kernel_ms"]
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        fp32_peak = ctx.fp32_peak
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this workload
            tj = json.loads((ROOT / "profiles" / "roofline_traffic.json").read_text()).get(args.workload)
            if tj and tj.get("kernel", "").split("<")[0] == m["dominant_kernel"].split("<")[0] and world == 1:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except (OSError, ValueError):
            pass
        ach_tflops = m["flops_alg"] / world / (kernel_ms * 1e-3) / 1e12  # per GPU, dominant kernel
        line = {
            "metric": "rays_per_second", "value": rays_per_frame / (m["ms_per_step"] * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(flat, rc, m["features"], m["desc"], args.seed),
            "run": {"parallelism": f"row tiles of {args.tile_rows} rows interleaved over {world} GPU(s), NCCL gather to rank 0",
                    "l2": "flushed between timed steps (256 MiB write)", "rays_per_frame": int(rays_per_frame),
                    "closest_rays": int(counts["closest_rays"]), "shadow_rays": int(counts["shadow_rays"])},
            "ms_per_frame": m["ms_per_step"],
            "e2e": {"value": rays_per_frame / (m["e2e_ms"] * 1e-3), "unit": "rays/s", "ms_per_frame": m["e2e_ms"],
                    "h2d_bytes_per_step": int(m["h2d"]), "d2h_bytes_per_step": int(m["d2h"]),
                    "scene_create_ms": m["scene_create_ms"],
                    "ms_per_frame_with_scene_create": m["e2e_ms"] + m["scene_create_ms"],
                    "note": "the Rust shim's render() flattens the scene and calls rt_scene_create on every call "
                            "(reference times exactly that call, src/main.rs:331-334): scene_create_ms = host BVH build + upload, once per scene"},
            "scene_create_ms": m["scene_create_ms"],
            "gpu_launches": int(args.steps * m["launches_per_step"]),
            "executed_tests_per_ray": {k: counts[k] / rays_per_frame for k in ("box_tests", "triangle_tests", "sphere_tests", "beam_triangle_tests")},
            "clocks": m["clocks"],
            "roofline": {"bound": "fp32", "achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": ach_tflops / fp32_peak if fp32_peak else None, "traffic": traffic,
                         "kernel": m["dominant_kernel"], "kernel_ms": kernel_ms, "all_render_kernels_ms": m["render_ms"],
                         "peak_source": "measured in this run: rt_measure_fp32_peak (register FFMA microbenchmark); MEASURED_PEAKS.json has no FP32 figure",
                         "algorithmic": f"EXECUTED tests x canonical flop (SURVEY 8d): {F_BOX}/box, {F_TRI}/triangle, {F_SPH}/sphere, {F_BEAM_TRI}/beam-triangle test; "
                                        f"{int(counts['box_tests'])} box, {int(counts['triangle_tests'])} triangle, {int(counts['sphere_tests'])} sphere, "
                                        f"{int(counts['beam_triangle_tests'])} beam-triangle tests per frame, of which the primary pre-pass kernel's "
                                        f"{int(counts['prepass_box_tests'])} / {int(counts['prepass_triangle_tests'])} / {int(counts['prepass_sphere_tests'])} are left out",
                         "per_ray_traversal_equivalent": {
                             "achieved": m["flops_ray"] / world / (kernel_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                             "frac": m["flops_ray"] / world / (kernel_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None,
                             "note": "the same frame's test counts when every shadow ray walks the BVH itself (RT_FLAG_PER_RAY_SHADOWS counter build: "
                                     f"{int(counts_ray['box_tests'])} box, {int(counts_ray['triangle_tests'])} triangle, {int(counts_ray['sphere_tests'])} sphere) "
                                     "over the measured kernel time: the work shadow beams avoid counts as done here, it does not in `achieved`"},
                         "traffic_source": traffic_src,
                         "hbm": {"achieved": m["bytes_alg"] / world / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": m["bytes_alg"] / world / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                                 "note": ("algorithmic node/primitive bytes; the scene is L1/L2 resident, so this is cache not DRAM traffic"
                                          if flat.n_primitives * 64 < 64e6 else
                                          "algorithmic node/primitive bytes; the primitive array exceeds L2, part of this is DRAM traffic")}},
            "scenes": scenes,
            "wall_s_timed_region": m["wall"],
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(flat, rc)
        emit(line)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
