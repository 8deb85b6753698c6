#!/usr/bin/env python3
"""Quick GPU-vs-oracle check used during bring-up (run under gpurun). Prints one line per config."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, RenderConfig, example_scene
from hslu_i.ba_raytracing.f2501_raytracer_b200 import _lib as L
from oracle import oracle


def compare(name, scene_name, features, stride=1, seed=0, threads=0):
    scene = example_scene(scene_name, features)
    flat = scene.flatten()
    rc = RenderConfig.from_features(features, seed)
    dev = DeviceScene(flat)
    buf, st = dev.render(rc)
    t0 = time.time()
    buf, st = dev.render(rc)
    t_gpu = time.time() - t0
    rcc = RenderConfig.from_features(features, seed, extra_flags=L.FLAG_COUNT_WORK)
    _, stc = dev.render(rcc)
    W, H = rc.cfg.width, rc.cfg.height
    ys, xs = np.mgrid[0:H:stride, 0:W:stride]
    pix = (ys * W + xs).astype(np.uint32).ravel()
    desc = flat.desc()
    t0 = time.time()
    argb, mask, _, cnt = oracle.render(desc, rc.cfg, oracle.BVH, threads, None if stride == 1 else pix)
    t_cpu = time.time() - t0
    g_argb, g_mask = buf.buffer[pix], buf.hit_mask[pix]
    o_argb, o_mask = argb[pix], mask[pix]
    mask_diff = int((g_mask != o_mask).sum())
    ch = lambda a, s: ((a >> s) & 255).astype(np.int32)
    d = np.max([np.abs(ch(g_argb, s) - ch(o_argb, s)) for s in (16, 8, 0)], axis=0)
    both = (g_mask == 1) & (o_mask == 1)
    print(f"{name}: {W}x{H} prims={flat.n_primitives} S={rc.samples_per_pixel} M={rc.cfg.light_multiplicator} "
          f"gpu_kernel_ms={st['kernel_ms']:.3f} gpu_total_ms={t_gpu*1e3:.2f} oracle_s={t_cpu:.2f} compared={len(pix)} "
          f"mask_diff={mask_diff} exact={(d[both]==0).mean():.5f} within1={(d[both]<=1).mean():.5f} maxdiff={int(d.max())} "
          f"hitfrac={g_mask.mean():.4f} gpu_rays=({stc['closest_rays']},{stc['shadow_rays']}) "
          f"oracle_rays=({cnt['closest_rays']},{cnt['shadow_rays']}) box={stc['box_tests']} tri={stc['triangle_tests']} sph={stc['sphere_tests']}",
          flush=True)
    if mask_diff or int(d.max()) > 1:
        bad = np.nonzero((g_mask != o_mask) | (d > 1))[0][:10]
        for b in bad:
            print("   bad pixel", int(pix[b] % W), int(pix[b] // W), hex(int(g_argb[b])), hex(int(o_argb[b])), int(g_mask[b]), int(o_mask[b]))
    dev.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c1lite", "c4", "c3"]
    if "c2" in which:
        compare("C2 test_scene", "test_scene", "simd_render,backface_culling,medium_resolution")
    if "c1lite" in which:
        compare("C1-lite", "main", "")
    if "c4" in which:
        compare("C4 realistic hires", "main", "realistic,high_resolution", stride=2)
    if "c3" in which:
        compare("C3 high_quality simd", "main", "high_quality,simd_render", stride=4)
    if "c3s" in which:
        compare("C3 high_quality scalar", "main", "high_quality", stride=4)
    if "c1" in which:
        compare("C1 default simd", "main", "default,simd_render", stride=8)
