"""Property tests of the shadow-beam criteria (hslu_i/.../csrc/rt_device.cuh: beam_axis / beam_child /
beam_filter), restated in numpy float32: whatever the criteria drop must be unreachable for EVERY ray of
the beam -- checked against the CPU oracle's shadow query (the reference's triangle test) on single-triangle
scenes and against an exact segment/box test.  Host logic only, no GPU: the device code is held to the
same property end to end by the bit-identical A/B tests in test_gpu_parity.py."""
import numpy as np
import pytest

F = np.float32
EPS = F(1.1920929e-7)


def cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], F)


def beam_filter_keeps(v1, v2, v3, P, lmin, lmax):
    """numpy float32 restatement of beam_filter's per-triangle decision."""
    v1, v2, v3, P, lmin, lmax = (np.asarray(x, F) for x in (v1, v2, v3, P, lmin, lmax))
    C = (lmin + lmax) * F(0.5)
    h = (lmax - lmin) * F(0.5) + F(1e-6) * np.abs(C).sum(dtype=F)
    Dc = C - P
    e1, e2 = v2 - v1, v3 - v1
    g = cross(e1, e2)
    q3w = F(64.0) * EPS * max(np.abs(e1).max(), np.abs(e2).max())
    b = v1 - P
    wu, wv = cross(e2, b), cross(b, e1)
    ww = (g - wu) - wv
    dot = lambda a, c: (a * c).sum(dtype=F)
    detc, uc, vc, wc = dot(Dc, g), dot(Dc, wu), dot(Dc, wv), dot(Dc, ww)
    hd, hu, hv, hw = dot(h, np.abs(g)), dot(h, np.abs(wu)), dot(h, np.abs(wv)), dot(h, np.abs(ww))
    dn = (np.abs(Dc) + h).sum(dtype=F)
    m = q3w * np.abs(b).sum(dtype=F) * dn
    tn = dot(b, g)
    mt = F(64.0) * EPS * dot(np.abs(b), np.abs(g))
    sg = F(np.copysign(1.0, detc))
    det_signed = abs(detc) > hd + m
    outside = (uc * sg < -(hu + m)) or (vc * sg < -(hv + m)) or (wc * sg < -(hw + F(3) * m)) or (tn * sg < -mt)
    return not (det_signed and outside)


def shadow_rays(P, targets, eps_d):
    """The M shadow rays of one shading point exactly as shade_pair builds them (float32)."""
    P = np.asarray(P, F)
    l2p = (targets - P).astype(F)
    ldir = (l2p / np.sqrt((l2p * l2p).sum(1, dtype=F))[:, None]).astype(F)
    q = (P + ldir * F(eps_d)).astype(F)
    maxd = np.sqrt(((targets - q) ** 2).sum(1, dtype=F)).astype(F)
    return np.concatenate([q, ldir], 1).astype(F), maxd


def test_beam_filter_never_drops_a_triangle_some_ray_hits(pkg, oracle_mod):
    rng = np.random.default_rng(2024)
    eps_d = 1.1920929e-7 * 100 * (1 + 0.8333333 + 0.9166667) / 3
    mat = pkg.material([0.5, 0.5, 0.5])
    kept = dropped = hit_cases = 0
    for case in range(500):
        scale = 10 ** rng.uniform(-2.6, -0.2)
        v1 = rng.uniform(0.1, 0.9, 3)
        v2, v3 = v1 + rng.normal(0, scale, 3), v1 + rng.normal(0, scale, 3)
        n = np.cross(v2 - v1, v3 - v1)
        if np.linalg.norm(n) < 1e-12:
            continue
        # aim the beam at a point of the triangle's plane in or near the triangle, from a point off the plane
        u, v = rng.uniform(-0.4, 1.4, 2)
        X = v1 + u * (v2 - v1) + v * (v3 - v1)
        P = X + rng.normal(0, 1, 3) * 10 ** rng.uniform(-2, 0)
        C = P + (X - P) * rng.uniform(1.05, 3.0)
        half = 10 ** rng.uniform(-3.3, -1.8, 3)
        lmin, lmax = (C - half).astype(F), (C + half).astype(F)
        targets = (lmin + rng.random((96, 3)).astype(F) * (lmax - lmin)).astype(F)
        targets[:8] = [[(lmin, lmax)[(k >> a) & 1][a] for a in range(3)] for k in range(8)]  # the corners too
        s = pkg.Scene()
        s.add_triangle(v1.astype(F), v2.astype(F), v3.astype(F), mat)
        flat = s.flatten()
        rays, maxd = shadow_rays(P.astype(F), targets, eps_d)
        occluded = oracle_mod.trace_shadow(flat.desc(), rays, maxd, 0, oracle_mod.BRUTE)[:, 0] == 1
        tri = flat.triangles[0]
        keep = beam_filter_keeps(tri["v1"], tri["v2"], tri["v3"], P.astype(F), lmin, lmax)
        if occluded.any():
            hit_cases += 1
            assert keep, f"case {case}: {int(occluded.sum())} rays of the beam hit the triangle but the filter drops it"
        kept += keep
        dropped += not keep
    assert hit_cases > 100 and dropped > 100  # the test exercised both outcomes


def beam_touches_box(P, lmin, lmax, pad, lo, hi):
    """numpy float32 restatement of make_beam + beam_child for one box."""
    P, lmin, lmax, lo, hi = (np.asarray(x, F) for x in (P, lmin, lmax, lo, hi))
    C = F(0.5) * (lmin + lmax)
    h = F(0.5) * (lmax - lmin) + F(pad)
    tn, tf, extra_done = F(0.0), F(1.001), False
    for a in range(3):
        D = C[a] - P[a]
        a1, a2 = D - h[a], D + h[a]
        a1 = F(-1e-20) if abs(a1) < 1e-20 else a1
        a2 = F(1e-20) if abs(a2) < 1e-20 else a2
        i1, i2 = F(1.0) / a1, F(1.0) / a2
        o1, o2 = (P[a] - F(pad)) * i1, (P[a] + F(pad)) * i2
        if a1 > 0:
            tn, tf = max(tn, lo[a] * i2 - o2), min(tf, hi[a] * i1 - o1)
        elif a2 < 0:
            tn, tf = max(tn, hi[a] * i1 - o1), min(tf, lo[a] * i2 - o2)
        else:  # the hull widens both ways: two near bounds (the second only for the first such axis), no far bound
            tn = max(tn, hi[a] * i1 - o1)
            if not extra_done:
                tn = max(tn, lo[a] * i2 - o2)
                extra_done = True
    return tn <= tf


def segment_hits_box(P, L, lo, hi):
    d = L - P
    t0, t1 = 0.0, 1.0
    for a in range(3):
        if abs(d[a]) < 1e-300:
            if P[a] < lo[a] or P[a] > hi[a]:
                return False
            continue
        ta, tb = (lo[a] - P[a]) / d[a], (hi[a] - P[a]) / d[a]
        t0, t1 = max(t0, min(ta, tb)), min(t1, max(ta, tb))
    return t0 <= t1


def test_beam_hull_touches_every_box_a_segment_crosses():
    rng = np.random.default_rng(7)
    pad = 0.25 * 1e-4 + 2 * 1.1e-5
    touched = missed = crossed = 0
    for case in range(3000):
        P = rng.uniform(0, 1, 3)
        C = rng.uniform(0, 1, 3)
        if case % 3 == 0:  # light nearly aligned with the shading point on one axis: the hull widens both ways there
            a = rng.integers(3)
            C[a] = P[a] + rng.normal(0, 2e-3)
        half = 10 ** rng.uniform(-3.3, -1.8, 3)
        lmin, lmax = (C - half).astype(F), (C + half).astype(F)
        centre = P + (C - P) * rng.uniform(0, 1) + rng.normal(0, 1, 3) * 10 ** rng.uniform(-3, -1)
        ext = 10 ** rng.uniform(-3, -0.7, 3)
        lo, hi = (centre - ext).astype(F), (centre + ext).astype(F)
        touch = beam_touches_box(P.astype(F), lmin, lmax, pad, lo, hi)
        Ls = lmin.astype(np.float64) + rng.random((64, 3)) * (lmax.astype(np.float64) - lmin.astype(np.float64))
        cross_any = any(segment_hits_box(P.astype(F).astype(np.float64), L, lo.astype(np.float64), hi.astype(np.float64)) for L in Ls)
        if cross_any:
            crossed += 1
            assert touch, f"case {case}: a segment of the beam crosses the box but the hull test skips it"
        touched += bool(touch)
        missed += not touch
    assert crossed > 500 and missed > 500
