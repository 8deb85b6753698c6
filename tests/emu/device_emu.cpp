// device_emu.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the product's device-side traversal / shadow-beam code
// (hslu_i/.../csrc/rt_device.cuh, unchanged) for the HOST so that the "not gpu" test suite can run it against the
// CPU oracle: BVH4 traversal == linear scan, beam candidate lists == per-ray traversal, on the real flattened
// scenes (rt_bvh.cpp builds them).  Nothing in the product links or loads this; a renderer on the CPU it is not
// (no kernels, no shading, no frame loop): it answers raw queries for the tests.
#define RT_HOST_EMULATION 1
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

#include <cuda_runtime.h>  // float3 / float4 / make_float3 ... (plain structs for a host compiler)

// ---- what the device code expects from CUDA, for a host compiler
using std::signbit;
static inline float __ldg(const float *p) { return *p; }
static inline float4 __ldg(const float4 *p) { return *p; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __fmul_rn(float a, float b) { return a * b; }  // (-ffp-contract=off: never fused)
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

#include "rt_device.cuh"

using namespace rtd;

struct EmuScene {
    rtb::FlatScene flat;
    DScene d;
};

extern "C" {

void *emu_scene_create(const rt_scene_desc *desc) {
    EmuScene *s = new EmuScene();
    if (rtb::build_flat_scene(desc, s->flat) != RT_OK) {
        delete s;
        return nullptr;
    }
    s->flat.nodes.resize(s->flat.nodes.size() + 32);  // slack: the device arrays are never read past their end either
    s->flat.prims.resize(std::max<size_t>(s->flat.prims.size(), 16));
    s->flat.shade_n.resize(std::max<size_t>(s->flat.shade_n.size(), 4));
    s->flat.mats.resize(std::max<size_t>(s->flat.mats.size(), 12));
    s->flat.lights.resize(std::max<size_t>(s->flat.lights.size(), 8));
    s->d.nodes = (const float4 *)s->flat.nodes.data();
    s->d.prims = (const float4 *)s->flat.prims.data();
    s->d.shade_n = (const float4 *)s->flat.shade_n.data();
    s->d.mats = (const float4 *)s->flat.mats.data();
    s->d.lights = (const float4 *)s->flat.lights.data();
    s->d.n_prims = s->flat.n_prims;
    s->d.n_nodes = s->flat.n_nodes;
    s->d.n_lights = s->flat.n_lights;
    s->d.nan_hit_slot = s->flat.nan_hit_slot;
    s->d.nan_hit_slot_culled = s->flat.nan_hit_slot_culled;
    s->d.box_pad = s->flat.box_pad;
    return s;
}
void emu_scene_destroy(void *h) { delete (EmuScene *)h; }
float emu_box_pad(void *h) { return ((EmuScene *)h)->d.box_pad; }

// the bodies of trace_closest_kernel / trace_shadow_kernel (rt_kernels.cuh)
void emu_trace_closest(void *h, uint32_t flags, const float *rays, uint32_t n, float *out_t, uint32_t *out_prim) {
    const DScene &sc = ((EmuScene *)h)->d;
    Counters cn = {0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < n; i++) {
        const float3 o = f3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        const float3 d = normalized3(f3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        const ClosestHit hit = trace_closest<false>(sc, o, d, (flags & RT_FLAG_BACKFACE_CULLING) != 0, cn);
        out_t[i] = hit.slot >= 0 ? hit.t : __int_as_float(0x7f800000);
        out_prim[i] = hit.slot >= 0 ? hit.prim : 0xFFFFFFFFu;
    }
}
static void put(float *out, const ShadowResult &s) {
    out[0] = s.occluded ? 1.0f : 0.0f;
    out[1] = s.opacity;
    out[2] = s.filter.x;
    out[3] = s.filter.y;
    out[4] = s.filter.z;
}
void emu_trace_shadow(void *h, uint32_t flags, const float *rays, const float *max_dist, uint32_t n, float *out) {
    const DScene &sc = ((EmuScene *)h)->d;
    Counters cn = {0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < n; i++) {
        const float3 o = f3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        const float3 d = normalized3(f3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        put(out + 5 * i, trace_shadow<false>(sc, o, d, max_dist[i], (flags & RT_FLAG_BACKFACE_CULLING) != 0, cn));
    }
}

// One shadow beam: shading point P, box [lmin, lmax] of the light copies, and the M copies themselves.  Builds the
// beam as direct_kernel does, then answers every copy's shadow query twice: from the beam's candidate list
// (trace_shadow_list, with the last-occluder probe carried along) and by its own BVH walk (trace_shadow).
// Returns the number of candidates (-1: the list overflowed, -2 = BEAM_COVERED: one opaque triangle occludes every
// ray of the beam -- out_list is then all "occluded"), the box / triangle / beam-triangle test counts.
int emu_beam_shadows(void *h, uint32_t flags, const float *P3, const float *lmin3, const float *lmax3, float eps_d, const float *copies, uint32_t m,
                     float *out_list, float *out_ray, uint64_t *out_counts) {
    const DScene &sc = ((EmuScene *)h)->d;
    const bool cull = (flags & RT_FLAG_BACKFACE_CULLING) != 0;
    Counters cb = {0, 0, 0, 0, 0, 0}, cl = {0, 0, 0, 0, 0, 0}, cr = {0, 0, 0, 0, 0, 0};
    const float3 P = f3(P3[0], P3[1], P3[2]), lmin = f3(lmin3[0], lmin3[1], lmin3[2]), lmax = f3(lmax3[0], lmax3[1], lmax3[2]);
    const float pad = fmaf(0.25f, sc.box_pad, 2.0f * eps_d);
    const Beam beam = make_beam(P, lmin, lmax, pad);
    const float3 C = (lmin + lmax) * 0.5f, hh = (lmax - lmin) * 0.5f + splat(1e-6f * (fabsf(C.x) + fabsf(C.y) + fabsf(C.z)));
    uint32_t cand[BEAM_MAX_CAND];
    const int n_cand = beam_collect<true>(sc, beam, P, C - P, hh, cand, cb, cull ? -1.0f : eps_d);  // (as direct_kernel calls it)
    int32_t last_occluder = -1;
    for (uint32_t c = 0; c < m; c++) {
        const float3 lpos = f3(copies[3 * c], copies[3 * c + 1], copies[3 * c + 2]);
        const float3 ldir = normalized3(lpos - P);  // the ray of shade_pair (rt_kernels.cuh), generic IEEE forms
        const float3 q = P + ldir * eps_d;
        const float maxd = mag3(lpos - q);
        const float3 d = normalized3(ldir);
        put(out_ray + 5 * c, trace_shadow<true>(sc, q, d, maxd, cull, cr));
        if (n_cand == BEAM_COVERED) {
            ShadowResult occ;
            occ.occluded = true;
            occ.opacity = 0.0f;
            occ.filter = splat(1.0f);
            put(out_list + 5 * c, occ);
        } else if (n_cand >= 0)
            put(out_list + 5 * c, trace_shadow_list<true>(sc, cand, n_cand, last_occluder, q, d, maxd, cull, cl));
        else
            put(out_list + 5 * c, trace_shadow<true>(sc, q, d, maxd, cull, cl));
    }
    if (out_counts && n_cand > 0) {  // (diagnostics) the first candidates' reference primitive indices
        for (int i = 0; i < 3; i++) out_counts[5 + i] = i < n_cand ? __float_as_uint(__ldg(sc.prims + 4 * (size_t)cand[i] + 3).x) : 0xFFFFFFFFull;
    }
    if (out_counts) {
        out_counts[0] = cb.box;        // beam: box tests of the one hull walk
        out_counts[1] = cb.beam_prim;  // beam: triangle culling tests
        out_counts[2] = cl.tri + cl.sph;  // list: primitive tests over all copies
        out_counts[3] = cr.box;        // per ray: box tests over all copies
        out_counts[4] = cr.tri + cr.sph;  // per ray: primitive tests over all copies
    }
    return n_cand;
}

// light_box_behind (rt_device.cuh) against the per-copy test of shade_pair (rt_kernels.cuh): for n cases (P, N, box),
// out_behind[i] = the box test; out_facing[i] = how many of the case's m copies pass `dot3(N, normalize(x - P)) > 0`,
// evaluated with shade_pair's own operations.  The product relies on: behind => no copy faces.
void emu_light_box_behind(const float *P3, const float *N3, const float *lmin3, const float *lmax3, const float *copies, uint32_t n, uint32_t m,
                          int *out_behind, int *out_facing) {
    for (uint32_t i = 0; i < n; i++) {
        const float3 P = f3(P3[3 * i], P3[3 * i + 1], P3[3 * i + 2]), N = f3(N3[3 * i], N3[3 * i + 1], N3[3 * i + 2]);
        const float3 lmin = f3(lmin3[3 * i], lmin3[3 * i + 1], lmin3[3 * i + 2]), lmax = f3(lmax3[3 * i], lmax3[3 * i + 1], lmax3[3 * i + 2]);
        const float3 C = (lmin + lmax) * 0.5f, hh = (lmax - lmin) * 0.5f + splat(1e-6f * (fabsf(C.x) + fabsf(C.y) + fabsf(C.z)));  // (as direct_kernel)
        out_behind[i] = light_box_behind(N, C - P, hh) ? 1 : 0;
        int facing = 0;
        for (uint32_t c = 0; c < m; c++) {
            const float *x = copies + 3 * ((size_t)i * m + c);
            const float3 l2p = f3(x[0], x[1], x[2]) - P;
            const float d2 = dot3(l2p, l2p);
            const bool d2_ok = (d2 > 1e-30f) & (d2 < 1e30f);
            float3 ldir = l2p * rcp_rn_normal(sqrt_rn_normal(d2));
            if (!d2_ok) ldir = normalized3(l2p);
            if (smax(dot3(N, ldir), 0.0f) > 0.0f) facing++;
        }
        out_facing[i] = facing;
    }
}

}  // extern "C"
