// rt_kernels.cuh -- the render kernels (sm_100a).
//
// primary_kernel: closest hit of every primary sample, one thread per sample.
// direct_kernel (configs without bounces, the headline path): persistent warps fetch units = (32 neighbouring
// samples, one light's M soft-shadow copies); each lane builds ONE shadow beam for its sample (rt_device.cuh:
// hull walk -> candidate list -> triangle culling -> cover test) and shades the M copies in order, the running
// sums chained from light to light through tagged records; very expensive units are deferred and finished by
// all warps together (DEFER variant).
// render_kernel (bounce configs): persistent warps.  Every lane owns one primary sample at a time and walks that
// sample's reflection/refraction tree depth-first with a private path stack; lanes refill from a
// global work counter (warp-aggregated atomicAdd) the moment their tree is finished, so the
// divergent closest-hit phase always runs with as many live lanes as there is work.  The
// dominant cost -- L*M shadow/transmittance rays per shading point (reference
// raytracer_renderer.rs:773-870) -- is NOT done per lane: the live hit records of the warp are
// compacted into shared memory and the (hit, light copy) pairs are flattened over all 32 lanes,
// so the shadow traversal runs at full warp occupancy however few lanes still have a live path,
// and neighbouring lanes trace near-identical rays (same origin, jittered copies of one light); when most
// lanes have a hit, each lights its own hit through shadow beams instead (beam_lighting).
// resolve_kernel: sums the samples of a pixel in the reference's order, encodes to 0xFFRRGGBB.
#pragma once
#include "rt_device.cuh"

namespace rtd {

// Occupancy sweep on B200 (profiles/README.md): the kernel is latency bound, 28 warps/SM at 72
// registers (a few hundred bytes of spills) beats 16 warps/SM at 125 registers without spills by 13 %.
#ifndef RT_EMPTY_LIST_SKIP
#define RT_EMPTY_LIST_SKIP 1  // a beam without candidates: no list walk (see RT_SELF_CULL in rt_device.cuh)
#endif
#ifndef RT_CTA_THREADS
#ifndef RT_FACING_FIRST
#define RT_FACING_FIRST 1  // no shadow query towards a light the surface faces away from, no beam towards a light box entirely behind it
#endif
#define RT_CTA_THREADS 128
#endif
#ifndef RT_MIN_CTAS_PER_SM
#define RT_MIN_CTAS_PER_SM 7
#endif
constexpr int CTA_THREADS = RT_CTA_THREADS;
constexpr int CTA_WARPS = CTA_THREADS / 32;
constexpr int PATH_STACK = 24;   // >= 1 + max(reflection, refraction) depth (21) + slack
constexpr int MAX_AA = 24;
constexpr int MAX_SM_LIGHTS = 8; // cloud indices cached in shared memory for up to 8 lights
constexpr int TILE_W = 8, TILE_H = 4;  // work-item order: 8x4-pixel tiles, samples innermost

// Wavefront form of the bounce configs (paths_kernel -> direct_kernel<.., POINTS> -> gather_paths_kernel): one
// shading point of a sample's reflection / refraction tree, 64 B.
struct PathPoint {
    float4 a;  // P.xyz, t (distance of the hit along its ray: attenuation)
    float4 b;  // N.xyz, bits(material index)
    float4 c;  // d.xyz (normalised direction of the ray that hit), bits(pixel x | pixel y << 16)
    float4 w;  // weight.rgb of the point's local colour in its sample, bits(link: the sample's NEXT point in path order, -1 at its last)
};
struct PathQueue {
    PathPoint *points;
    float4 *contrib;           // per point: weight * local colour (direct_kernel<.., POINTS>)
    int32_t *sample_head;      // per sample of the chunk: its FIRST point (-1: the primary ray missed)
    unsigned int *n_points;    // points appended so far (may exceed capacity: overflow)
    uint32_t capacity;
    uint32_t first_item, n_chunk_items;  // the chunk of primary samples (work items) being traced
};

struct RenderParams {
    DScene sc;
    uint32_t width, height, flags;
    uint32_t S;            // primary samples per pixel
    uint32_t M, LM;        // light copies per light, total light copies
    float inv_M;           // 1 / M as f32 (light.rs:194)
    float w2s_w, w2s_h, w2s_d;
    float focus_x, focus_y, focus_z;
    float eps_d;           // default_epsilon_distance component (vector.rs:697-700)
    float aa_off[2 * MAX_AA];
    float sample_scale;    // 1 / S
    const float4 *clouds;  // K x M cloud points, already scaled to scene units: p * (w2s_w, w2s_h, w2s_d) (scale_clouds_kernel)
    float cloud_lo[3], cloud_hi[3];  // component-wise bounds of all cloud points (shadow beams)
    uint32_t K;
    unsigned long long seed;
    uint32_t refl_depth, refr_depth;
    uint32_t tile_rows, tile_rank, tile_world, local_rows;
    uint32_t tiles_x, n_items;
    uint32_t tail_div;     // 2 x warps in the grid: divisor of the guided self-scheduling
    uint32_t heavy_cap;    // direct_kernel: deferred units per launch (<= HEAVY_CAP; beyond: processed in place)
    uint32_t defer_all;    // tests only (RT_DEFER_ALL): every unit asks for deferral
    uint32_t defer_cand;   // DEFER: a beam with more candidates than this makes its unit a heavy one (overflowed beams always do)
    uint32_t chain_window; // direct_kernel: batches per window of the chained records (see direct_kernel)
    uint32_t lights_per_unit;  // direct_kernel (soft shadows, not DEFER): lights a unit takes in a row (>= 1)
    unsigned int *chain_done;  // per batch of a window: tag of the window's last unit of that batch (slot reuse)
    PathQueue pq;          // bounce configs: the shading points of the chunk being rendered
    float4 *samples;               // local_rows * width * S : (r, g, b, valid)
    uint4 *primary_hits;           // n_items: closest hit of every primary sample (primary_kernel): bits t, slot, prim, meta
    unsigned int *work_counter;
    unsigned long long *counters;  // closest, shadow, box, tri, sph
};

RT_DEV uint32_t cloud_index(unsigned long long seed, uint32_t x, uint32_t y, uint32_t light, uint32_t n) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(x + 1u) +
                           0xBF58476D1CE4E5B9ull * (unsigned long long)(y + 1u) + 0x94D049BB133111EBull * (unsigned long long)(light + 1u);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(((z >> 32) * (unsigned long long)n) >> 32);
}

RT_DEV uint32_t tile_image_row(const RenderParams &p, uint32_t local_row) {
    if (p.tile_world <= 1 || p.tile_rows == 0) return local_row;
    uint32_t t = local_row / p.tile_rows, r = local_row - t * p.tile_rows;
    return (t * p.tile_world + p.tile_rank) * p.tile_rows + r;
}

// Work item -> pixel and sample. Items are ordered in 8x4-pixel tiles of the rank's local image,
// samples innermost, so the 32 items of a warp are 2 neighbouring pixels x 16 samples (coherent).
struct ItemInfo {
    bool valid;
    uint32_t x, ly, s;
};
RT_DEV ItemInfo decode_item(const RenderParams &p, uint32_t item) {
    ItemInfo it;
    uint32_t pix = item / p.S;
    it.s = item - pix * p.S;
    uint32_t tile = pix / (TILE_W * TILE_H), within = pix - tile * (TILE_W * TILE_H);
    uint32_t ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    it.x = tx * TILE_W + (within % TILE_W);
    it.ly = ty * TILE_H + (within / TILE_W);
    it.valid = item < p.n_items && it.x < p.width && it.ly < p.local_rows;
    return it;
}

// get_pixel_color / render_pixel_colors (raytracer_renderer.rs:1140-1357): origin on the z = 0 plane
// (+ AA offset), direction = pixel centre - focus (un-normalised; Ray::new normalises)
RT_DEV void primary_ray(const RenderParams &p, uint32_t pix_x, uint32_t pix_y, uint32_t s, float3 &o, float3 &dir) {
    float cx = (float)pix_x * p.w2s_w, cy = (float)pix_y * p.w2s_h;
    dir = f3(cx, cy, 0.0f) - f3(p.focus_x, p.focus_y, p.focus_z);
    o = f3(cx + p.aa_off[2 * s], cy + p.aa_off[2 * s + 1], 0.0f);
}

// One (shading point, light copy) pair: the body of the `for light in lights` loop of
// calculate_lighting (raytracer_renderer.rs:773-870) with PointLight::to_point_light_cloud
// (light.rs:183-225), the shadow query and PointLight::calculate_contribution_at (light.rs:261-299).
// Returns the diffuse and specular terms this light copy adds (zero when occluded / facing away).
// One light copy as the lighting loop sees it (PointLight::to_point_light_cloud, light.rs:183-225).
struct LightCopy {
    float3 pos, color;
    float intensity;
};
// light `li` itself; with soft shadows the intensity is already the per-copy intensity / M (light.rs:194)
RT_DEV LightCopy load_light(const RenderParams &p, uint32_t li) {
    const float4 l0 = __ldg(p.sc.lights + 2 * li), l1 = __ldg(p.sc.lights + 2 * li + 1);
    LightCopy L;
    L.pos = f3(l0.x, l0.y, l0.z);
    L.color = f3(l1.x, l1.y, l1.z);
    L.intensity = p.M > 1 ? p.inv_M * l0.w : l0.w;
    return L;
}
// copy `cpy` of cloud `cloud`: position + p * (w2s_w, w2s_h, w2s_d) (the product is precomputed per point)
RT_DEV float3 copy_position(const RenderParams &p, float3 lpos, uint32_t cloud, uint32_t cpy) {
    const float4 rp = __ldg(p.clouds + (size_t)cloud * p.M + cpy);
    return lpos + f3(rp.x, rp.y, rp.z);
}

// `cand` / `n_cand` (>= 0): the beam's candidate list for this (shading point, light), see beam_collect;
// n_cand < 0: the shadow query walks the BVH itself.
// MODE 0: the shadow query walks the BVH itself (inlined).  MODE 1: the caller built a beam; its overflow
// fallback (n_cand < 0) is the out-of-line traversal.  MODE 2: out-of-line traversal only.
// Returns false when the pair adds nothing (light behind the surface or completely occluded: dc = spc = 0, and the
// reference does not add either) -- the caller may leave its running sums alone.
template <bool COUNT, int MODE = 0>
RT_DEV bool shade_pair(const RenderParams &p, float3 hp, float3 hn, float3 view, float3 mcolor, float mshininess, float3 lpos, float lint,
                       float3 lcol, bool cull, Counters &cn, float3 &dc, float3 &spc, const uint32_t *cand = nullptr, int n_cand = -1,
                       int32_t *last_occluder = nullptr) {
    const DScene &sc = p.sc;
    dc = splat(0.0f);
    spc = splat(0.0f);
    // The shadow ray exactly as the reference builds it (light.rs:261-299, raytracer_renderer.rs:773-800):
    //   ldir = normalize(lpos - hp), q = hp + ldir * eps_d, maxd = |lpos - q|, d = normalize(ldir) (Ray::new).
    // Three IEEE square roots and two IEEE reciprocals per ray; in the normal range (always, short of a light
    // sitting on the surface) they are taken without nvcc's per-call range checks (rt_device.cuh).
    const float3 l2p = lpos - hp;
    const float d2 = dot3(l2p, l2p);
    const bool d2_ok = (d2 > 1e-30f) & (d2 < 1e30f);  // inside the range of the sequences above (false for NaN)
    float3 ldir = l2p * rcp_rn_normal(sqrt_rn_normal(d2));
    if (!d2_ok) ldir = normalized3(l2p);
    // `light_valid` (raytracer_renderer.rs:849): a light the surface does not face adds exactly nothing, whatever its
    // shadow query answers -- so the query is not made (RT_FACING_FIRST; the reference casts the ray and drops the result).
    const float diffuse = smax(dot3(hn, ldir), 0.0f);
    if (RT_FACING_FIRST && !(diffuse > 0.0f)) {
        if (COUNT) cn.shadow++;
        return false;
    }
    ShadowResult sh;
    sh.occluded = false;
    sh.opacity = 1.0f;
    sh.filter = splat(1.0f);
    // An empty candidate list (and no occluder remembered from an earlier copy): the query is "nothing on the
    // segment" whatever the ray, so the ray itself -- origin offset, length, second normalisation: two of the three
    // square roots -- is not built at all.
    const bool nothing_to_test = RT_EMPTY_LIST_SKIP && MODE == 1 && n_cand == 0 && *last_occluder < 0;
    if (nothing_to_test) {
        if (COUNT) cn.shadow++;
    } else {
        const float3 q = hp + ldir * p.eps_d;
        const float3 lq = lpos - q;
        const float m2 = dot3(lq, lq);
        float maxd;
        float3 d;
        if (d2_ok & (m2 > 1e-30f)) {
            maxd = sqrt_rn_normal(m2);
            d = ldir * rcp_rn_normal(sqrt_rn_normal(dot3(ldir, ldir)));
        } else {
            maxd = mag3(lq);
            d = normalized3(ldir);
        }
        if (MODE == 1)
            sh = n_cand >= 0 ? trace_shadow_list<COUNT>(sc, cand, n_cand, *last_occluder, q, d, maxd, cull, cn)
                             : trace_shadow_outlined<COUNT>(sc, q, d, maxd, cull, &cn);
        else if (MODE == 2)
            sh = trace_shadow_outlined<COUNT>(sc, q, d, maxd, cull, &cn);
        else
            sh = trace_shadow<COUNT>(sc, q, d, maxd, cull, cn);
    }
    // Everything below only matters when the light is unoccluded AND the surface faces it
    // (`light_valid`, raytracer_renderer.rs:849): otherwise the pair adds exactly zero.
    if (sh.occluded || !(diffuse > 0.0f)) return false;
    const float ldist = sqrt_fast(dot3(l2p, l2p)) + EPS;  // from here on: colour only (fast math, rt_device.cuh)
    const float cosn = dot3(l2p, hn);                       // cosi = cosn / ldist > 0 <=> cosn > 0 (ldist > 0)
    if (cosn > 0.0f) {  // incident_angle_pos; else colour and intensity are zero
        const float cosi = cosn * rcp_fast(ldist);
        const float att = 0.95f * (EPS + ldist + ldist * ldist);
        const float sig = (tanh_pos_fast(att) + 1.0f) * 0.5f;
        const float cint = cosi * lint * sclamp(sig, 0.0f, 1.0f);
        const float3 ccolor = mcolor * lcol;
        // colour / filter; x / 1.0f == x exactly, and the filter is (1,1,1) unless the ray crossed a
        // transmissive primitive
        const bool unit_filter = sh.filter.x == 1.0f && sh.filter.y == 1.0f && sh.filter.z == 1.0f;
        const float3 lcs = unit_filter ? ccolor : ccolor * f3(rcp_fast(sh.filter.x), rcp_fast(sh.filter.y), rcp_fast(sh.filter.z));
        float spec = 0.0f;
        const bool has_spec = mshininess > 0.0f;
        if (has_spec) {
            const float3 rr = reflected3(ldir, hn);
            const float sd = smax(dot3(rr, view) * rsqrt_fast(dot3(rr, rr)), 0.0f);
            spec = pow_pos_fast(sd, smax(mshininess * 512.0f, 1.0f));
        }
        const float lf = diffuse * cint * sh.opacity;
        const float sf = cint * sh.opacity * spec;
        dc = mcolor * lcs * lf;
        if (has_spec) spc = lcol * sf;
    } else {
        // cosi <= 0 with diffuse > 0 (rounding only): contribution colour and intensity are zero; the
        // specular term is l.color * (0 * opacity * spec) = 0 exactly.
        dc = mcolor * (splat(0.0f) / sh.filter) * (diffuse * 0.0f * sh.opacity);
    }
    return true;
}

struct PathRec {  // one pending ray of a sample's reflection/refraction tree (48 B)
    float4 a;     // origin.xyz, ior
    float4 b;     // direction.xyz (normalised once; Ray::new normalises again), bits(depth | flags)
    float4 c;     // weight.rgb, unused
};
constexpr uint32_t PATH_REFL_CHILD = 1u << 16;  // whole subtree is scaled by attenuation(t) of its first hit

struct WarpShared {
    float4 rec[32][4];               // hit records: (P, n.x) (n.yz, v.xy) (v.z, bits mat, bits pixel x, bits pixel y) (material colour, shininess)
    float res[32][6];                // per pair: diffuse rgb, specular rgb (only when fewer than 32 hits are live)
    uint32_t cloud[32][MAX_SM_LIGHTS];
};

struct LightSums {  // running diffuse / specular light sums of one sample
    float3 d, s;
};
// All L x M light copies for ONE shading point, in the reference's light order, with one shadow beam per
// light (rt_device.cuh): what a lane of render_kernel runs for its own hit when most lanes of the warp have
// one.  Out of line, parameters from their global-memory copy: render_kernel is register-starved as it is.
constexpr uint32_t BEAM_MIN_LIVE = 12;  // below: the (hit, copy) pairs are spread over all 32 lanes instead
template <bool COUNT>
__device__ __noinline__ LightSums beam_lighting(const RenderParams *gp, float3 P, float3 N, float3 view, float3 mcolor, float mshininess,
                                                uint32_t cloud_x, uint32_t pix_y, Counters *cn_out) {
    const RenderParams &p = *gp;
    const DScene &sc = p.sc;
    const bool cull = (p.flags & RT_FLAG_BACKFACE_CULLING) != 0;
    Counters cn = {0, 0, 0, 0, 0, 0};
    LightSums out;
    out.d = out.s = splat(0.0f);
    const float3 w2s = f3(p.w2s_w, p.w2s_h, p.w2s_d);
    const float3 clo = f3(p.cloud_lo[0], p.cloud_lo[1], p.cloud_lo[2]) * w2s, chi = f3(p.cloud_hi[0], p.cloud_hi[1], p.cloud_hi[2]) * w2s;
    const float pad = fmaf(0.25f, sc.box_pad, 2.0f * p.eps_d);
    uint32_t cand[BEAM_MAX_CAND];
    for (uint32_t li = 0; li < sc.n_lights; li++) {
        const LightCopy L = load_light(p, li);
        const uint32_t cloud = cloud_index(p.seed, cloud_x, pix_y, li, p.K);
        const float3 lmin = L.pos + clo, lmax = L.pos + chi;  // the copies' `lpos + rp * w2s` is monotonic in rp
        const float3 C = (lmin + lmax) * 0.5f, hh = (lmax - lmin) * 0.5f + splat(1e-6f * (fabsf(C.x) + fabsf(C.y) + fabsf(C.z)));
        int n_cand = BEAM_COVERED;  // (a light box behind the surface is treated like a covered beam: nothing to add)
        if (!(RT_FACING_FIRST && light_box_behind(N, C - P, hh))) {
            const Beam beam = make_beam(P, lmin, lmax, pad);
            n_cand = beam_collect<COUNT>(sc, beam, P, C - P, hh, cand, cn, cull ? -1.0f : p.eps_d);
        }
        if (n_cand == BEAM_COVERED) {  // one opaque triangle occludes all M copies: M occluded queries, nothing to add
            if (COUNT) cn.shadow += p.M;
            continue;
        }
        int32_t last_occluder = -1;
        for (uint32_t c = 0; c < p.M; c++) {
            float3 dc, spc;
            shade_pair<COUNT, 1>(p, P, N, view, mcolor, mshininess, copy_position(p, L.pos, cloud, c), L.intensity, L.color, cull, cn, dc, spc, cand, n_cand,
                                 &last_occluder);
            out.d = out.d + dc;
            out.s = out.s + spc;
        }
    }
    if (COUNT && cn_out) {
        cn_out->closest += cn.closest; cn_out->shadow += cn.shadow; cn_out->box += cn.box;
        cn_out->tri += cn.tri; cn_out->sph += cn.sph; cn_out->beam_prim += cn.beam_prim;
    }
    return out;
}

// gp: global-memory copy of `p` when shadow beams may be used (M > 1, usable cloud bounds), else NULL
template <bool COUNT>
__global__ void __launch_bounds__(CTA_THREADS, RT_MIN_CTAS_PER_SM) render_kernel(const RenderParams p, const RenderParams *gp) {
    __shared__ WarpShared smem[CTA_WARPS];
    WarpShared &ws = smem[threadIdx.x >> 5];
    stage_top_nodes(p.sc);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t FULL = 0xFFFFFFFFu;
    const DScene &sc = p.sc;
    const bool cull = (p.flags & RT_FLAG_BACKFACE_CULLING) != 0;
    const bool refl_on = (p.flags & RT_FLAG_REFLECTIONS) != 0;
    const bool refr_on = (p.flags & RT_FLAG_REFRACTIONS) != 0;
    const bool packet_noaa = (p.flags & RT_FLAG_PACKET_SEMANTICS) && !(p.flags & RT_FLAG_ANTI_ALIASING);
    Counters cn = {0, 0, 0, 0, 0, 0};

    PathRec stack[PATH_STACK];
    int sp = 0;
    bool have_item = false, have_ray = false;
    uint32_t sample_index = 0, item_index = 0, pix_x = 0, pix_y = 0;
    float3 acc = splat(0.0f);
    bool sample_valid = false;
    // current ray
    float3 ro = splat(0.0f), rd = splat(0.0f), rw = splat(0.0f);
    float r_ior = AIR_IOR;
    int r_depth = -1;  // -1 = None
    bool r_refl_child = false, r_primary = false;

    for (;;) {
        // ---------------------------------------------------------------- refill
        if (!have_ray && sp > 0) {
            PathRec pr = stack[--sp];
            ro = f3(pr.a.x, pr.a.y, pr.a.z);
            r_ior = pr.a.w;
            rd = f3(pr.b.x, pr.b.y, pr.b.z);
            uint32_t bits = __float_as_uint(pr.b.w);
            r_depth = (int)(bits & 0xFFFFu);
            r_refl_child = (bits & PATH_REFL_CHILD) != 0;
            r_primary = false;
            rw = f3(pr.c.x, pr.c.y, pr.c.z);
            have_ray = true;
        }
        if (!have_ray && have_item) {  // tree finished: publish the sample
            p.samples[sample_index] = make_float4(acc.x, acc.y, acc.z, sample_valid ? 1.0f : 0.0f);
            have_item = false;
        }
        bool exhausted = false;  // the fetch of this round found the queue empty
        {
            bool need = !have_ray;
            uint32_t need_mask = __ballot_sync(FULL, need);
            if (need_mask) {
                // Guided self-scheduling: a warp normally takes one item per idle lane (up to 32), but
                // near the end of the frame only its fair share of what is left, so that the last
                // rounds are short and all warps finish together. Small rounds cost nothing extra:
                // the (hit, light) flattening below fills the lanes whatever the number of live hits.
                uint32_t base = 0, grant = 0;
                const int leader = __ffs(need_mask) - 1;
                if ((int)lane == leader) {
                    const uint32_t done = *(volatile unsigned int *)p.work_counter;
                    const uint32_t left = done < p.n_items ? p.n_items - done : 0u;
                    grant = min((uint32_t)__popc(need_mask), max(min(left / p.tail_div + 1u, 32u), 4u));
                    base = atomicAdd(p.work_counter, grant);
                }
                base = __shfl_sync(FULL, base, leader);
                grant = __shfl_sync(FULL, grant, leader);
                exhausted = base >= p.n_items;
                const uint32_t rank = __popc(need_mask & ((1u << lane) - 1u));
                const uint32_t item = base + rank;
                const ItemInfo it = decode_item(p, item);
                if (need && rank < grant && it.valid) {
                    pix_x = it.x;
                    pix_y = tile_image_row(p, it.ly);
                    sample_index = (it.ly * p.width + it.x) * p.S + it.s;
                    item_index = item;
                    primary_ray(p, pix_x, pix_y, it.s, ro, rd);
                    rw = splat(1.0f);
                    r_ior = AIR_IOR;
                    r_depth = -1;
                    r_refl_child = false;
                    r_primary = true;
                    acc = splat(0.0f);
                    sample_valid = false;
                    have_item = true;
                    have_ray = true;
                }
            }
        }
        if (!__any_sync(FULL, have_ray)) {  // (a round of nothing but padding items must not end the warp: see paths_kernel)
            if (exhausted) break;
            continue;
        }

        // ---------------------------------------------------------------- closest hit (per lane)
        float3 d = splat(0.0f), P = splat(0.0f), N = splat(0.0f);
        ClosestHit h;
        h.slot = -1;
        h.t = 0.0f;
        h.meta = 0;
        h.prim = 0;
        if (have_ray) {
            d = normalized3(rd);  // Ray::new_with_mask (ray.rs:52-57)
            if (r_primary) {  // traced by primary_kernel at full width; here only fetched
                const uint4 ph = p.primary_hits[item_index];
                h.t = __uint_as_float(ph.x);
                h.slot = (int32_t)ph.y;
                h.prim = ph.z;
                h.meta = ph.w;
                if (COUNT) cn.closest++;
            } else {
                h = trace_closest<COUNT>(sc, ro, d, cull, cn);
            }
        }
        const bool hit = have_ray && h.slot >= 0;
        if (hit) {
            P = ray_at(ro, d, h.t);
            if (h.meta & rtb::PRIM_SPHERE) {
                float4 q0 = __ldg(sc.prims + 4 * (size_t)h.slot);
                N = normalized3(P - f3(q0.x, q0.y, q0.z));
            } else {
                float4 sn = __ldg(sc.shade_n + h.slot);
                N = f3(sn.x, sn.y, sn.z);
            }
            if (r_primary) sample_valid = true;
        }

        // ---------------------------------------------------------------- cooperative lighting
        const uint32_t live = __ballot_sync(FULL, hit);
        const uint32_t n_live = __popc(live);
        const uint32_t my_slot = __popc(live & ((1u << lane) - 1u));
        if (hit) {
            ws.rec[my_slot][0] = make_float4(P.x, P.y, P.z, N.x);
            ws.rec[my_slot][1] = make_float4(N.y, N.z, d.x, d.y);
            ws.rec[my_slot][2] = make_float4(d.z, __uint_as_float(h.meta & rtb::PRIM_MATERIAL_MASK), __uint_as_float(pix_x), __uint_as_float(pix_y));
            const float4 ma = __ldg(sc.mats + 3 * (h.meta & rtb::PRIM_MATERIAL_MASK)), mb = __ldg(sc.mats + 3 * (h.meta & rtb::PRIM_MATERIAL_MASK) + 1);
            ws.rec[my_slot][3] = make_float4(ma.x, ma.y, ma.z, mb.x);
        }
        // lane s < n_live accumulates the light sums of hit slot s (in light order, like the
        // reference's `for light in lights`); when all 32 lanes are live, slot == lane == owner.
        float3 sum_d = splat(0.0f), sum_s = splat(0.0f);
        // most lanes have a hit and the lights have M > 1 copies: every lane lights its OWN hit, one shadow beam
        // per light (2.5x fewer instructions per ray than a BVH walk per copy; same sums in the same order)
        const bool lane_beams = gp != nullptr && n_live >= BEAM_MIN_LIVE;
        if (lane_beams) {
            if (hit) {
                const float4 ma = __ldg(sc.mats + 3 * (h.meta & rtb::PRIM_MATERIAL_MASK)), mb = __ldg(sc.mats + 3 * (h.meta & rtb::PRIM_MATERIAL_MASK) + 1);
                const LightSums ls = beam_lighting<COUNT>(gp, P, N, d, f3(ma.x, ma.y, ma.z), mb.x, packet_noaa ? pix_x / 8u : pix_x, pix_y, COUNT ? &cn : nullptr);
                sum_d = ls.d;
                sum_s = ls.s;
            }
        }
        __syncwarp();
        if (!lane_beams) {
        if (p.M > 1) {  // SceneLightSource::preprocess: one cloud per (pixel, light) (light.rs:183-225)
            for (uint32_t i = lane; i < n_live * MAX_SM_LIGHTS; i += 32) {
                uint32_t s = i / MAX_SM_LIGHTS, li = i % MAX_SM_LIGHTS;
                if (li < sc.n_lights) {
                    uint32_t x = __float_as_uint(ws.rec[s][2].z), y = __float_as_uint(ws.rec[s][2].w);
                    ws.cloud[s][li] = cloud_index(p.seed, packet_noaa ? x / 8u : x, y, li, p.K);
                }
            }
            __syncwarp();
        }
        const uint32_t total = n_live * p.LM;
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t pair = base + lane;
            float3 dc = splat(0.0f), spc = splat(0.0f);
            if (pair < total) {
                // light-major flattening: a batch holds ONE light copy for (up to) 32 neighbouring hit
                // points, so the 32 shadow rays converge on the same point and traverse together
                // (n_live == 32: l = base / 32 and s = lane, no division)
                const uint32_t l = n_live == 32 ? (base >> 5) : pair / n_live, s = pair - l * n_live;
                const float4 r0 = ws.rec[s][0], r1 = ws.rec[s][1], r2 = ws.rec[s][2], r3 = ws.rec[s][3];
                const float3 hp = f3(r0.x, r0.y, r0.z), hn = f3(r0.w, r1.x, r1.y), view = f3(r1.z, r1.w, r2.x);
                const float3 mcolor = f3(r3.x, r3.y, r3.z);
                const float mshininess = r3.w;
                uint32_t li = l, cpy = 0;
                if (p.M > 1) {
                    li = l / p.M;
                    cpy = l - li * p.M;
                }
                const uint32_t ci = p.M > 1 ? (li < MAX_SM_LIGHTS ? ws.cloud[s][li]
                                                                   : cloud_index(p.seed, packet_noaa ? __float_as_uint(r2.z) / 8u : __float_as_uint(r2.z),
                                                                                 __float_as_uint(r2.w), li, p.K))
                                            : 0u;
                const LightCopy L = load_light(p, li);
                shade_pair<COUNT>(p, hp, hn, view, mcolor, mshininess, p.M > 1 ? copy_position(p, L.pos, ci, cpy) : L.pos, L.intensity, L.color, cull,
                                  cn, dc, spc);
            }
            if (n_live == 32) {
                sum_d = sum_d + dc;  // pair (l, s = lane): this lane's own hit, next light in order
                sum_s = sum_s + spc;
            } else {
                ws.res[lane][0] = dc.x; ws.res[lane][1] = dc.y; ws.res[lane][2] = dc.z;
                ws.res[lane][3] = spc.x; ws.res[lane][4] = spc.y; ws.res[lane][5] = spc.z;
                __syncwarp();
                if (lane < n_live) {
                    // batch entry j holds pair base + j = (light (base + j) / n_live, slot (base + j) % n_live)
                    const uint32_t j0 = (lane + n_live - base % n_live) % n_live;
                    for (uint32_t j = j0; j < 32u && base + j < total; j += n_live) {  // increasing light index
                        sum_d = sum_d + f3(ws.res[j][0], ws.res[j][1], ws.res[j][2]);
                        sum_s = sum_s + f3(ws.res[j][3], ws.res[j][4], ws.res[j][5]);
                    }
                }
                __syncwarp();
            }
        }
        // hand the sums of slot `my_slot` (held by lane my_slot) to the lane that owns the hit
        {
            const int src = hit ? (int)my_slot : (int)lane;
            sum_d = f3(__shfl_sync(FULL, sum_d.x, src), __shfl_sync(FULL, sum_d.y, src), __shfl_sync(FULL, sum_d.z, src));
            sum_s = f3(__shfl_sync(FULL, sum_s.x, src), __shfl_sync(FULL, sum_s.y, src), __shfl_sync(FULL, sum_s.z, src));
        }
        }  // !lane_beams

        // ---------------------------------------------------------------- combine + spawn (owner lane)
        if (hit) {
            const Material m = load_material(sc, h.meta & rtb::PRIM_MATERIAL_MASK);
            float3 light_color = sum_d;
            float3 specular = sum_s;
            const float3 ambient = (m.color * splat(1.0f)) * 0.08f;
            float3 direct = ambient + light_color;
            const float df = attenuation(h.t);
            direct = direct * df;
            specular = specular * df;
            // single_raytrace (raytracer_renderer.rs:147-264), linearised: a reflection child's whole
            // subtree is scaled by attenuation(child.t) (raytracer_renderer.rs:722-727)
            const float3 w = r_refl_child ? (rw * df) : rw;
            const bool is_transmissive = m.transmissive;
            const float3 local = is_transmissive ? specular : (direct + specular);
            acc = acc + w * local;
            const bool is_reflective = (m.metallic > 0.0f) || is_transmissive;
            const float start_ior = r_ior;
            if (refl_on && is_reflective) {  // calculate_reflection (raytracer_renderer.rs:526-729)
                float ct = dot3(d, N);
                bool inside = ct < 0.0f;
                float3 inormal = inside ? -N : N;
                float new_ior = inside ? m.ior : AIR_IOR;
                float eta = inside ? (new_ior / start_ior) : (start_ior / new_ior);
                float ci = fabsf(ct);
                float sin2 = eta * eta * (1.0f - ci * ci);
                bool tir = sin2 >= 1.0f;
                bool reflective = (m.metallic > 0.0f) || (is_transmissive && tir);
                if (reflective) {
                    float3 rdir = normalized3(reflected3(d, N));
                    bool none = (fabsf(rdir.x - 0.0f) <= EPS) && (fabsf(rdir.y - 0.0f) <= EPS) && (fabsf(rdir.z - 0.0f) <= EPS);
                    int next = r_depth < 0 ? (int)p.refl_depth : max(r_depth - 1, 0);
                    if (!none && next > 0 && sp < PATH_STACK) {
                        float3 reflectance, tmp;
                        compute_fresnel(m, inormal, -d, start_ior, reflectance, tmp);
                        float3 co = P + rdir * p.eps_d;
                        float3 cw = w * reflectance;
                        stack[sp].a = make_float4(co.x, co.y, co.z, start_ior);
                        stack[sp].b = make_float4(rdir.x, rdir.y, rdir.z, __uint_as_float((uint32_t)next | PATH_REFL_CHILD));
                        stack[sp].c = make_float4(cw.x, cw.y, cw.z, 0.0f);
                        sp++;
                    }
                }
            }
            if (refr_on && is_transmissive) {  // calculate_refractions (raytracer_renderer.rs:279-524)
                float ct = dot3(d, N);
                bool entering = ct <= 0.0f;
                float3 inormal = entering ? -N : N;
                float new_ior = entering ? m.ior : AIR_IOR;
                float eta = entering ? (new_ior / start_ior) : (start_ior / new_ior);
                float inv_eta = 1.0f / eta;
                float3 tmp, transmittance;
                compute_fresnel(m, inormal, d, inv_eta, tmp, transmittance);
                float3 rdir = normalized3(refracted3(d, -inormal, inv_eta));
                float op = m.opacity;
                int step = (op < 0.5f) ? 2 : 1;
                int factor = (op <= 0.3f) ? 3 : ((op < 0.5f) ? 2 : 1);
                int next = r_depth < 0 ? (int)p.refr_depth / factor : max(r_depth - step, 0);
                if (next > 0 && sp < PATH_STACK) {
                    float boost = m.boost + 1.0f;
                    float3 co = P + rdir * p.eps_d;
                    float3 cw = w * (transmittance * boost);
                    stack[sp].a = make_float4(co.x, co.y, co.z, new_ior);
                    stack[sp].b = make_float4(rdir.x, rdir.y, rdir.z, __uint_as_float((uint32_t)next));
                    stack[sp].c = make_float4(cw.x, cw.y, cw.z, 0.0f);
                    sp++;
                }
            }
        }
        have_ray = false;
        __syncwarp();
    }

    if (COUNT) {
        atomicAdd(p.counters + 0, cn.closest);
        atomicAdd(p.counters + 1, cn.shadow);
        atomicAdd(p.counters + 2, cn.box);
        atomicAdd(p.counters + 3, cn.tri);
        atomicAdd(p.counters + 4, cn.sph);
        atomicAdd(p.counters + 5, cn.beam_prim);
    }
}

// ---- wavefront form of the bounce configs -------------------------------------------------------------------
// render_kernel above lights every shading point where it finds it, inside the divergent path walk: thin bounce
// tails leave most lanes of a warp idle exactly where 95-140 shadow rays per point are due.  The wavefront form
// separates the two: paths_kernel walks the reflection / refraction trees (closest-hit rays only -- which children a
// hit spawns and with which weights depends on geometry and material alone, not on the lighting) and appends every
// shading point to a queue; direct_kernel<.., POINTS> then lights the queue like it lights primary hits -- full
// warps of 32 points, one shadow beam per (point, light), chained light groups, deferral of very expensive units;
// gather_paths_kernel adds a sample's contributions in the order the path walk met them (the same float sums as
// render_kernel's `acc + w * local`, so the frame is bit-identical).  The frame is processed in chunks of primary
// samples so that the queue stays bounded.
template <bool COUNT>
__global__ void __launch_bounds__(CTA_THREADS, 4) paths_kernel(const RenderParams p) {
    stage_top_nodes(p.sc);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t FULL = 0xFFFFFFFFu;
    const DScene &sc = p.sc;
    const PathQueue &pq = p.pq;
    const bool cull = (p.flags & RT_FLAG_BACKFACE_CULLING) != 0;
    const bool refl_on = (p.flags & RT_FLAG_REFLECTIONS) != 0;
    const bool refr_on = (p.flags & RT_FLAG_REFRACTIONS) != 0;
    Counters cn = {0, 0, 0, 0, 0, 0};

    PathRec stack[PATH_STACK];
    int sp = 0;
    bool have_item = false, have_ray = false;
    uint32_t sample_slot = 0, item_index = 0, pix_x = 0, pix_y = 0;
    int32_t first_point = -1, last_point = -1;  // the sample's point list: forward links, written by this lane as the points arrive
    float3 ro = splat(0.0f), rd = splat(0.0f), rw = splat(0.0f);
    float r_ior = AIR_IOR;
    int r_depth = -1;  // -1 = None
    bool r_refl_child = false, r_primary = false;

    for (;;) {
        // ---------------------------------------------------------------- refill (as render_kernel)
        if (!have_ray && sp > 0) {
            PathRec pr = stack[--sp];
            ro = f3(pr.a.x, pr.a.y, pr.a.z);
            r_ior = pr.a.w;
            rd = f3(pr.b.x, pr.b.y, pr.b.z);
            uint32_t bits = __float_as_uint(pr.b.w);
            r_depth = (int)(bits & 0xFFFFu);
            r_refl_child = (bits & PATH_REFL_CHILD) != 0;
            r_primary = false;
            rw = f3(pr.c.x, pr.c.y, pr.c.z);
            have_ray = true;
        }
        if (!have_ray && have_item) {  // tree finished: publish the head of the sample's point list
            pq.sample_head[sample_slot] = first_point;
            have_item = false;
        }
        bool exhausted = false;  // the fetch of this round found the queue empty
        {
            const bool need = !have_ray;
            const uint32_t need_mask = __ballot_sync(FULL, need);
            if (need_mask) {
                uint32_t base = 0, grant = 0;
                const int leader = __ffs(need_mask) - 1;
                if ((int)lane == leader) {
                    const uint32_t done = *(volatile unsigned int *)p.work_counter;
                    const uint32_t left = done < pq.n_chunk_items ? pq.n_chunk_items - done : 0u;
                    grant = min((uint32_t)__popc(need_mask), max(min(left / p.tail_div + 1u, 32u), 4u));
                    base = atomicAdd(p.work_counter, grant);
                }
                base = __shfl_sync(FULL, base, leader);
                grant = __shfl_sync(FULL, grant, leader);
                exhausted = base >= pq.n_chunk_items;
                const uint32_t rank = __popc(need_mask & ((1u << lane) - 1u));
                const uint32_t in_chunk = base + rank;
                const uint32_t item = pq.first_item + in_chunk;
                const ItemInfo it = decode_item(p, item);
                if (need && rank < grant && in_chunk < pq.n_chunk_items) {
                    if (it.valid) {
                        pix_x = it.x;
                        pix_y = tile_image_row(p, it.ly);
                        sample_slot = in_chunk;
                        item_index = item;
                        primary_ray(p, pix_x, pix_y, it.s, ro, rd);
                        rw = splat(1.0f);
                        r_ior = AIR_IOR;
                        r_depth = -1;
                        r_refl_child = false;
                        r_primary = true;
                        first_point = last_point = -1;
                        have_item = true;
                        have_ray = true;
                    } else {
                        pq.sample_head[in_chunk] = -1;  // padding item of an edge tile
                    }
                }
            }
        }
        if (!__any_sync(FULL, have_ray)) {
            // Nobody got a ray.  That ends the warp only when the queue is empty: a round may hand out nothing but
            // padding items (the lower half of the last tile row when the height is not a multiple of 4), and a warp
            // that left then would leave the items behind it to the others -- at 1140x950 every warp of the grid
            // left that way before the row was finished (round-2 finding: stale list heads in the last two rows).
            if (exhausted) break;
            continue;
        }

        // ---------------------------------------------------------------- closest hit (per lane)
        float3 d = splat(0.0f), P = splat(0.0f), N = splat(0.0f);
        ClosestHit h;
        h.slot = -1;
        h.t = 0.0f;
        h.meta = 0;
        h.prim = 0;
        if (have_ray) {
            d = normalized3(rd);  // Ray::new_with_mask (ray.rs:52-57)
            if (r_primary) {  // traced by primary_kernel at full width; here only fetched
                const uint4 ph = p.primary_hits[item_index];
                h.t = __uint_as_float(ph.x);
                h.slot = (int32_t)ph.y;
                h.prim = ph.z;
                h.meta = ph.w;
                if (COUNT) cn.closest++;
            } else {
                h = trace_closest<COUNT>(sc, ro, d, cull, cn);
            }
        }
        const bool hit = have_ray && h.slot >= 0;
        // ---------------------------------------------------------------- append the shading point, spawn the children
        const uint32_t hit_mask = __ballot_sync(FULL, hit);
        if (hit_mask) {
            uint32_t base = 0;
            const int leader = __ffs(hit_mask) - 1;
            if ((int)lane == leader) base = atomicAdd(pq.n_points, (unsigned int)__popc(hit_mask));
            base = __shfl_sync(FULL, base, leader);
            if (hit) {
                P = ray_at(ro, d, h.t);
                if (h.meta & rtb::PRIM_SPHERE) {
                    float4 q0 = __ldg(sc.prims + 4 * (size_t)h.slot);
                    N = normalized3(P - f3(q0.x, q0.y, q0.z));
                } else {
                    float4 sn = __ldg(sc.shade_n + h.slot);
                    N = f3(sn.x, sn.y, sn.z);
                }
                const Material m = load_material(sc, h.meta & rtb::PRIM_MATERIAL_MASK);
                const float df = attenuation(h.t);
                // single_raytrace (raytracer_renderer.rs:147-264), linearised: a reflection child's whole subtree is
                // scaled by attenuation(child.t) (raytracer_renderer.rs:722-727)
                const float3 w = r_refl_child ? (rw * df) : rw;
                const uint32_t idx = base + __popc(hit_mask & ((1u << lane) - 1u));
                if (idx < pq.capacity) {  // (beyond: the host sees n_points > capacity and renders the chunk in two halves)
                    PathPoint *pt = pq.points + idx;
                    pt->a = make_float4(P.x, P.y, P.z, h.t);
                    pt->b = make_float4(N.x, N.y, N.z, __uint_as_float(h.meta & rtb::PRIM_MATERIAL_MASK));
                    pt->c = make_float4(d.x, d.y, d.z, __uint_as_float(pix_x | (pix_y << 16)));
                    pt->w = make_float4(w.x, w.y, w.z, __int_as_float(-1));
                    // the previous point of this sample (this lane wrote it) now has a successor
                    if (last_point >= 0) pq.points[last_point].w.w = __int_as_float((int32_t)idx);
                }
                if (first_point < 0) first_point = (int32_t)idx;
                last_point = (int32_t)idx;
                const bool is_transmissive = m.transmissive;
                const bool is_reflective = (m.metallic > 0.0f) || is_transmissive;
                const float start_ior = r_ior;
                if (refl_on && is_reflective) {  // calculate_reflection (raytracer_renderer.rs:526-729)
                    float ct = dot3(d, N);
                    bool inside = ct < 0.0f;
                    float3 inormal = inside ? -N : N;
                    float new_ior = inside ? m.ior : AIR_IOR;
                    float eta = inside ? (new_ior / start_ior) : (start_ior / new_ior);
                    float ci = fabsf(ct);
                    float sin2 = eta * eta * (1.0f - ci * ci);
                    bool tir = sin2 >= 1.0f;
                    bool reflective = (m.metallic > 0.0f) || (is_transmissive && tir);
                    if (reflective) {
                        float3 rdir = normalized3(reflected3(d, N));
                        bool none = (fabsf(rdir.x - 0.0f) <= EPS) && (fabsf(rdir.y - 0.0f) <= EPS) && (fabsf(rdir.z - 0.0f) <= EPS);
                        int next = r_depth < 0 ? (int)p.refl_depth : max(r_depth - 1, 0);
                        if (!none && next > 0 && sp < PATH_STACK) {
                            float3 reflectance, tmp;
                            compute_fresnel(m, inormal, -d, start_ior, reflectance, tmp);
                            float3 co = P + rdir * p.eps_d;
                            float3 cw = w * reflectance;
                            stack[sp].a = make_float4(co.x, co.y, co.z, start_ior);
                            stack[sp].b = make_float4(rdir.x, rdir.y, rdir.z, __uint_as_float((uint32_t)next | PATH_REFL_CHILD));
                            stack[sp].c = make_float4(cw.x, cw.y, cw.z, 0.0f);
                            sp++;
                        }
                    }
                }
                if (refr_on && is_transmissive) {  // calculate_refractions (raytracer_renderer.rs:279-524)
                    float ct = dot3(d, N);
                    bool entering = ct <= 0.0f;
                    float3 inormal = entering ? -N : N;
                    float new_ior = entering ? m.ior : AIR_IOR;
                    float eta = entering ? (new_ior / start_ior) : (start_ior / new_ior);
                    float inv_eta = 1.0f / eta;
                    float3 tmp, transmittance;
                    compute_fresnel(m, inormal, d, inv_eta, tmp, transmittance);
                    float3 rdir = normalized3(refracted3(d, -inormal, inv_eta));
                    float op = m.opacity;
                    int step = (op < 0.5f) ? 2 : 1;
                    int factor = (op <= 0.3f) ? 3 : ((op < 0.5f) ? 2 : 1);
                    int next = r_depth < 0 ? (int)p.refr_depth / factor : max(r_depth - step, 0);
                    if (next > 0 && sp < PATH_STACK) {
                        float boost = m.boost + 1.0f;
                        float3 co = P + rdir * p.eps_d;
                        float3 cw = w * (transmittance * boost);
                        stack[sp].a = make_float4(co.x, co.y, co.z, new_ior);
                        stack[sp].b = make_float4(rdir.x, rdir.y, rdir.z, __uint_as_float((uint32_t)next));
                        stack[sp].c = make_float4(cw.x, cw.y, cw.z, 0.0f);
                        sp++;
                    }
                }
            }
        }
        have_ray = false;
    }
    if (COUNT) {
        atomicAdd(p.counters + 0, cn.closest);
        atomicAdd(p.counters + 2, cn.box);
        atomicAdd(p.counters + 3, cn.tri);
        atomicAdd(p.counters + 4, cn.sph);
    }
}

// One thread per primary sample of the chunk: the sample's colour = its points' contributions added in the order
// the path walk met them (render_kernel's `acc = acc + w * local`).  The list is linked forwards (paths_kernel gives
// every point its successor when that one is appended), so it is one walk with one memory latency per point.
// (Before: backward links, reversed in place and then summed -- two dependent walks with a store in between; 0.98 -> 0.11 ms on C4.)
__global__ void gather_paths_kernel(const RenderParams p) {
    const uint32_t in_chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const PathQueue &pq = p.pq;
    if (in_chunk >= pq.n_chunk_items) return;
    const ItemInfo it = decode_item(p, pq.first_item + in_chunk);
    if (!it.valid) return;
    const uint32_t sample_index = (it.ly * p.width + it.x) * p.S + it.s;
    int32_t cur = pq.sample_head[in_chunk];
    float3 acc = splat(0.0f);
    const bool valid = cur >= 0;  // the first point is the primary hit
    while (cur >= 0) {  // path order: one latency per point (the contribution and the link are independent loads)
        const float4 c = pq.contrib[cur];
        const int32_t next = __float_as_int(pq.points[cur].w.w);
        acc = acc + f3(c.x, c.y, c.z);
        cur = next;
    }
    p.samples[sample_index] = make_float4(acc.x, acc.y, acc.z, valid ? 1.0f : 0.0f);
}

// Closest hit of every primary sample, one thread per work item (same item order as render_kernel).
// Splitting this off keeps the divergent closest-hit traversal at full warp width and lets
// render_kernel fetch work in small chunks (fine-grained load balance) at no cost.
template <bool COUNT>
__global__ void __launch_bounds__(128) primary_kernel(const RenderParams p) {
    stage_top_nodes(p.sc);
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    const ItemInfo it = decode_item(p, item);
    if (!it.valid) return;
    Counters cn = {0, 0, 0, 0, 0, 0};
    float3 o, dir;
    primary_ray(p, it.x, tile_image_row(p, it.ly), it.s, o, dir);
    const ClosestHit h = trace_closest<COUNT>(p.sc, o, normalized3(dir), (p.flags & RT_FLAG_BACKFACE_CULLING) != 0, cn);
    p.primary_hits[item] = make_uint4(__float_as_uint(h.t), (uint32_t)h.slot, h.prim, h.meta);
    if (COUNT) {  // rays are counted where they are consumed (render_kernel); tests are counted here
        atomicAdd(p.counters + 2, cn.box);
        atomicAdd(p.counters + 3, cn.tri);
        atomicAdd(p.counters + 4, cn.sph);
        atomicAdd(p.counters + 8, cn.box);  // ... and once more as the pre-pass's own share
        atomicAdd(p.counters + 9, cn.tri);
        atomicAdd(p.counters + 10, cn.sph);
    }
}

// Direct lighting without bounces (no `reflections` / `refractions` feature): every sample is one
// primary hit + its L*M shadow rays, so the frame is processed as a stream of small, uniform work
// units instead of per-sample trees.  Unit = (batch of 32 neighbouring samples, light group), a
// group being the M soft-shadow copies of one light (or all lights when M == 1).  Persistent warps
// fetch units in light-major order with one atomicAdd each; each lane shades its own sample against
// the group's light copies IN ORDER, continuing the running per-sample sums left by the unit of
// the previous group -- the reference's `for light in lights` order.  Units of one batch are
// chained through a tagged 2 x 16-byte record in global memory (the tag travels inside each
// 16-byte store, so no fence -- and no L1 invalidation -- is needed); the predecessor was handed
// out a whole group earlier, so the consumer practically never waits.  19 shadow batches per unit
// instead of 95 and a single tail per frame is what keeps 8-GPU strong scaling efficient.
//
// Record storage is BOUNDED: the batches are processed in windows of `chain_window` batches (all light groups of
// a window before the next window starts), and a window keeps two generations of records, indexed by the parity
// of the light group -- the writer of (group g, batch b) has waited for group g-1, which read group g-2 before it
// wrote, so the slot of g-2 is free.  Across windows the slot's last reader is found through the chain as well:
// the LAST group's unit of (window w, batch b) publishes a tag in chain_done[b] (st.release), and the FIRST
// group's unit of (window w+1, batch b) waits for it (ld.acquire) before it overwrites the slot; in practice it
// never waits (that unit finished a whole window ago).  32 MB of records for any frame size instead of
// S*W*H*(L-1)*32 B (1 GB for the headline frame), and the records are consumed while still in L2.
// (The DEFER variant -- small launches only -- keeps one window and every generation: its deferred units are
// finished at the end of the launch and need their predecessors' records until then.)
struct DirectState {  // running light sums of one sample: (sum.rgb, bits tag) twice
    float4 d, s;
};
// gpu-scope relaxed 16-byte accesses for the chained records (volatile asm: the spin loop must
// really re-load, and the compiler may not assume the memory is unchanged)
RT_DEV float4 ld_relaxed_gpu(const float4 *ptr) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr) : "memory");
    return v;
}
RT_DEV void st_relaxed_gpu(float4 *ptr, float4 v) {
    asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

RT_DEV unsigned int ld_acquire_gpu_u32(const unsigned int *ptr) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
RT_DEV void st_release_gpu_u32(unsigned int *ptr, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

#ifdef RT_DBG_TAIL  // tuning experiments only: per-warp timeline of direct_kernel (rt_dbg_tail in rt_api.cu)
__device__ unsigned long long g_dbg_tail[12 * 8192];
RT_DEV unsigned long long dbg_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

// ---- very expensive units ---------------------------------------------------------------------
// A few units cost 20-30x the average: their samples sit where the M shadow rays of one light run
// lengthwise through dense geometry, the beams overflow and every ray walks the BVH itself, while the
// other lanes of the warp wait (SIMT).  Handed out late, one such unit is the tail of the whole launch
// -- and with a rank's share of a multi-GPU frame every unit is "late".  They are therefore DEFERRED:
// the unit is put on a list (and marked in the chain, so that the units that continue its samples'
// sums defer themselves too), and once the unit queue has run dry all warps work the list off together:
// one sub-task = one light copy for the 32 samples of a unit.  Contributions go to a scratch slab; the
// warp that completes a unit's last sub-task adds them to the predecessor's sums in copy order, i.e.
// in the reference's light order: the image is bit-identical to processing the unit in place.
constexpr uint32_t HEAVY_CAP = 16384;             // most deferred units per launch (RenderParams::heavy_cap; beyond: processed in place)
constexpr uint32_t CHAIN_DEFERRED = 0x80000000u;  // flag in a chain tag: the unit is on the heavy list
struct HeavyQueue {
    unsigned int n_heavy;       // list slots handed out (a unit reserves its own slot and those of its successors at once)
    unsigned int warps_done_a;  // warps that have left the unit loop (after that the list is final)
    unsigned int next_task;     // next sub-task ticket: position in `ready` * M + copy   } read together (8-byte aligned)
    unsigned int n_ready;       // units whose slot has been published, in publication order }
    unsigned int unit_counter;  // direct_kernel's unit queue (here so that one memset resets everything)
    unsigned int pad[3];
    unsigned int done[HEAVY_CAP];   // finished sub-tasks per slot
    unsigned int unit[HEAVY_CAP];   // per slot: unit + 1; 0 until written
    unsigned int ready[HEAVY_CAP];  // publication order of the units that are worked off as sub-tasks: slot + 1; 0 until written
    unsigned int claim[HEAVY_CAP];  // per slot: 1 once a warp has taken over the slot's final sums (exactly one does)
};
RT_DEV unsigned int ld_volatile_u32(const unsigned int *p) { return *(const volatile unsigned int *)p; }
// Every wait loop of direct_kernel is bounded: a protocol error must surface as an error code (rt_render /
// rt_check_frame fail with RT_ERR_CUDA), never as a hung GPU.  The watchdog word is STICKY: it is only reset by
// the host after it has reported the error, so an asynchronous caller (rt_render_device without stats) that
// checks late still sees a trip of any frame since its last check.
constexpr uint32_t WAIT_SPINS = 1u << 23;   // plain polling loops (seconds)
constexpr uint32_t WAIT_SLEEPS = 1u << 21;  // loops that sleep 1 us per poll (seconds)
RT_DEV void watchdog_trip(unsigned int *watchdog, unsigned int code) {
    if (watchdog) atomicMax(watchdog, code);
}

// one sample of a batch, rebuilt from the primary pre-pass record (cheap next to its shadow rays)
struct SampleRec {
    bool valid, hit, transmissive;
    float3 P, N, d, mcolor;
    float mshininess, t;
    uint32_t sample_index, pix_x, pix_y;
};
// POINTS: `item` is a shading point of the path queue instead of a primary sample
template <bool POINTS>
RT_DEV SampleRec rebuild_sample(const RenderParams &p, uint32_t item) {
    const DScene &sc = p.sc;
    SampleRec r;
    if (POINTS) {
        r.valid = r.hit = item < p.n_items;
        r.sample_index = item;
        r.P = r.N = r.d = r.mcolor = splat(0.0f);
        r.mshininess = r.t = 0.0f;
        r.transmissive = false;
        r.pix_x = r.pix_y = 0;
        if (r.hit) {
            const PathPoint *pt = p.pq.points + item;
            const float4 a = __ldg(&pt->a), b = __ldg(&pt->b), c = __ldg(&pt->c);
            r.P = f3(a.x, a.y, a.z);
            r.t = a.w;
            r.N = f3(b.x, b.y, b.z);
            r.d = f3(c.x, c.y, c.z);
            const uint32_t pix = __float_as_uint(c.w);
            r.pix_x = pix & 0xFFFFu;
            r.pix_y = pix >> 16;
            const Material m = load_material(sc, __float_as_uint(b.w));
            r.mcolor = m.color;
            r.mshininess = m.shininess;
            r.transmissive = m.transmissive;
        }
        return r;
    }
    const ItemInfo it = decode_item(p, item);
    uint4 ph = make_uint4(0u, 0xFFFFFFFFu, 0u, 0u);
    if (it.valid) ph = p.primary_hits[item];
    r.valid = it.valid;
    r.hit = it.valid && (int32_t)ph.y >= 0;
    r.sample_index = (it.ly * p.width + it.x) * p.S + it.s;
    r.pix_x = it.x;
    r.pix_y = tile_image_row(p, it.ly);
    r.P = r.N = r.d = r.mcolor = splat(0.0f);
    r.mshininess = r.t = 0.0f;
    r.transmissive = false;
    if (r.hit) {
        float3 o, dir;
        primary_ray(p, r.pix_x, r.pix_y, it.s, o, dir);
        r.d = normalized3(dir);
        r.t = __uint_as_float(ph.x);
        const int32_t slot = (int32_t)ph.y;
        const uint32_t meta = ph.w;
        r.P = ray_at(o, r.d, r.t);
        if (meta & rtb::PRIM_SPHERE) {
            const float4 q0 = __ldg(sc.prims + 4 * (size_t)slot);
            r.N = normalized3(r.P - f3(q0.x, q0.y, q0.z));
        } else {
            const float4 sn = __ldg(sc.shade_n + slot);
            r.N = f3(sn.x, sn.y, sn.z);
        }
        const Material m = load_material(sc, meta & rtb::PRIM_MATERIAL_MASK);
        r.mcolor = m.color;
        r.mshininess = m.shininess;
        r.transmissive = m.transmissive;
    }
    return r;
}

// single_raytrace without bounces (raytracer_renderer.rs:195-257): the last light group publishes the sample
// POINTS: the point's contribution to its sample, weight * local colour (the `acc + w * local` term of the path walk)
template <bool POINTS>
RT_DEV void store_direct_sample(const RenderParams &p, const SampleRec &r, float3 sum_d, float3 sum_s) {
    const float3 ambient = (r.mcolor * splat(1.0f)) * 0.08f;
    const float df = attenuation(r.t);
    const float3 direct = (ambient + sum_d) * df;
    const float3 specular = sum_s * df;
    const float3 local = r.transmissive ? specular : (direct + specular);
    if (POINTS) {
        const float4 w = __ldg(&p.pq.points[r.sample_index].w);
        const float3 c = f3(w.x, w.y, w.z) * local;
        p.pq.contrib[r.sample_index] = make_float4(c.x, c.y, c.z, 1.0f);
    } else {
        p.samples[r.sample_index] = make_float4(local.x, local.y, local.z, 1.0f);
    }
}

// The final sums of a listed unit (slot idx = light group g of batch b): the predecessor's sums continued with the unit's
// M per-copy contributions from the scratch slab, in copy order -- the reference's light order -- then the unit's own
// record (or, for the last group, the sample).  NEVER blocks: if the predecessor's record is not final yet, the call
// returns, and the warp that finishes the predecessor carries on into this slot (it looks at done[idx + 1] after it has
// written its record; this call looked at the record after done[idx] was complete: with a fence on both sides at least
// one of the two sees the other).  `claim` makes sure exactly one warp does a slot.  After a slot the same warp goes on to
// the next group's slot if that one's contributions are complete -- a chain of listed units is finished by whoever
// completes its last missing piece, at once, instead of by waiting warps.
template <bool POINTS>
__device__ __noinline__ void finish_chain(const RenderParams *gp, HeavyQueue *hq, float4 *hscratch, DirectState *state, uint32_t n_groups, uint32_t frame_tag,
                                          uint32_t idx, uint32_t g, uint32_t b) {
    const RenderParams &p = *gp;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t n_batches = (p.n_items + 31u) / 32u;
    const size_t chain_stride = (size_t)n_batches * 32u;
    const uint32_t item = b * 32u + lane;
    const SampleRec r = rebuild_sample<POINTS>(p, item);
    float3 sum_d = splat(0.0f), sum_s = splat(0.0f);
    bool ready = true;
    if (g > 0 && r.hit) {
        const uint32_t want = frame_tag + g - 1u;
        const float4 sd = ld_relaxed_gpu(&state[(size_t)(g - 1u) * chain_stride + item].d), ss = ld_relaxed_gpu(&state[(size_t)(g - 1u) * chain_stride + item].s);
        ready = __float_as_uint(sd.w) == want && __float_as_uint(ss.w) == want;
        sum_d = f3(sd.x, sd.y, sd.z);
        sum_s = f3(ss.x, ss.y, ss.z);
    }
    if (!__all_sync(FULL, ready)) return;  // the predecessor's finisher will come here
    unsigned int won = 0;
    if (lane == 0) won = atomicCAS(&hq->claim[idx], 0u, 1u) == 0u;
    if (!__shfl_sync(FULL, won, 0)) return;
    for (;;) {
        if (r.hit) {
            const float4 *slab = hscratch + ((size_t)idx * p.M * 32u + lane) * 2u;  // + copy * 64
            for (uint32_t k = 0; k < p.M; k++) {
                const float4 a = ld_relaxed_gpu(slab + (size_t)k * 64u), b2 = ld_relaxed_gpu(slab + (size_t)k * 64u + 1);
                sum_d = sum_d + f3(a.x, a.y, a.z);
                sum_s = sum_s + f3(b2.x, b2.y, b2.z);
            }
            if (g + 1 != n_groups) {
                const float tag = __uint_as_float(frame_tag + g);
                st_relaxed_gpu(&state[(size_t)g * chain_stride + item].d, make_float4(sum_d.x, sum_d.y, sum_d.z, tag));
                st_relaxed_gpu(&state[(size_t)g * chain_stride + item].s, make_float4(sum_s.x, sum_s.y, sum_s.z, tag));
            } else {
                store_direct_sample<POINTS>(p, r, sum_d, sum_s);
            }
        }
        if (g + 1 == n_groups) return;
        // the next group's slot was reserved together with this one: are its contributions complete?
        __threadfence();
        __syncwarp();
        won = 0;
        if (lane == 0) won = ld_volatile_u32(&hq->done[idx + 1u]) == p.M && atomicCAS(&hq->claim[idx + 1u], 0u, 1u) == 0u;
        if (!__shfl_sync(FULL, won, 0)) return;
        __threadfence();
        idx++;
        g++;
    }
}

// Phase B of direct_kernel (see above).  Out of line, parameters read from their global-memory copy, so
// that it does not weigh on the register allocation of the unit loop.
template <bool COUNT, bool POINTS>
__device__ __noinline__ void finish_heavy_units(const RenderParams *gp, HeavyQueue *hq, float4 *hscratch, DirectState *state, uint32_t n_groups,
                                                uint32_t frame_tag, uint32_t total_warps, Counters *cn_out, unsigned int *watchdog) {
    const RenderParams &p = *gp;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t FULL = 0xFFFFFFFFu;
    const bool cull = (p.flags & RT_FLAG_BACKFACE_CULLING) != 0;
    const bool packet_noaa = (p.flags & RT_FLAG_PACKET_SEMANTICS) && !(p.flags & RT_FLAG_ANTI_ALIASING);
    const uint32_t n_batches = (p.n_items + 31u) / 32u;
    const size_t chain_stride = (size_t)n_batches * 32u;
    Counters cn = {0, 0, 0, 0, 0, 0};
    // The unit queue has run dry for this warp: it works sub-tasks off until the list is final (every warp has left the
    // unit loop) and empty.  Tickets come from one atomicAdd each and follow the PUBLICATION order of the units
    // (`ready[]`), so a ticket never refers to a unit that is listed later.  (Measured and dropped, profiles/README.md:
    // warps lending a hand between two units as soon as a unit is listed -- +2 % on a rank's eighth of the headline frame.)
    if (lane == 0) {
        __threadfence();  // this warp's publications before its "left the unit loop"
        atomicAdd(&hq->warps_done_a, 1u);
    }
    for (;;) {
        uint32_t idx = 0xFFFFFFFFu, c = 0, unit = 0;
        if (lane == 0) {
            const unsigned int ticket = atomicAdd(&hq->next_task, 1u);
            const uint32_t pos = ticket / p.M;
            for (uint32_t spins = 0;; spins++) {
                if (spins > WAIT_SLEEPS) {
                    watchdog_trip(watchdog, 1u);
                    break;
                }
                const bool final_list = ld_volatile_u32(&hq->warps_done_a) >= total_warps;
                if (final_list) __threadfence();  // (the flag first, then the counter: a final list has all its entries counted)
                if (pos < ld_volatile_u32(&hq->n_ready)) {
                    c = ticket - pos * p.M;
                    uint32_t w = 0;
                    while ((idx = ld_volatile_u32(&hq->ready[pos]) - 1u) == 0xFFFFFFFFu)  // (counted, being written)
                        if (++w > WAIT_SPINS) {
                            watchdog_trip(watchdog, 2u);
                            break;
                        }
                    if (idx != 0xFFFFFFFFu) unit = ld_volatile_u32(&hq->unit[idx]) - 1u;  // (written before the slot was counted)
                    break;
                }
                if (final_list) break;  // the launch has no such sub-task
                __nanosleep(1000);
            }
        }
        idx = __shfl_sync(FULL, idx, 0);
        if (idx == 0xFFFFFFFFu) break;
        c = __shfl_sync(FULL, c, 0);
        unit = __shfl_sync(FULL, unit, 0);
        const uint32_t g = unit / n_batches, b = unit - g * n_batches;
        const uint32_t item = b * 32u + lane;
        const SampleRec r = rebuild_sample<POINTS>(p, item);
        float4 *slab = hscratch + ((size_t)idx * p.M * 32u + lane) * 2u;  // + copy * 64
        if (r.hit) {
            const LightCopy L = load_light(p, g);
            const uint32_t cloud = cloud_index(p.seed, packet_noaa ? r.pix_x / 8u : r.pix_x, r.pix_y, g, p.K);
            float3 dc, spc;
            shade_pair<COUNT, 2>(p, r.P, r.N, r.d, r.mcolor, r.mshininess, copy_position(p, L.pos, cloud, c), L.intensity, L.color, cull, cn, dc, spc);
            st_relaxed_gpu(slab + (size_t)c * 64u, make_float4(dc.x, dc.y, dc.z, 0.0f));
            st_relaxed_gpu(slab + (size_t)c * 64u + 1, make_float4(spc.x, spc.y, spc.z, 0.0f));
        }
        __threadfence();
        __syncwarp();
        unsigned int finished = 0;
        if (lane == 0) finished = atomicAdd(&hq->done[idx], 1u) + 1u;
        finished = __shfl_sync(FULL, finished, 0);
        if (finished != p.M) continue;
        // last sub-task of the unit: its sums can be finished as soon as the predecessor's are (finish_chain)
        __threadfence();
        finish_chain<POINTS>(gp, hq, hscratch, state, n_groups, frame_tag, idx, g, b);
    }
    if (COUNT && cn_out) {
        cn_out->closest += cn.closest; cn_out->shadow += cn.shadow; cn_out->box += cn.box;
        cn_out->tri += cn.tri; cn_out->sph += cn.sph; cn_out->beam_prim += cn.beam_prim;
    }
}

#ifndef RT_DIRECT_MIN_CTAS
#define RT_DIRECT_MIN_CTAS RT_MIN_CTAS_PER_SM
#endif
// puts `unit` on the heavy list (slot idx): marks its chain records (the mark carries the slot: the unit that
// continues these samples' sums defers itself too, into the slot after this one, which was reserved together with
// it), zeroes the samples of lanes that missed (last group), publishes
__device__ __noinline__ void defer_unit(HeavyQueue *hq, DirectState *own_rec, float4 *sample, uint32_t unit, uint32_t idx, uint32_t deferred_tag,
                                        bool mark_chain, bool zero_sample) {
    if (mark_chain) {
        const float tag = __uint_as_float(deferred_tag);
        st_relaxed_gpu(&own_rec->d, make_float4(__uint_as_float(idx), 0.0f, 0.0f, tag));
        st_relaxed_gpu(&own_rec->s, make_float4(0.0f, 0.0f, 0.0f, tag));
    }
    if (zero_sample) *sample = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __threadfence();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        *(volatile unsigned int *)&hq->unit[idx] = unit + 1u;
        __threadfence();
        const unsigned int r = atomicAdd(&hq->n_ready, 1u);  // sub-task tickets follow the publication order
        *(volatile unsigned int *)&hq->ready[r] = idx + 1u;
    }
}

// DEFER: with the deferral of very expensive units compiled in.  It costs a few per cent of steady-state
// throughput (register pressure in the unit loop) and removes the tail of the launch, so the host selects
// it when a launch has few units per warp -- a rank's share of a multi-GPU frame.
#ifndef RT_WAIT_BACKOFF_NS
#define RT_WAIT_BACKOFF_NS 200  // sleep between two polls of a chained record that is not there yet (0: plain polling)
#endif
#ifndef RT_DEFER_MIN_CTAS
#define RT_DEFER_MIN_CTAS 7  // (B200, units of all lights: a rank's eighth of C3 takes 1.38 ms with 5, 6 or 7 CTAs per SM; the full frame 10.11 / 10.02 / 9.78 ms)
#endif
template <bool COUNT, bool BEAM, bool DEFER, bool POINTS = false>
__global__ void __launch_bounds__(RT_CTA_THREADS, DEFER ? RT_DEFER_MIN_CTAS : RT_DIRECT_MIN_CTAS) direct_kernel(const RenderParams p, uint32_t n_groups, uint32_t frame_tag,
                                                                                    unsigned int *counter, DirectState *state, HeavyQueue *hq,
                                                                                    float4 *hscratch, const RenderParams *gp, unsigned int *watchdog) {
    stage_top_nodes(p.sc);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t FULL = 0xFFFFFFFFu;
    const DScene &sc = p.sc;
    const bool cull = (p.flags & RT_FLAG_BACKFACE_CULLING) != 0;
    const bool packet_noaa = (p.flags & RT_FLAG_PACKET_SEMANTICS) && !(p.flags & RT_FLAG_ANTI_ALIASING);
    const uint32_t n_batches = (p.n_items + 31u) / 32u;
    const uint32_t n_units = n_batches * n_groups;
    const uint32_t window = DEFER ? n_batches : p.chain_window;  // batches per window (>= 1)
    const uint32_t lpu = p.M > 1 ? p.lights_per_unit : 1u;  // lights a unit takes in a row (soft shadows)
    // DEFER keeps its records, tags and list entries PER LIGHT whatever the unit size: a unit that defers at its light g
    // becomes, from there on, the per-light units g, g+1, .. of its batch (below)
    const uint32_t n_lights_pl = p.M > 1 ? sc.n_lights : 1u;
    const uint32_t window_units = window * n_groups;
    const size_t chain_stride = (size_t)window * 32u;  // records of one generation
    const bool deferring = DEFER && BEAM && hq != nullptr;
    Counters cn = {0, 0, 0, 0, 0, 0};
#ifdef RT_DBG_TAIL
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long dbg_t0 = dbg_globaltimer(), dbg_spin = 0, dbg_last_start = 0, dbg_worst = 0, dbg_worst_info = 0, dbg_cur_info = 0;
    unsigned long long dbg_ph0 = 0, dbg_ph1 = 0, dbg_ph2 = 0, dbg_help_time = 0;
    uint32_t dbg_units = 0, dbg_helped = 0;
#endif
    for (;;) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(counter, 1u);
        unit = __shfl_sync(FULL, unit, 0);
        if (unit >= n_units) break;
#ifdef RT_DBG_TAIL
        dbg_units++;
        dbg_last_start = dbg_globaltimer();
        dbg_cur_info = (unsigned long long)unit << 32;
#endif
        // unit -> (window, light group, batch): all groups of a window's batches before the next window
        const uint32_t win = unit / window_units, in_win = unit - win * window_units;
        const uint32_t win_batches = min(window, n_batches - win * window);  // (the last window may be short)
        const uint32_t group_index = in_win / win_batches, b_in_win = in_win - group_index * win_batches;
        const uint32_t batch = win * window + b_in_win;
        const bool last = group_index + 1 == n_groups;
        const uint32_t item = batch * 32u + lane;
        const SampleRec r = rebuild_sample<POINTS>(p, item);
        // generation of a record: the parity of the group, or (DEFER: one window, every generation kept) the LIGHT the
        // sums stand after -- the unit's last light for its own record, the light before its first one for the predecessor's
        const uint32_t first_light = group_index * lpu, last_light = min(first_light + lpu, n_lights_pl) - 1u;  // (DEFER)
        const uint32_t gen = DEFER ? last_light : (group_index & 1u), prev_gen = DEFER ? first_light - 1u : (gen ^ 1u);
        DirectState *const own_rec = state + (size_t)gen * chain_stride + (size_t)b_in_win * 32u + lane;
        const DirectState *const prev_rec = state + (size_t)prev_gen * chain_stride + (size_t)b_in_win * 32u + lane;  // (only dereferenced when group_index > 0)
        const uint32_t tag_base = frame_tag + win * n_groups;  // tags are unique per (frame, window, group); DEFER: frame_tag + light
        if (!DEFER && win > 0 && (group_index == 0 || last) && n_groups > 1) {
            // The FIRST unit of a batch column in a window: its records reuse the slots of the previous window -- wait
            // until that window's last unit of this column, and with it (through the chain) every reader of those slots,
            // has finished.  The LAST unit waits for the same tag before it publishes its own: lanes without a hit take no
            // part in the record chain, so a last unit none of whose lanes hit would otherwise publish at once -- ahead
            // of the previous window's tag, which the first unit of its own window may still be waiting for.  Tags only
            // grow (per window, per frame), so "reached" is a signed comparison.
            const unsigned int want_done = tag_base - 1u;  // = tag of (win - 1, last group)
            if (lane == 0) {
                for (uint32_t spins = 0; (int32_t)(ld_acquire_gpu_u32(p.chain_done + b_in_win) - want_done) < 0; spins++) {
                    if (spins > WAIT_SPINS) {
                        watchdog_trip(watchdog, 6u);
                        break;
                    }
                    __nanosleep(200);
                }
            }
            __syncwarp();
        }
        if (COUNT && !POINTS && r.valid && group_index == 0) cn.closest++;  // (POINTS: counted where the ray was traced)
        // lanes of a batch that missed sit the batch out but stay with the warp
        float3 sum_d = splat(0.0f), sum_s = splat(0.0f);
        const uint32_t n_in_group = p.M > 1 ? p.M : p.LM;
        uint32_t cand[BEAM_MAX_CAND];  // shadow beam: one BVH walk for the M copies of this light (rt_device.cuh)
        uint32_t inherited_slot = 0xFFFFFFFFu;  // the predecessor is on the heavy list: the slot that was reserved for this unit
        bool unit_parked = false;               // (DEFER) the unit went to the heavy list or parked its contributions: nothing to store here
        if (r.hit) {
            if (group_index > 0) {
                // wait for (this sample, previous group): gpu-scope loads see the producer's stores in L2
                const uint32_t want = DEFER ? frame_tag + first_light - 1u : tag_base + group_index - 1u;
                float4 sd, ss;
#ifdef RT_DBG_TAIL
                const unsigned long long dbg_s0 = dbg_globaltimer();
#endif
                for (uint32_t spins = 0;; spins++) {
                    if (spins > WAIT_SPINS) {
                        watchdog_trip(watchdog, 4u);
                        break;
                    }
                    if (RT_WAIT_BACKOFF_NS > 0 && spins > 4u) __nanosleep(RT_WAIT_BACKOFF_NS);  // (a waiting warp must not take issue slots from the working ones)
                    sd = ld_relaxed_gpu(&prev_rec->d);
                    ss = ld_relaxed_gpu(&prev_rec->s);
                    const uint32_t td = __float_as_uint(sd.w), ts = __float_as_uint(ss.w);
                    if (td == want && ts == want) break;
                    if (DEFER && td == (want | CHAIN_DEFERRED)) {  // the predecessor is on the heavy list: so is this unit
                        inherited_slot = __float_as_uint(sd.x) + 1u;
                        break;
                    }
                }
#ifdef RT_DBG_TAIL
                dbg_spin += dbg_globaltimer() - dbg_s0;
#endif
                sum_d = f3(sd.x, sd.y, sd.z);
                sum_s = f3(ss.x, ss.y, ss.z);
            }
        }
        // the unit's lights, in the reference's order (soft shadows: `lpu` lights of M copies each; hard shadows: one
        // group of all lights).  The running sums stay in registers from one light to the next.
        for (uint32_t gl = 0; gl < lpu; gl++) {
        const uint32_t li0 = p.M > 1 ? group_index * lpu + gl : 0u;
        if (gl > 0 && li0 >= sc.n_lights) break;
        uint32_t cloud = 0u;
        int n_cand = -1;
        int32_t last_occluder = -1;
        bool heavy = false;  // this lane's beam overflowed: M separate BVH walks, everybody else in the warp waiting
        LightCopy L = {splat(0.0f), splat(0.0f), 0.0f};
        if (p.LM > 0) L = load_light(p, li0);
        if (r.hit) {
            if (p.M > 1) cloud = cloud_index(p.seed, packet_noaa ? r.pix_x / 8u : r.pix_x, r.pix_y, li0, p.K);
            if (BEAM) {
                const float3 w2s = f3(p.w2s_w, p.w2s_h, p.w2s_d);
                // the same float operations as the copies' `lpos + rp * w2s`, which are monotonic in rp
                const float3 lmin = L.pos + f3(p.cloud_lo[0], p.cloud_lo[1], p.cloud_lo[2]) * w2s;
                const float3 lmax = L.pos + f3(p.cloud_hi[0], p.cloud_hi[1], p.cloud_hi[2]) * w2s;
                const float pad = fmaf(0.25f, sc.box_pad, 2.0f * p.eps_d);
                const float3 C = (lmin + lmax) * 0.5f, hh = (lmax - lmin) * 0.5f + splat(1e-6f * (fabsf(C.x) + fabsf(C.y) + fabsf(C.z)));
                if (RT_FACING_FIRST && light_box_behind(r.N, C - r.P, hh)) {
                    n_cand = BEAM_COVERED;  // like a covered beam: M queries nobody needs, nothing to add (rt_device.cuh)
                } else {
                    const Beam beam = make_beam(r.P, lmin, lmax, pad);
                    n_cand = beam_collect<COUNT>(sc, beam, r.P, C - r.P, hh, cand, cn, cull ? -1.0f : p.eps_d);
                }
                if (DEFER) heavy = n_cand == -1 || n_cand > (int)p.defer_cand;  // (a long list is M long walks too)
            }
            if (DEFER && p.defer_all) heavy = true;  // (tests)
        }
        if (DEFER && BEAM && deferring) {
            __syncwarp();
            const unsigned heavy_lanes = __ballot_sync(FULL, heavy), inherit = __ballot_sync(FULL, inherited_slot != 0xFFFFFFFFu);
            if (heavy_lanes | inherit) {
                // A unit never waits for a listed predecessor: either it inherits the slot that was reserved for it when
                // the first light of its chain was listed, or (fresh: one of its beams is a heavy one) it reserves slots
                // for this light AND every later light of its batch at once -- all or nothing, so the list never holds a
                // light whose successors cannot follow it.  No room: processed in place (its predecessor is then not on
                // the list: nothing to wait for).  From the light it defers at, a unit of several lights goes on as the
                // per-light units of its batch would: every later light inherits the next slot.
                const bool last_l = li0 + 1u == n_lights_pl;
                DirectState *const rec_l = state + (size_t)li0 * chain_stride + (size_t)b_in_win * 32u + lane;  // (DEFER: generation = light)
                unsigned int idx = 0xFFFFFFFFu;
                if (inherit) {
                    idx = __shfl_sync(FULL, inherited_slot, __ffs(inherit) - 1);
                } else {
                    if (lane == 0) {
                        const unsigned int k = n_lights_pl - li0;
                        unsigned int old = ld_volatile_u32(&hq->n_heavy);
                        for (;;) {
                            if (old + k > p.heavy_cap) break;
                            const unsigned int seen = atomicCAS(&hq->n_heavy, old, old + k);
                            if (seen == old) {
                                idx = old;
                                break;
                            }
                            old = seen;
                        }
                    }
                    idx = __shfl_sync(FULL, idx, 0);
                    if (idx != 0xFFFFFFFFu && gl > 0 && r.hit) {
                        // deferred in the middle of the unit: the sums so far are the record the list continues from
                        const float tag = __uint_as_float(frame_tag + li0 - 1u);
                        DirectState *const before = rec_l - chain_stride;
                        st_relaxed_gpu(&before->d, make_float4(sum_d.x, sum_d.y, sum_d.z, tag));
                        st_relaxed_gpu(&before->s, make_float4(sum_s.x, sum_s.y, sum_s.z, tag));
                    }
                }
                if (idx != 0xFFFFFFFFu && heavy_lanes) {
                    // very expensive: worked off as M sub-tasks by all warps once the unit queue is dry
                    defer_unit(hq, rec_l, p.samples + r.sample_index, li0 * n_batches + batch, idx, (frame_tag + li0) | CHAIN_DEFERRED, !last_l && r.hit,
                               last_l && r.valid && !r.hit);
#ifdef RT_DBG_TAIL
                    dbg_helped++;
#endif
                    unit_parked = true;
                    inherited_slot = idx + 1u;  // (the unit's next light, if any)
                    continue;
                }
                if (idx != 0xFFFFFFFFu) {
                    // Only its predecessor is listed, the light itself is an ordinary one: its M contributions are computed
                    // here and now, through the beams like any other's, and parked in the slot's scratch slab; what
                    // waits for the predecessor is just the M ordered additions (finish_chain).  (Before: all M x 32 rays of
                    // such a unit were traced again by BVH walks as sub-tasks -- most of the work of the final phase.)
                    if (!last_l && r.hit) {  // successors must not wait for the record either
                        const float tag = __uint_as_float((frame_tag + li0) | CHAIN_DEFERRED);
                        st_relaxed_gpu(&rec_l->d, make_float4(__uint_as_float(idx), 0.0f, 0.0f, tag));
                        st_relaxed_gpu(&rec_l->s, make_float4(0.0f, 0.0f, 0.0f, tag));
                    }
                    if (last_l && r.valid && !r.hit) p.samples[r.sample_index] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (r.hit) {
                        float4 *slab = hscratch + ((size_t)idx * p.M * 32u + lane) * 2u;  // + copy * 64
                        if (COUNT && n_cand == BEAM_COVERED) cn.shadow += n_in_group;
                        for (uint32_t l = 0; l < n_in_group; l++) {
                            float3 dc = splat(0.0f), spc = splat(0.0f);
                            if (n_cand != BEAM_COVERED)
                                shade_pair<COUNT, 1>(p, r.P, r.N, r.d, r.mcolor, r.mshininess, copy_position(p, L.pos, cloud, l), L.intensity, L.color, cull, cn, dc,
                                                     spc, cand, n_cand, &last_occluder);
                            st_relaxed_gpu(slab + (size_t)l * 64u, make_float4(dc.x, dc.y, dc.z, 0.0f));
                            st_relaxed_gpu(slab + (size_t)l * 64u + 1, make_float4(spc.x, spc.y, spc.z, 0.0f));
                        }
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) atomicAdd(&hq->done[idx], p.M);  // all M contributions at once
                    __threadfence();
                    finish_chain<POINTS>(gp, hq, hscratch, state, n_lights_pl, frame_tag, idx, li0, batch);
                    unit_parked = true;
                    inherited_slot = idx + 1u;
                    continue;
                }
            }
        }
#ifdef RT_DBG_TAIL
        const unsigned long long dbg_tb = dbg_globaltimer();
        {
            int mx = n_cand < 0 ? 255 : n_cand;
            for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
            dbg_cur_info |= (unsigned long long)(mx & 0xFF);
        }
#endif
        // every lane shades its own sample against the group's light copies, in order (a beam that one opaque
        // triangle occludes entirely is M occluded queries: nothing to trace, nothing to add)
        if (COUNT && r.hit && n_cand == BEAM_COVERED) cn.shadow += n_in_group;
        if (r.hit && n_cand != BEAM_COVERED) {
            for (uint32_t l = 0; l < n_in_group; l++) {
                float3 dc, spc;
                if (!BEAM && p.M == 1) L = load_light(p, l);  // hard shadows: the group is all lights
                const float3 lpos = p.M > 1 ? copy_position(p, L.pos, cloud, l) : L.pos;
                if (shade_pair<COUNT, BEAM ? 1 : 0>(p, r.P, r.N, r.d, r.mcolor, r.mshininess, lpos, L.intensity, L.color, cull, cn, dc, spc, cand, n_cand,
                                                    &last_occluder)) {
                    sum_d = sum_d + dc;  // (x + 0 == x for every sum that can occur here: the sums start at +0 and never become -0)
                    sum_s = sum_s + spc;
                }
            }
        }
#ifdef RT_DBG_TAIL
        {
            const unsigned long long now = dbg_globaltimer();
            if (now - dbg_last_start > dbg_worst) {
                dbg_worst = now - dbg_last_start;
                dbg_worst_info = dbg_cur_info;
                dbg_ph0 = dbg_tb - dbg_last_start;  // rebuild + chain wait + beam
                dbg_ph1 = now - dbg_tb;             // copies
                dbg_ph2 = 0ull;
            }
        }
#endif
        }  // lights of the unit
        if (DEFER && unit_parked) continue;
        if (r.hit) {
            if (!last) {
                const float tag = __uint_as_float(DEFER ? frame_tag + last_light : tag_base + group_index);
                st_relaxed_gpu(&own_rec->d, make_float4(sum_d.x, sum_d.y, sum_d.z, tag));
                st_relaxed_gpu(&own_rec->s, make_float4(sum_s.x, sum_s.y, sum_s.z, tag));
            } else {
                store_direct_sample<POINTS>(p, r, sum_d, sum_s);
            }
        } else if (r.valid && last) {
            p.samples[r.sample_index] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        if (!DEFER && last && n_groups > 1 && window < n_batches) {
            // the window's last unit of this batch: every lane has read its predecessor's record -- the slots of this
            // batch may be reused by the next window (release: ordered after the loads of all lanes)
            __syncwarp();
            if (lane == 0) st_release_gpu_u32(p.chain_done + b_in_win, tag_base + group_index);
        }
    }
    // the queue is empty: work off the deferred units together
    if (DEFER && BEAM && deferring) {
#ifdef RT_DBG_TAIL
        const unsigned long long dbg_h0 = dbg_globaltimer();
#endif
        finish_heavy_units<COUNT, POINTS>(gp, hq, hscratch, state, n_lights_pl, frame_tag, gridDim.x * (blockDim.x >> 5), COUNT ? &cn : nullptr, watchdog);
#ifdef RT_DBG_TAIL
        dbg_help_time += dbg_globaltimer() - dbg_h0;
#endif
    }
    if (COUNT) {
        atomicAdd(p.counters + 0, cn.closest);
        atomicAdd(p.counters + 1, cn.shadow);
        atomicAdd(p.counters + 2, cn.box);
        atomicAdd(p.counters + 3, cn.tri);
        atomicAdd(p.counters + 4, cn.sph);
        atomicAdd(p.counters + 5, cn.beam_prim);
    }
#ifdef RT_DBG_TAIL
    if (lane == 0) {  // per warp: start, end, start of its last unit, time spent waiting on chained records, units
        unsigned long long *o = g_dbg_tail + 12 * (size_t)warp_id;
        o[8] = dbg_ph0; o[9] = dbg_ph1; o[10] = dbg_ph2; o[11] = dbg_help_time;
        o[0] = dbg_t0; o[1] = dbg_globaltimer(); o[2] = dbg_last_start; o[3] = dbg_spin; o[4] = dbg_units; o[5] = dbg_worst; o[6] = dbg_worst_info;
        o[7] = dbg_helped;
    }
#endif
}

// raw cloud points (K x M x [f32;3], as handed over) -> float4 table scaled to scene units, the
// `p * (w2s_w, w2s_h, w2s_d)` of to_point_light_cloud (light.rs:208-216) evaluated once per point
__global__ void scale_clouds_kernel(const float *__restrict__ raw, float4 *__restrict__ out, uint32_t n, float w2s_w, float w2s_h, float w2s_d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_float4(raw[3 * i] * w2s_w, raw[3 * i + 1] * w2s_h, raw[3 * i + 2] * w2s_d, 0.0f);
}

// antialiased_raytrace's reduction (raytracer_renderer.rs:974-1015) + to_output + ChunkView::set.
// to_frame == 0: writes the rank's compact slab, local row j = image row tile_image_row(j).
// to_frame == 1: writes the pixel at its place in a FULL frame (row tile_image_row(j)).  With rt_multi_render that
// frame lives on device 0 and is reached through NVLink peer memory: the tile gather is these stores, there is no
// separate collective and no placement pass.
__global__ void resolve_kernel(const RenderParams p, uint32_t *__restrict__ out_argb, uint8_t *__restrict__ out_mask, int to_frame) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = p.local_rows * p.width;
    if (i >= n) return;
    size_t o = i;
    if (to_frame) {
        const uint32_t lr = i / p.width;
        o = (size_t)tile_image_row(p, lr) * p.width + (i - lr * p.width);
    }
    const float4 *s = p.samples + (size_t)i * p.S;
    float3 color;
    bool any = false;
    if (!(p.flags & RT_FLAG_ANTI_ALIASING)) {
        float4 v = s[0];
        color = f3(v.x, v.y, v.z);
        any = v.w != 0.0f;
    } else {
        const float scale = p.sample_scale;
        auto term = [&](uint32_t k) {
            float4 v = s[k];
            any = any || (v.w != 0.0f);
            return v.w != 0.0f ? f3(v.x, v.y, v.z) * scale : splat(0.0f);
        };
        if (!(p.flags & RT_FLAG_PACKET_SEMANTICS)) {
            float3 res = splat(0.0f);
            for (uint32_t k = 1; k < p.S; k++) res = term(k) + res;
            color = res + term(0);
        } else {
            float3 lane[8];
#pragma unroll
            for (uint32_t j = 0; j < 8; j++) {
                float3 res = splat(0.0f);
                for (uint32_t q = 1; q < p.S / 8; q++) res = term(8 * q + j) + res;
                lane[j] = res + term(j);
            }
            float3 q0 = lane[0] + lane[4], q1 = lane[1] + lane[5], q2 = lane[2] + lane[6], q3 = lane[3] + lane[7];
            color = (q0 + q2) + (q1 + q3);
        }
    }
    out_argb[o] = any ? encode_argb(color) : 0u;
    if (out_mask) out_mask[o] = any ? 1 : 0;
}

// debug/parity kernels: raw BVH queries for arbitrary rays
__global__ void trace_closest_kernel(DScene sc, uint32_t flags, const float *__restrict__ rays, uint32_t n, float *out_t, uint32_t *out_prim) {
    stage_top_nodes(sc);
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Counters cn = {0, 0, 0, 0, 0, 0};
    float3 o = f3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
    float3 d = normalized3(f3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
    ClosestHit h = trace_closest<false>(sc, o, d, (flags & RT_FLAG_BACKFACE_CULLING) != 0, cn);
    out_t[i] = h.slot >= 0 ? h.t : __int_as_float(0x7f800000);
    out_prim[i] = h.slot >= 0 ? h.prim : 0xFFFFFFFFu;
}

__global__ void trace_shadow_kernel(DScene sc, uint32_t flags, const float *__restrict__ rays, const float *__restrict__ max_dist, uint32_t n,
                                    float *out) {
    stage_top_nodes(sc);
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Counters cn = {0, 0, 0, 0, 0, 0};
    float3 o = f3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
    float3 d = normalized3(f3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
    ShadowResult s = trace_shadow<false>(sc, o, d, max_dist[i], (flags & RT_FLAG_BACKFACE_CULLING) != 0, cn);
    out[5 * i] = s.occluded ? 1.0f : 0.0f;
    out[5 * i + 1] = s.opacity;
    out[5 * i + 2] = s.filter.x;
    out[5 * i + 3] = s.filter.y;
    out[5 * i + 4] = s.filter.z;
}

// after the NCCL gather: place one rank's compact slab rows at their image rows
__global__ void scatter_rows_kernel(const uint8_t *__restrict__ slab, uint8_t *__restrict__ frame, uint32_t row_bytes, uint32_t local_rows,
                                    uint32_t tile_rows, uint32_t tile_rank, uint32_t tile_world) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // in 4-byte words when row_bytes % 4 == 0, else bytes
    size_t total = (size_t)row_bytes * local_rows;
    if (i >= total) return;
    uint32_t lr = (uint32_t)(i / row_bytes), off = (uint32_t)(i - (size_t)lr * row_bytes);
    uint32_t row = lr;
    if (tile_world > 1 && tile_rows) {
        uint32_t t = lr / tile_rows, r = lr - t * tile_rows;
        row = (t * tile_world + tile_rank) * tile_rows + r;
    }
    frame[(size_t)row * row_bytes + off] = slab[i];
}

// after the NCCL gather, all ranks at once: frame pixel (row, x) comes from rank (row / tile_rows) % world,
// local row (row / tile_rows / world) * tile_rows + row % tile_rows of that rank's slab (slabs are `slab_stride`
// u32 apart).  A hit pixel is 0xFFrrggbb and a miss is 0 (resolve_kernel), so the hit mask is slab != 0.
__global__ void place_tiles_kernel(const uint32_t *__restrict__ slabs, size_t slab_stride, uint32_t *__restrict__ frame, uint8_t *__restrict__ mask,
                                   uint32_t width, uint32_t height, uint32_t tile_rows, uint32_t tile_world) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)width * height) return;
    const uint32_t row = (uint32_t)(i / width), x = (uint32_t)(i - (size_t)row * width);
    const uint32_t t = row / tile_rows, r = row - t * tile_rows;
    const uint32_t rank = t % tile_world, lr = (t / tile_world) * tile_rows + r;
    const uint32_t v = slabs[rank * slab_stride + (size_t)lr * width + x];
    frame[i] = v;
    if (mask) mask[i] = v != 0u;
}

}  // namespace rtd
