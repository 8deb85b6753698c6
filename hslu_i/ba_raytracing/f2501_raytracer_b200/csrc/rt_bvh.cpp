// rt_bvh.cpp -- host BVH build (binned SAH BVH2, collapsed to BVH4, breadth-first node order)
// and packing of the scene into the float4 device layout described in rt_bvh.h.
#include "rt_bvh.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <queue>
#include <thread>

#include "rt_internal.h"

namespace rtb {
namespace {

const float INF = std::numeric_limits<float>::infinity();
constexpr float F32_EPS = 1.1920929e-7f;

struct Box {
    float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
    void grow(const Box &b) {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], b.lo[k]);
            hi[k] = std::max(hi[k], b.hi[k]);
        }
    }
    void grow_pt(const float *p) {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], p[k]);
            hi[k] = std::max(hi[k], p[k]);
        }
    }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0.0f;
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct Node2 {
    Box box;
    int32_t left = -1, right = -1;  // internal
    uint32_t first = 0, count = 0;  // leaf (count > 0)
};

struct Builder {
    std::vector<Box> pbox;       // padded primitive boxes
    std::vector<float> centroid; // 3 per primitive
    std::vector<uint32_t> order;
    std::vector<Node2> nodes;
    bool median_only = false;
    float c_isect = 1.25f;
    uint32_t max_leaf = MAX_LEAF;
    // what build() works on: the arrays above, or (a subtree built by a worker thread) its parent's
    const Box *pb = nullptr;
    const float *ce = nullptr;
    uint32_t *ord = nullptr;
    // parallel build: subtrees of at most `task_size` references (0 = off) are set aside as tasks while the
    // top of the tree is built, then built by worker threads and spliced in.  A subtree depends only on its
    // range of `order`, so the tree is the one the sequential recursion builds.
    uint32_t task_size = 0;
    struct Task {
        uint32_t first, count;
        int depth;
        int32_t parent;
        bool left;
    };
    std::vector<Task> tasks;

    unsigned wide_threads = 1;  // threads for the scans of big nodes (the sequential top of a parallel build)
    template <typename Fn>
    static void run_chunks(uint32_t first, uint32_t count, unsigned nthr, Fn fn) {
        std::vector<std::thread> pool;
        const uint32_t per = (count + nthr - 1) / nthr;
        for (unsigned t = 1; t < nthr; t++) {
            const uint32_t lo_i = first + std::min(count, t * per), hi_i = first + std::min(count, (t + 1) * per);
            pool.emplace_back([=]() { fn(t, lo_i, hi_i); });
        }
        fn(0u, first, first + std::min(count, per));
        for (auto &th : pool) th.join();
    }

    int32_t child(uint32_t first, uint32_t count, int depth, int32_t parent, bool left) {
        if (task_size && count <= task_size && count >= 1024) {
            tasks.push_back({first, count, depth, parent, left});
            return -2;  // patched by build_parallel
        }
        return build(first, count, depth);
    }

    int32_t build(uint32_t first, uint32_t count, int depth) {
        Node2 nd;
        Box cb;
        // big nodes (the top of a large tree) are scanned by all threads: min / max and counts merge exactly, so the
        // result does not depend on how the range is cut
        const unsigned nthr = (wide_threads > 1 && count >= 262144u) ? wide_threads : 1u;
        if (nthr > 1) {
            std::vector<Box> nb(nthr), cbs(nthr);
            run_chunks(first, count, nthr, [&](unsigned t, uint32_t lo_i, uint32_t hi_i) {
                for (uint32_t i = lo_i; i < hi_i; i++) {
                    nb[t].grow(pb[ord[i]]);
                    cbs[t].grow_pt(&ce[3 * ord[i]]);
                }
            });
            for (unsigned t = 0; t < nthr; t++) {
                nd.box.grow(nb[t]);
                cb.grow(cbs[t]);
            }
        } else {
            for (uint32_t i = first; i < first + count; i++) {
                nd.box.grow(pb[ord[i]]);
                cb.grow_pt(&ce[3 * ord[i]]);
            }
        }
        int32_t idx = (int32_t)nodes.size();
        nodes.push_back(nd);
        auto make_leaf = [&]() {
            nodes[idx].first = first;
            nodes[idx].count = count;
            return idx;
        };
        if (count == 1) return make_leaf();
        int axis = -1;
        uint32_t mid = 0;
        const int NB = 16;
        const float C_TRAV = 1.0f, C_ISECT = c_isect;
        float best = INF;
        if (!median_only && depth < 28) {
            float parent_area = nd.box.area();
            for (int a = 0; a < 3; a++) {
                float lo = cb.lo[a], ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0.0f)) continue;
                Box bb[NB];
                uint32_t bc[NB] = {0};
                float scale = (float)NB / ext;
                if (nthr > 1) {
                    std::vector<Box> tb((size_t)nthr * NB);
                    std::vector<uint32_t> tc((size_t)nthr * NB, 0u);
                    run_chunks(first, count, nthr, [&](unsigned t, uint32_t lo_i, uint32_t hi_i) {
                        for (uint32_t i = lo_i; i < hi_i; i++) {
                            uint32_t p = ord[i];
                            int b = std::min(NB - 1, std::max(0, (int)((ce[3 * p + a] - lo) * scale)));
                            tc[(size_t)t * NB + b]++;
                            tb[(size_t)t * NB + b].grow(pb[p]);
                        }
                    });
                    for (unsigned t = 0; t < nthr; t++)
                        for (int b = 0; b < NB; b++) {
                            bc[b] += tc[(size_t)t * NB + b];
                            bb[b].grow(tb[(size_t)t * NB + b]);
                        }
                } else {
                    for (uint32_t i = first; i < first + count; i++) {
                        uint32_t p = ord[i];
                        int b = std::min(NB - 1, std::max(0, (int)((ce[3 * p + a] - lo) * scale)));
                        bc[b]++;
                        bb[b].grow(pb[p]);
                    }
                }
                float rarea[NB];
                uint32_t rcnt[NB];
                Box acc;
                uint32_t c = 0;
                for (int b = NB - 1; b > 0; b--) {
                    acc.grow(bb[b]);
                    c += bc[b];
                    rarea[b] = acc.area();
                    rcnt[b] = c;
                }
                acc = Box();
                c = 0;
                for (int b = 0; b < NB - 1; b++) {
                    acc.grow(bb[b]);
                    c += bc[b];
                    if (c == 0 || rcnt[b + 1] == 0) continue;
                    float cost = C_TRAV + C_ISECT * (acc.area() * (float)c + rarea[b + 1] * (float)rcnt[b + 1]) / std::max(parent_area, 1e-30f);
                    if (cost < best) {
                        best = cost;
                        axis = a;
                        mid = b + 1;  // bins [0, mid) go left
                    }
                }
            }
            if (axis >= 0) {
                if (count <= max_leaf && best >= C_ISECT * (float)count) return make_leaf();
                float lo = cb.lo[axis], scale = (float)NB / (cb.hi[axis] - cb.lo[axis]);
                uint32_t split_bin = mid;
                uint32_t *it = std::partition(ord + first, ord + first + count, [&](uint32_t p) {
                    int b = std::min(NB - 1, std::max(0, (int)((ce[3 * p + axis] - lo) * scale)));
                    return (uint32_t)b < split_bin;
                });
                mid = (uint32_t)(it - ord);
                if (mid == first || mid == first + count) axis = -1;
            }
        }
        if (axis < 0) {
            if (count <= max_leaf) return make_leaf();
            // median split along the widest centroid axis (also the fallback that bounds the depth)
            int a = 0;
            for (int k = 1; k < 3; k++)
                if (cb.hi[k] - cb.lo[k] > cb.hi[a] - cb.lo[a]) a = k;
            mid = first + count / 2;
            std::nth_element(ord + first, ord + mid, ord + first + count, [&](uint32_t x, uint32_t y) { return ce[3 * x + a] < ce[3 * y + a]; });
        }
        int32_t l = child(first, mid - first, depth + 1, idx, true);
        int32_t r = child(mid, first + count - mid, depth + 1, idx, false);
        nodes[idx].left = l;
        nodes[idx].right = r;
        return idx;
    }

    // root of the tree over all of `order`, built with `threads` threads
    int32_t build_parallel(uint32_t n, unsigned threads) {
        pb = pbox.data();
        ce = centroid.data();
        ord = order.data();
        tasks.clear();
        wide_threads = threads;
        task_size = (threads > 1 && n >= 65536) ? std::max<uint32_t>(n / (4 * threads), 2048) : 0;
        const int32_t root = build(0, n, 0);
        if (tasks.empty()) return root;
        std::vector<std::vector<Node2>> sub(tasks.size());
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= tasks.size()) return;
                Builder w;
                w.median_only = median_only;
                w.c_isect = c_isect;
                w.max_leaf = max_leaf;
                w.pb = pb;
                w.ce = ce;
                w.ord = ord;
                w.nodes.reserve(2 * (size_t)tasks[i].count);
                w.build(tasks[i].first, tasks[i].count, tasks[i].depth);
                sub[i] = std::move(w.nodes);
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; t++) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
        for (size_t i = 0; i < tasks.size(); i++) {  // splice: the subtree's root is its node 0
            const int32_t offset = (int32_t)nodes.size();
            for (Node2 nd : sub[i]) {
                if (nd.count == 0) {
                    nd.left += offset;
                    nd.right += offset;
                }
                nodes.push_back(nd);
            }
            (tasks[i].left ? nodes[tasks[i].parent].left : nodes[tasks[i].parent].right) = offset;
        }
        return root;
    }
};

inline float u2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

}  // namespace

int check_scene_desc(const rt_scene_desc *d) {
    if (!d) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_create: null scene description");
    if ((d->n_triangles && !d->triangles) || (d->n_spheres && !d->spheres) || (d->n_materials && !d->materials) ||
        (d->n_lights && !d->lights))
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_create: null array with non-zero count");
    const uint64_t n64 = (uint64_t)d->n_triangles + d->n_spheres;
    if (n64 > LEAF_FIRST_MASK) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_create: too many primitives (%llu)", (unsigned long long)n64);
    if (d->n_materials > PRIM_MATERIAL_MASK) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_create: too many materials");
    return RT_OK;
}

// materials (48 B each) and lights (32 B each) in their device layout; transmissive[m] = TransmissionProperties::mask
void pack_materials_and_lights(const rt_scene_desc *d, FlatScene &out, std::vector<uint8_t> &transmissive) {
    out.n_mats = d->n_materials;
    out.n_lights = d->n_lights;
    transmissive.assign(d->n_materials, 0);
    out.mats.resize(12 * (size_t)d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const rt_material &m = d->materials[i];
        // TransmissionProperties::mask (material.rs:44-50)
        bool tm = m.has_opacity && !(std::fabs(m.opacity - 0.0f) <= F32_EPS);
        transmissive[i] = tm;
        float *o = &out.mats[12 * (size_t)i];
        o[0] = m.color[0]; o[1] = m.color[1]; o[2] = m.color[2]; o[3] = m.metallic;
        o[4] = m.shininess; o[5] = m.ior; o[6] = m.opacity; o[7] = m.boost;
        // compute_fresnel against air-like surroundings (other_ior = 1.0: every shadow / transmittance query,
        // raytracer.rs:60): its two divisions depend on the material alone -- evaluated here once, in f32, the same
        // IEEE operations (material.rs:480-500)
        const float f0r = (1.0f - m.ior) / (1.0f + m.ior);
        o[8] = u2f(m.has_opacity ? 1u : 0u); o[9] = u2f(tm ? 1u : 0u); o[10] = f0r * f0r; o[11] = 1.0f / m.ior;
    }
    out.lights.resize(8 * (size_t)d->n_lights);
    for (uint32_t i = 0; i < d->n_lights; i++) {
        const rt_light &l = d->lights[i];
        float *o = &out.lights[8 * (size_t)i];
        o[0] = l.position[0]; o[1] = l.position[1]; o[2] = l.position[2]; o[3] = l.intensity;
        o[4] = l.color[0]; o[5] = l.color[1]; o[6] = l.color[2]; o[7] = 0;
    }
}

int build_flat_scene(const rt_scene_desc *d, FlatScene &out) {
    if (int rc = check_scene_desc(d)) return rc;
    const uint32_t n = d->n_triangles + d->n_spheres, ns = d->n_spheres, nt = d->n_triangles;
    for (uint32_t i = 0; i < ns; i++)
        if (d->spheres[i].material >= d->n_materials) return rt_fail(RT_ERR_INVALID_ARGUMENT, "sphere %u: material index out of range", i);
    for (uint32_t i = 0; i < nt; i++)
        if (d->triangles[i].material >= d->n_materials) return rt_fail(RT_ERR_INVALID_ARGUMENT, "triangle %u: material index out of range", i);

    const bool verbose = getenv("RT_BVH_VERBOSE") != nullptr;  // stage timings on stderr (tuning)
    auto t_last = std::chrono::steady_clock::now();
    auto stage = [&](const char *name) {
        if (!verbose) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[rt_bvh] %-28s %8.1f ms\n", name, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    out = FlatScene();
    std::vector<uint8_t> transmissive;
    pack_materials_and_lights(d, out, transmissive);

    // ---- BVH references. Every primitive gets one reference. Optionally (RT_BVH_SPLIT_DIV=d > 0)
    // triangles whose longest edge exceeds scene_extent / d are referenced once per cell of a k x k
    // barycentric subdivision (k*k sub-triangles with tight boxes); the device accepts a hit only
    // in the reference whose cell contains the hit's (u, v) (rt_device.cuh tri_cell), so each
    // (ray, triangle) hit is still reported exactly once -- required because shadow queries
    // accumulate over transmissive hits. OFF by default: on the reference's scenes most shading
    // points lie ON the large wall/floor triangles, whose shadow rays then cross many cell boxes
    // (measured on B200, high_quality workload: d = 24 -> 49 box tests per ray instead of 21 and
    // 98 ms instead of 48 ms per frame; parity unchanged). Kept for scenes with large occluders.
    Builder b;
    if (const char *e = getenv("RT_BVH_CISECT")) b.c_isect = (float)atof(e);   // tuning knobs for experiments
    if (const char *e = getenv("RT_BVH_MAX_LEAF")) b.max_leaf = std::min<uint32_t>(MAX_LEAF, std::max(1, atoi(e)));
    float split_div = 0.0f;
    if (const char *e = getenv("RT_BVH_SPLIT_DIV")) split_div = (float)atof(e);
    unsigned build_threads = std::min(32u, std::max(1u, std::thread::hardware_concurrency()));  // (memory-bound beyond that)
    if (const char *e = getenv("RT_BVH_THREADS")) build_threads = (unsigned)std::max(1, atoi(e));
    Box exact_scene;
    float max_abs = 1.0f;
    for (uint32_t i = 0; i < ns; i++) {
        const rt_sphere &s = d->spheres[i];
        float r = std::fabs(s.radius);
        for (int k = 0; k < 3; k++) {
            float c[2] = {s.center[k] - r, s.center[k] + r};
            if (!std::isfinite(c[0]) || !std::isfinite(c[1])) return rt_fail(RT_ERR_INVALID_ARGUMENT, "sphere %u has non-finite coordinates", i);
            exact_scene.lo[k] = std::min(exact_scene.lo[k], c[0]);
            exact_scene.hi[k] = std::max(exact_scene.hi[k], c[1]);
        }
    }
    for (uint32_t i = 0; i < nt; i++) {
        const rt_triangle &t = d->triangles[i];
        for (const float *v : {t.v1, t.v2, t.v3}) {
            for (int k = 0; k < 3; k++)
                if (!std::isfinite(v[k])) return rt_fail(RT_ERR_INVALID_ARGUMENT, "triangle %u has non-finite coordinates", i);
            exact_scene.grow_pt(v);
        }
    }
    float max_extent = 0.0f;
    for (int k = 0; k < 3; k++) {
        if (n) {
            max_abs = std::max(max_abs, std::max(std::fabs(exact_scene.lo[k]), std::fabs(exact_scene.hi[k])));
            max_extent = std::max(max_extent, exact_scene.hi[k] - exact_scene.lo[k]);
        }
    }
    // Padding: the absolute part covers the slab arithmetic (|o| * 2^-24 relative error in t) and the
    // displacement of hits the float triangle solve accepts just outside the exact triangle; the
    // relative part scales with the reference's own size (DESIGN.md "Conservative traversal").
    out.box_pad = 1e-4f * max_abs;
    std::vector<uint32_t> ref_prim, ref_tag;
    auto add_ref = [&](uint32_t prim, uint32_t tag, Box bx) {
        for (int k = 0; k < 3; k++) {
            float pad = out.box_pad + 1e-3f * (bx.hi[k] - bx.lo[k]);
            bx.lo[k] -= pad;
            bx.hi[k] += pad;
        }
        ref_prim.push_back(prim);
        ref_tag.push_back(tag);
        b.pbox.push_back(bx);
        for (int k = 0; k < 3; k++) b.centroid.push_back(0.5f * (bx.lo[k] + bx.hi[k]));
    };
    for (uint32_t i = 0; i < ns; i++) {
        const rt_sphere &s = d->spheres[i];
        float r = std::fabs(s.radius);
        Box bx;
        for (int k = 0; k < 3; k++) {
            bx.lo[k] = s.center[k] - r;
            bx.hi[k] = s.center[k] + r;
        }
        add_ref(i, 0, bx);
    }
    const float split_len = (split_div > 0.0f && max_extent > 0.0f) ? max_extent / split_div : INF;
    for (uint32_t i = 0; i < nt; i++) {
        const rt_triangle &t = d->triangles[i];
        double e1[3], e2[3], longest = 0.0;
        for (int k = 0; k < 3; k++) {
            e1[k] = (double)t.v2[k] - t.v1[k];
            e2[k] = (double)t.v3[k] - t.v1[k];
        }
        for (int a = 0; a < 3; a++) {  // edge lengths |e1|, |e2|, |e2 - e1|
            double l2 = 0.0;
            for (int k = 0; k < 3; k++) {
                double c = a == 0 ? e1[k] : (a == 1 ? e2[k] : e2[k] - e1[k]);
                l2 += c * c;
            }
            longest = std::max(longest, std::sqrt(l2));
        }
        uint32_t kdiv = 1;
        if (longest > split_len) kdiv = (uint32_t)std::min(32.0, std::ceil(longest / split_len));
        if (kdiv < 2 || ref_prim.size() + (size_t)kdiv * kdiv > (size_t)LEAF_FIRST_MASK / 2) {
            Box bx;
            bx.grow_pt(t.v1);
            bx.grow_pt(t.v2);
            bx.grow_pt(t.v3);
            add_ref(ns + i, 0, bx);
            continue;
        }
        auto corner = [&](uint32_t iu, uint32_t jv, float *outp) {
            for (int k = 0; k < 3; k++) outp[k] = (float)((double)t.v1[k] + e1[k] * ((double)iu / kdiv) + e2[k] * ((double)jv / kdiv));
        };
        for (uint32_t iu = 0; iu < kdiv; iu++) {
            for (uint32_t jv = 0; iu + jv < kdiv; jv++) {
                float a[3], bb[3], c[3];
                Box bx;  // upright cell (iu, jv)
                corner(iu, jv, a);
                corner(iu + 1, jv, bb);
                corner(iu, jv + 1, c);
                bx.grow_pt(a);
                bx.grow_pt(bb);
                bx.grow_pt(c);
                add_ref(ns + i, TRI_CELL(kdiv, iu, jv, 0), bx);
                if (iu + jv + 2 <= kdiv) {  // inverted cell sharing the same (iu, jv) square
                    Box by;
                    corner(iu + 1, jv + 1, a);
                    by.grow_pt(a);
                    by.grow_pt(bb);
                    by.grow_pt(c);
                    add_ref(ns + i, TRI_CELL(kdiv, iu, jv, 1), by);
                }
            }
        }
    }
    stage("primitive boxes");
    const uint32_t n_refs = (uint32_t)ref_prim.size();
    out.n_prims = n_refs;  // slots = references
    Box scene;
    for (uint32_t i = 0; i < n_refs; i++) scene.grow(b.pbox[i]);
    for (int k = 0; k < 3; k++) {
        out.scene_lo[k] = n_refs ? scene.lo[k] : 0.0f;
        out.scene_hi[k] = n_refs ? scene.hi[k] : 0.0f;
    }

    // BVH2 -> BVH4, retrying with balanced median splits if the tree is too deep for the device stack
    const uint32_t MAX_BVH4_DEPTH = 20;  // device stack holds 3 * depth + 1 references (rt_device.cuh)
    struct Ref4 {
        int32_t node2;
        uint32_t depth;
    };
    for (int attempt = 0; attempt < 2; attempt++) {
        b.median_only = attempt == 1;
        b.order.resize(n_refs);
        for (uint32_t i = 0; i < n_refs; i++) b.order[i] = i;
        b.nodes.clear();
        b.nodes.reserve(2 * (size_t)n_refs + 1);
        int32_t root = n_refs ? b.build_parallel(n_refs, build_threads) : -1;
        stage("BVH2 build");

        out.nodes.clear();
        out.max_depth = 0;
        std::queue<Ref4> q;  // breadth-first so the top levels are contiguous at the front
        auto new_node = [&]() {
            size_t at = out.nodes.size();
            out.nodes.resize(at + 32);
            for (int c = 0; c < 4; c++) {
                for (int k = 0; k < 3; k++) {
                    out.nodes[at + 4 * k + c] = INF;         // lo
                    out.nodes[at + 4 * (3 + k) + c] = -INF;  // hi
                }
                out.nodes[at + 24 + c] = u2f(CHILD_EMPTY);
                out.nodes[at + 28 + c] = 0.0f;
            }
            return (uint32_t)(at / 32);
        };
        new_node();  // node 0 = root
        if (root >= 0) q.push({root, 1});
        uint32_t next_index = 0;
        // node `next_index` is filled from the queue front; children get indices in push order
        std::vector<int32_t> kids;
        uint32_t allocated = 1;
        while (!q.empty()) {
            Ref4 cur = q.front();
            q.pop();
            uint32_t me = next_index++;
            out.max_depth = std::max(out.max_depth, cur.depth);
            kids.clear();
            const Node2 &n2 = b.nodes[cur.node2];
            if (n2.count > 0) {
                kids.push_back(cur.node2);  // root is a leaf: single child
            } else {
                kids.push_back(n2.left);
                kids.push_back(n2.right);
                while (kids.size() < 4) {
                    int bi = -1;
                    float ba = -1.0f;
                    for (size_t i = 0; i < kids.size(); i++) {
                        const Node2 &k2 = b.nodes[kids[i]];
                        if (k2.count == 0 && k2.box.area() > ba) {
                            ba = k2.box.area();
                            bi = (int)i;
                        }
                    }
                    if (bi < 0) break;
                    int32_t victim = kids[bi];
                    kids[bi] = b.nodes[victim].left;
                    kids.push_back(b.nodes[victim].right);
                }
            }
            size_t at = 32 * (size_t)me;
            for (size_t c = 0; c < kids.size(); c++) {
                const Node2 &k2 = b.nodes[kids[c]];
                for (int k = 0; k < 3; k++) {
                    out.nodes[at + 4 * k + c] = k2.box.lo[k];
                    out.nodes[at + 4 * (3 + k) + c] = k2.box.hi[k];
                }
                uint32_t ref;
                if (k2.count > 0) {
                    ref = CHILD_LEAF | ((k2.count - 1) << 27) | k2.first;
                } else {
                    ref = allocated++;
                    new_node();
                    q.push({kids[c], cur.depth + 1});
                }
                out.nodes[at + 24 + c] = u2f(ref);
            }
        }
        out.n_nodes = (uint32_t)(out.nodes.size() / 32);
        stage("collapse to BVH4");
        if (out.max_depth <= MAX_BVH4_DEPTH) break;
        if (attempt == 1) return rt_fail(RT_ERR_INVALID_ARGUMENT, "BVH depth %u exceeds the traversal stack", out.max_depth);
    }
    out.root_ref = 0;

    // primitive records in slot order
    out.prims.assign(16 * (size_t)n_refs, 0.0f);
    out.shade_n.assign(4 * (size_t)n_refs, 0.0f);
    uint32_t best_any = 0xFFFFFFFFu, best_tm = 0xFFFFFFFFu;
    for (uint32_t slot = 0; slot < n_refs; slot++) {
        uint32_t p = ref_prim[b.order[slot]];
        float *q = &out.prims[16 * (size_t)slot];
        if (p < ns) {
            const rt_sphere &s = d->spheres[p];
            q[0] = s.center[0]; q[1] = s.center[1]; q[2] = s.center[2];
            q[3] = s.radius * s.radius;  // SphereData::with_material (sphere.rs:43)
            q[12] = u2f(p);
            q[13] = u2f(s.material | PRIM_SPHERE | (transmissive[s.material] ? PRIM_TRANSMISSIVE : 0u));
        } else {
            const rt_triangle &t = d->triangles[p - ns];
            float e1[3], e2[3];
            for (int k = 0; k < 3; k++) {
                e1[k] = t.v2[k] - t.v1[k];  // triangle.rs:89-90
                e2[k] = t.v3[k] - t.v1[k];
            }
            // first row of ultraviolet's Mat3::inversed for columns [d, -e1, -e2]: (-e1) x (-e2),
            // with cross = (y.mul_add(oz, -z*oy), ...); the sign flips cancel exactly.
            float g[3] = {std::fmaf(e1[1], e2[2], -(e1[2] * e2[1])), std::fmaf(e1[2], e2[0], -(e1[0] * e2[2])),
                          std::fmaf(e1[0], e2[1], -(e1[1] * e2[0]))};
            q[0] = t.v1[0]; q[1] = t.v1[1]; q[2] = t.v1[2]; q[3] = e1[0];
            q[4] = e1[1]; q[5] = e1[2]; q[6] = e2[0]; q[7] = e2[1];
            q[8] = e2[2]; q[9] = g[0]; q[10] = g[1]; q[11] = g[2];
            q[12] = u2f(p);
            q[13] = u2f(t.material | (transmissive[t.material] ? PRIM_TRANSMISSIVE : 0u));
            q[14] = u2f(ref_tag[b.order[slot]]);
            float emax = 0.0f;
            for (int k = 0; k < 3; k++) emax = std::max(emax, std::max(std::fabs(e1[k]), std::fabs(e2[k])));
            q[15] = 64.0f * 1.1920929e-7f * emax;  // error bound factor of tri_test's early-out (rt_device.cuh)
            float *sn = &out.shade_n[4 * (size_t)slot];
            sn[0] = t.normal[0]; sn[1] = t.normal[1]; sn[2] = t.normal[2];
            if (p < best_any) {
                best_any = p;
                out.nan_hit_slot = (int32_t)slot;
            }
            if (transmissive[t.material] && p < best_tm) {
                best_tm = p;
                out.nan_hit_slot_culled = (int32_t)slot;
            }
        }
    }
    stage("primitive records");
    return RT_OK;
}

}  // namespace rtb

// Debug / parity entry point (include/rt_api.h): host-only build + digest of the flattened arrays.
extern "C" int rt_bvh_digest(const rt_scene_desc *desc, uint64_t *out_digest, uint32_t *out_nodes, uint32_t *out_depth) {
    if (!out_digest) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_bvh_digest: null output");
    rtb::FlatScene flat;
    int rc = rtb::build_flat_scene(desc, flat);
    if (rc) return rc;
    uint64_t h = 1469598103934665603ull;  // FNV-1a over nodes, primitive records and stored normals
    auto mix = [&](const std::vector<float> &v) {
        const unsigned char *p = (const unsigned char *)v.data();
        for (size_t i = 0, n = v.size() * sizeof(float); i < n; i++) {
            h ^= p[i];
            h *= 1099511628211ull;
        }
    };
    mix(flat.nodes);
    mix(flat.prims);
    mix(flat.shade_n);
    *out_digest = h;
    if (out_nodes) *out_nodes = flat.n_nodes;
    if (out_depth) *out_depth = flat.max_depth;
    return RT_OK;
}

