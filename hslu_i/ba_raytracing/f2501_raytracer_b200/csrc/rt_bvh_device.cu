// rt_bvh_device.cu -- device-side BVH build for large scenes (sm_100a).
//
// The reference has no acceleration structure at all (src/raytracing/raytracer.rs:48,180 are linear scans), so the
// builder is new work on the path the reference times end to end (src/main.rs:331-334: the Rust shim calls
// rt_scene_create inside render()).  The host builder (rt_bvh.cpp, binned SAH, multi-threaded) takes seconds for the
// 4.2 M-triangle scene; this one builds the same device layout (rt_bvh.h) on the GPU in milliseconds:
//
//   1. exact scene bounds + validation (finite coordinates, material indices)          one reduction pass
//   2. padded primitive boxes (the host's padding rule, same float operations) and 63-bit Morton codes of
//      their centres
//   3. radix sort of (code, primitive)                                                 cub::DeviceRadixSort
//   4. binary radix tree over the sorted codes (Karras 2012; equal codes are split by index, which keeps runs of
//      duplicates balanced), boxes fitted bottom-up with one atomic flag per node
//   5. collapse to BVH4, level by level from the root: every BVH4 node opens the child with the largest area until
//      it has four (the host's rule); subtrees of <= MAX_LEAF primitives become leaves; a prefix sum per level gives
//      the children their indices, so the node array is breadth-first and its numbering deterministic
//   6. primitive records (64 B) and stored normals in leaf (= sorted) order
//
// Everything the traversal's exactness rests on is unchanged: boxes are padded by the same rule, leaves hold the
// same records, the tie rule travels in the primitive index.  If the tree comes out deeper than the traversal
// stack allows, the caller falls back to the host builder.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "rt_bvh.h"
#include "rt_bvh_device.h"
#include "rt_internal.h"

namespace rtb {
namespace {

#define CUB_OK(expr)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            rc = rt_fail(_e == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY : RT_ERR_CUDA,    \
                         "device BVH build: %s failed: %s", #expr, cudaGetErrorString(_e));       \
            goto done;                                                                            \
        }                                                                                         \
    } while (0)

constexpr uint32_t LEAF_BIT = 0x80000000u;  // child of a radix-tree node: bit 31 set = sorted primitive position
constexpr int TPB = 256;

struct Bounds {  // ordered-int encodings of floats for atomicMin / atomicMax
    int lo[3], hi[3], max_abs;
    unsigned int bad_coord, bad_material;
};
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__host__ __device__ __forceinline__ float ord2f(int i) {
    int j = i >= 0 ? i : i ^ 0x7FFFFFFF;
#ifdef __CUDA_ARCH__
    return __int_as_float(j);
#else
    float f;
    memcpy(&f, &j, 4);
    return f;
#endif
}

// exact box of primitive i (spheres first, then triangles: the reference iteration order)
__device__ __forceinline__ void prim_box(const rt_sphere *sph, const rt_triangle *tri, uint32_t ns, uint32_t i, float lo[3], float hi[3]) {
    if (i < ns) {
        const rt_sphere s = sph[i];
        const float r = fabsf(s.radius);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = s.center[k] - r;
            hi[k] = s.center[k] + r;
        }
    } else {
        const rt_triangle *t = tri + (i - ns);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float a = t->v1[k], b = t->v2[k], c = t->v3[k];
            lo[k] = fminf(a, fminf(b, c));
            hi[k] = fmaxf(a, fmaxf(b, c));
        }
    }
}

__global__ void bounds_kernel(const rt_sphere *sph, const rt_triangle *tri, uint32_t ns, uint32_t n, uint32_t n_materials, Bounds *out) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    unsigned bad_c = 0, bad_m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float l[3], h[3];
        prim_box(sph, tri, ns, i, l, h);
        const uint32_t mat = i < ns ? sph[i].material : tri[i - ns].material;
        bad_m |= mat >= n_materials;
        // NaN or infinite coordinates (fminf / fmaxf drop a NaN operand, so the raw values are looked at)
        float sum = 0.0f;
        if (i < ns) {
            sum = ((sph[i].center[0] + sph[i].center[1]) + sph[i].center[2]) + sph[i].radius;
        } else {
            const rt_triangle *t = tri + (i - ns);
#pragma unroll
            for (int k = 0; k < 3; k++) sum += (fabsf(t->v1[k]) + fabsf(t->v2[k])) + fabsf(t->v3[k]);
        }
        bad_c |= !(sum - sum == 0.0f);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            bad_c |= !(l[k] - l[k] == 0.0f) || !(h[k] - h[k] == 0.0f);
            lo[k] = fminf(lo[k], l[k]);
            hi[k] = fmaxf(hi[k], h[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
        }
    }
    bad_c = __any_sync(0xFFFFFFFFu, bad_c);
    bad_m = __any_sync(0xFFFFFFFFu, bad_m);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (lo[k] <= hi[k]) {
                atomicMin(&out->lo[k], f2ord(lo[k]));
                atomicMax(&out->hi[k], f2ord(hi[k]));
            }
        }
        if (bad_c) atomicOr(&out->bad_coord, 1u);
        if (bad_m) atomicOr(&out->bad_material, 1u);
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {  // 21 bits -> every third bit
    x &= 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// padded boxes (host rule: pad = box_pad + 1e-3 * extent per axis, rt_bvh.cpp add_ref) and Morton codes of their centres
__global__ void refs_kernel(const rt_sphere *sph, const rt_triangle *tri, uint32_t ns, uint32_t n, float box_pad, float3 slo, float3 sinv,
                            float4 *blo, float4 *bhi, unsigned long long *keys, uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lo[3], hi[3];
    prim_box(sph, tri, ns, i, lo, hi);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float pad = box_pad + 1e-3f * (hi[k] - lo[k]);
        lo[k] -= pad;
        hi[k] += pad;
    }
    blo[i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
    bhi[i] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    const float cx = (0.5f * (lo[0] + hi[0]) - slo.x) * sinv.x, cy = (0.5f * (lo[1] + hi[1]) - slo.y) * sinv.y, cz = (0.5f * (lo[2] + hi[2]) - slo.z) * sinv.z;
    const float S = 2097151.0f;  // 2^21 - 1
    const unsigned long long x = (unsigned long long)fminf(fmaxf(cx * S, 0.0f), S), y = (unsigned long long)fminf(fmaxf(cy * S, 0.0f), S),
                             z = (unsigned long long)fminf(fmaxf(cz * S, 0.0f), S);
    keys[i] = spread21(x) | (spread21(y) << 1) | (spread21(z) << 2);
    vals[i] = i;
}

// common-prefix length of sorted positions i and j (equal codes: continue into the index bits)
__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a != b) return __clzll((long long)(a ^ b));
    return 64 + __clz(i ^ j);
}

// Karras 2012: internal node i of the binary radix tree over n sorted keys (n >= 2)
__global__ void radix_tree_kernel(const unsigned long long *keys, int n, uint32_t *left, uint32_t *right, uint32_t *parent_int, uint32_t *parent_leaf,
                                  uint32_t *first, uint32_t *last) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t lc = lo == gamma ? (LEAF_BIT | (uint32_t)gamma) : (uint32_t)gamma;
    const uint32_t rc = hi == gamma + 1 ? (LEAF_BIT | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    left[i] = lc;
    right[i] = rc;
    first[i] = (uint32_t)lo;
    last[i] = (uint32_t)hi;
    if (lc & LEAF_BIT) parent_leaf[gamma] = (uint32_t)i; else parent_int[gamma] = (uint32_t)i;
    if (rc & LEAF_BIT) parent_leaf[gamma + 1] = (uint32_t)i; else parent_int[gamma + 1] = (uint32_t)i;
    if (i == 0) parent_int[0] = 0xFFFFFFFFu;
}

// boxes bottom-up: the second thread to arrive at a node fits it and moves on
__global__ void refit_kernel(int n, const uint32_t *vals, const float4 *blo, const float4 *bhi, const uint32_t *left, const uint32_t *right,
                             const uint32_t *parent_int, const uint32_t *parent_leaf, float4 *nlo, float4 *nhi, unsigned int *flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t cur = parent_leaf[i];
    while (cur != 0xFFFFFFFFu) {
        __threadfence();
        if (atomicAdd(&flags[cur], 1u) == 0u) return;  // first to arrive: the sibling subtree is not fitted yet
        __threadfence();
        const uint32_t lc = left[cur], rc = right[cur];
        // (children fitted by other threads: read past L1, their stores were fenced before the flag)
        const float4 llo = (lc & LEAF_BIT) ? blo[vals[lc & ~LEAF_BIT]] : __ldcg(nlo + lc), lhi = (lc & LEAF_BIT) ? bhi[vals[lc & ~LEAF_BIT]] : __ldcg(nhi + lc);
        const float4 rlo = (rc & LEAF_BIT) ? blo[vals[rc & ~LEAF_BIT]] : __ldcg(nlo + rc), rhi = (rc & LEAF_BIT) ? bhi[vals[rc & ~LEAF_BIT]] : __ldcg(nhi + rc);
        nlo[cur] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
        nhi[cur] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
        cur = parent_int[cur];
    }
}

struct TreeView {
    const uint32_t *left, *right, *first, *last, *vals;
    const float4 *blo, *bhi, *nlo, *nhi;
};
__device__ __forceinline__ uint32_t child_count(const TreeView &t, uint32_t c) { return (c & LEAF_BIT) ? 1u : t.last[c] - t.first[c] + 1u; }
__device__ __forceinline__ bool child_is_leaf(const TreeView &t, uint32_t c) { return child_count(t, c) <= MAX_LEAF; }
__device__ __forceinline__ void child_box(const TreeView &t, uint32_t c, float4 &lo, float4 &hi) {
    if (c & LEAF_BIT) {
        const uint32_t p = t.vals[c & ~LEAF_BIT];
        lo = t.blo[p];
        hi = t.bhi[p];
    } else {
        lo = t.nlo[c];
        hi = t.nhi[c];
    }
}
__device__ __forceinline__ float box_area(float4 lo, float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}
// the (up to four) children of the BVH4 node made from radix-tree node `node`: open the internal child with the
// largest area until there are four (rt_bvh.cpp's rule); a subtree of <= MAX_LEAF primitives is a leaf, never opened
__device__ __forceinline__ int gather_kids(const TreeView &t, uint32_t node, uint32_t kids[4]) {
    int nk = 2;
    kids[0] = t.left[node];
    kids[1] = t.right[node];
    while (nk < 4) {
        int bi = -1;
        float ba = -1.0f;
        for (int i = 0; i < nk; i++) {
            if (child_is_leaf(t, kids[i])) continue;
            float4 lo, hi;
            child_box(t, kids[i], lo, hi);
            const float a = box_area(lo, hi);
            if (a > ba) {
                ba = a;
                bi = i;
            }
        }
        if (bi < 0) break;
        const uint32_t victim = kids[bi];
        kids[bi] = t.left[victim];
        kids[nk++] = t.right[victim];
    }
    return nk;
}

// one level of the collapse, pass 1: how many internal children does each node of the level have
__global__ void level_count_kernel(TreeView t, const uint32_t *level, uint32_t n_level, uint32_t *counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_level) return;
    uint32_t kids[4];
    const int nk = gather_kids(t, level[i], kids);
    uint32_t c = 0;
    for (int k = 0; k < nk; k++) c += child_is_leaf(t, kids[k]) ? 0u : 1u;
    counts[i] = c;
}
// pass 2: write the BVH4 nodes of the level (node index level_base + i) and the next level's node list
__global__ void level_emit_kernel(TreeView t, const uint32_t *level, uint32_t n_level, uint32_t level_base, const uint32_t *offsets, uint32_t next_base,
                                  float4 *nodes, uint32_t *next_level) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_level) return;
    uint32_t kids[4];
    const int nk = gather_kids(t, level[i], kids);
    float lo[3][4], hi[3][4];
    uint32_t refs[4];
    uint32_t slot = offsets[i];
    for (int k = 0; k < 4; k++) {
        if (k < nk) {
            float4 l, h;
            child_box(t, kids[k], l, h);
            lo[0][k] = l.x; lo[1][k] = l.y; lo[2][k] = l.z;
            hi[0][k] = h.x; hi[1][k] = h.y; hi[2][k] = h.z;
            if (child_is_leaf(t, kids[k])) {
                const uint32_t first = (kids[k] & LEAF_BIT) ? (kids[k] & ~LEAF_BIT) : t.first[kids[k]];
                refs[k] = CHILD_LEAF | ((child_count(t, kids[k]) - 1u) << 27) | first;
            } else {
                refs[k] = next_base + slot;
                next_level[slot] = kids[k];
                slot++;
            }
        } else {
            for (int a = 0; a < 3; a++) {
                lo[a][k] = INFINITY;
                hi[a][k] = -INFINITY;
            }
            refs[k] = CHILD_EMPTY;
        }
    }
    float4 *np = nodes + 8 * (size_t)(level_base + i);
    for (int a = 0; a < 3; a++) {
        np[a] = make_float4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
        np[3 + a] = make_float4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
    }
    np[6] = make_float4(__uint_as_float(refs[0]), __uint_as_float(refs[1]), __uint_as_float(refs[2]), __uint_as_float(refs[3]));
    np[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// primitive records and stored normals in slot (= sorted) order: the float operations of rt_bvh.cpp
__global__ void records_kernel(const rt_sphere *sph, const rt_triangle *tri, uint32_t ns, uint32_t n, const uint32_t *vals, const unsigned char *transmissive,
                               float4 *prims, float4 *shade_n, unsigned long long *nan_any, unsigned long long *nan_tm) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const uint32_t p = vals[slot];
    float4 *q = prims + 4 * (size_t)slot;
    if (p < ns) {
        const rt_sphere s = sph[p];
        q[0] = make_float4(s.center[0], s.center[1], s.center[2], s.radius * s.radius);
        q[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        q[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        q[3] = make_float4(__uint_as_float(p), __uint_as_float(s.material | PRIM_SPHERE | (transmissive[s.material] ? PRIM_TRANSMISSIVE : 0u)), 0.0f, 0.0f);
        shade_n[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    } else {
        const rt_triangle *t = tri + (p - ns);
        float e1[3], e2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            e1[k] = __fsub_rn(t->v2[k], t->v1[k]);
            e2[k] = __fsub_rn(t->v3[k], t->v1[k]);
        }
        const float g0 = __fmaf_rn(e1[1], e2[2], -__fmul_rn(e1[2], e2[1])), g1 = __fmaf_rn(e1[2], e2[0], -__fmul_rn(e1[0], e2[2])),
                    g2 = __fmaf_rn(e1[0], e2[1], -__fmul_rn(e1[1], e2[0]));
        float emax = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) emax = fmaxf(emax, fmaxf(fabsf(e1[k]), fabsf(e2[k])));
        const uint32_t mat = t->material;
        const bool tm = transmissive[mat] != 0;
        q[0] = make_float4(t->v1[0], t->v1[1], t->v1[2], e1[0]);
        q[1] = make_float4(e1[1], e1[2], e2[0], e2[1]);
        q[2] = make_float4(e2[2], g0, g1, g2);
        q[3] = make_float4(__uint_as_float(p), __uint_as_float(mat | (tm ? PRIM_TRANSMISSIVE : 0u)), 0.0f, __fmul_rn(__fmul_rn(64.0f, 1.1920929e-7f), emax));
        shade_n[slot] = make_float4(t->normal[0], t->normal[1], t->normal[2], 0.0f);
        const unsigned long long key = ((unsigned long long)p << 32) | slot;
        atomicMin(nan_any, key);
        if (tm) atomicMin(nan_tm, key);
    }
}

template <typename T>
struct Tmp {  // device scratch that frees itself
    T *p = nullptr;
    ~Tmp() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t count) { return cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T)); }
};

}  // namespace

int build_scene_on_device(const rt_scene_desc *d, const std::vector<unsigned char> &transmissive, DeviceBvh &out) {
    out = DeviceBvh();
    const uint32_t ns = d->n_spheres, nt = d->n_triangles, n = ns + nt;
    if (n < 2) return rt_fail(RT_ERR_INVALID_ARGUMENT, "device BVH build needs at least two primitives");
    int rc = RT_OK;
    const bool verbose = getenv("RT_BVH_VERBOSE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto stage = [&](const char *name) {
        if (!verbose) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[rt_bvh_device] %-28s %8.2f ms\n", name, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    // every temporary lives in ONE allocation (a cudaMalloc of hundreds of MB costs milliseconds; two dozen of them
    // were a third of the build time)
    Tmp<unsigned char> arena;
    size_t arena_size = 0;
    auto reserve = [&](size_t bytes) {  // returns the offset of a 256-byte aligned block
        const size_t at = arena_size;
        arena_size += (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
        return at;
    };
    rt_sphere *d_sph = nullptr;
    rt_triangle *d_tri = nullptr;
    unsigned char *d_tm = nullptr, *cub_tmp = nullptr;
    Bounds *d_bounds = nullptr;
    float4 *blo = nullptr, *bhi = nullptr, *nlo = nullptr, *nhi = nullptr, *nodes_tmp = nullptr;
    unsigned long long *keys_in = nullptr, *keys_out = nullptr, *d_nan = nullptr;
    uint32_t *vals_in = nullptr, *vals_out = nullptr, *left = nullptr, *right = nullptr, *parent_int = nullptr, *parent_leaf = nullptr, *first = nullptr,
             *last = nullptr, *level_a = nullptr, *level_b = nullptr, *counts = nullptr, *offsets = nullptr;
    unsigned int *flags = nullptr;
    size_t sort_bytes = 0, scan_bytes = 0, tmp_bytes = 0;
    float4 *d_nodes = nullptr, *d_prims = nullptr, *d_shade = nullptr;
    Bounds hb;
    const int grid_n = (int)((n + TPB - 1) / TPB);
    {
        CUB_OK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (uint32_t *)nullptr,
                                               (uint32_t *)nullptr, (int)n, 0, 63));
        CUB_OK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n));
        tmp_bytes = std::max(sort_bytes, scan_bytes);
        const size_t o_sph = reserve((size_t)ns * sizeof(rt_sphere)), o_tri = reserve((size_t)nt * sizeof(rt_triangle)), o_tm = reserve(transmissive.size()),
                     o_bounds = reserve(sizeof(Bounds)), o_blo = reserve((size_t)n * 16), o_bhi = reserve((size_t)n * 16), o_nlo = reserve((size_t)n * 16),
                     o_nhi = reserve((size_t)n * 16), o_kin = reserve((size_t)n * 8), o_kout = reserve((size_t)n * 8), o_nan = reserve(16),
                     o_vin = reserve((size_t)n * 4), o_vout = reserve((size_t)n * 4), o_left = reserve((size_t)n * 4), o_right = reserve((size_t)n * 4),
                     o_pint = reserve((size_t)n * 4), o_pleaf = reserve((size_t)n * 4), o_first = reserve((size_t)n * 4), o_last = reserve((size_t)n * 4),
                     o_la = reserve((size_t)n * 4), o_lb = reserve((size_t)n * 4), o_counts = reserve((size_t)n * 4), o_offsets = reserve((size_t)(n + 1) * 4),
                     o_flags = reserve((size_t)n * 4), o_cub = reserve(tmp_bytes), o_nodes = reserve((size_t)n * 128);  // (fewer than n BVH4 nodes)
        CUB_OK(arena.alloc(arena_size));
        unsigned char *base = arena.p;
        d_sph = (rt_sphere *)(base + o_sph); d_tri = (rt_triangle *)(base + o_tri); d_tm = base + o_tm; d_bounds = (Bounds *)(base + o_bounds);
        blo = (float4 *)(base + o_blo); bhi = (float4 *)(base + o_bhi); nlo = (float4 *)(base + o_nlo); nhi = (float4 *)(base + o_nhi);
        keys_in = (unsigned long long *)(base + o_kin); keys_out = (unsigned long long *)(base + o_kout); d_nan = (unsigned long long *)(base + o_nan);
        vals_in = (uint32_t *)(base + o_vin); vals_out = (uint32_t *)(base + o_vout); left = (uint32_t *)(base + o_left); right = (uint32_t *)(base + o_right);
        parent_int = (uint32_t *)(base + o_pint); parent_leaf = (uint32_t *)(base + o_pleaf); first = (uint32_t *)(base + o_first); last = (uint32_t *)(base + o_last);
        level_a = (uint32_t *)(base + o_la); level_b = (uint32_t *)(base + o_lb); counts = (uint32_t *)(base + o_counts); offsets = (uint32_t *)(base + o_offsets);
        flags = (unsigned int *)(base + o_flags); cub_tmp = base + o_cub; nodes_tmp = (float4 *)(base + o_nodes);
        stage("scratch arena");
        if (ns) CUB_OK(cudaMemcpy(d_sph, d->spheres, (size_t)ns * sizeof(rt_sphere), cudaMemcpyHostToDevice));
        if (nt) CUB_OK(cudaMemcpy(d_tri, d->triangles, (size_t)nt * sizeof(rt_triangle), cudaMemcpyHostToDevice));
        if (!transmissive.empty()) CUB_OK(cudaMemcpy(d_tm, transmissive.data(), transmissive.size(), cudaMemcpyHostToDevice));
        stage("upload primitives");
        // ---- 1. bounds + validation
        Bounds init;
        for (int k = 0; k < 3; k++) {
            init.lo[k] = 0x7FFFFFFF;
            init.hi[k] = (int)0x80000000;
        }
        init.max_abs = 0;
        init.bad_coord = init.bad_material = 0;
        CUB_OK(cudaMemcpy(d_bounds, &init, sizeof(init), cudaMemcpyHostToDevice));
        bounds_kernel<<<std::min(grid_n, 148 * 8), TPB>>>(d_sph, d_tri, ns, n, d->n_materials, d_bounds);
        CUB_OK(cudaGetLastError());
        CUB_OK(cudaMemcpy(&hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost));
        if (hb.bad_coord) {
            rc = rt_fail(RT_ERR_INVALID_ARGUMENT, "a primitive has non-finite coordinates");
            goto done;
        }
        if (hb.bad_material) {
            rc = rt_fail(RT_ERR_INVALID_ARGUMENT, "a primitive's material index is out of range");
            goto done;
        }
        float slo[3], shi[3], max_abs = 1.0f;
        for (int k = 0; k < 3; k++) {
            slo[k] = ord2f(hb.lo[k]);
            shi[k] = ord2f(hb.hi[k]);
            max_abs = std::max(max_abs, std::max(std::fabs(slo[k]), std::fabs(shi[k])));
        }
        out.box_pad = 1e-4f * max_abs;  // rt_bvh.cpp
        stage("bounds");
        // ---- 2. references
        float3 lo3 = make_float3(slo[0], slo[1], slo[2]);
        float3 inv3 = make_float3(shi[0] > slo[0] ? 1.0f / (shi[0] - slo[0]) : 0.0f, shi[1] > slo[1] ? 1.0f / (shi[1] - slo[1]) : 0.0f,
                                  shi[2] > slo[2] ? 1.0f / (shi[2] - slo[2]) : 0.0f);
        refs_kernel<<<grid_n, TPB>>>(d_sph, d_tri, ns, n, out.box_pad, lo3, inv3, blo, bhi, keys_in, vals_in);
        CUB_OK(cudaGetLastError());
        stage("boxes + Morton codes");
        // ---- 3. sort
        {
            size_t sb = tmp_bytes;
            CUB_OK(cub::DeviceRadixSort::SortPairs(cub_tmp, sb, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 63));
        }
        stage("radix sort");
        // ---- 4. radix tree + boxes
        CUB_OK(cudaMemset(flags, 0, (size_t)n * sizeof(unsigned int)));
        radix_tree_kernel<<<grid_n, TPB>>>(keys_out, (int)n, left, right, parent_int, parent_leaf, first, last);
        CUB_OK(cudaGetLastError());
        refit_kernel<<<grid_n, TPB>>>((int)n, vals_out, blo, bhi, left, right, parent_int, parent_leaf, nlo, nhi, flags);
        CUB_OK(cudaGetLastError());
        stage("radix tree + refit");
        // ---- 5. collapse to BVH4, level by level (breadth-first numbering)
        TreeView tv{left, right, first, last, vals_out, blo, bhi, nlo, nhi};
        uint32_t root = 0;
        CUB_OK(cudaMemcpy(level_a, &root, sizeof(root), cudaMemcpyHostToDevice));
        uint32_t n_level = 1, level_base = 0, depth = 0;
        uint32_t *cur = level_a, *nxt = level_b;
        // (the root of a scene with <= MAX_LEAF primitives would be a leaf: such scenes never come here)
        while (n_level) {
            depth++;
            const int g = (int)((n_level + TPB - 1) / TPB);
            level_count_kernel<<<g, TPB>>>(tv, cur, n_level, counts);
            CUB_OK(cudaGetLastError());
            size_t sb = tmp_bytes;
            CUB_OK(cub::DeviceScan::ExclusiveSum(cub_tmp, sb, counts, offsets, (int)n_level));
            const uint32_t next_base = level_base + n_level;
            level_emit_kernel<<<g, TPB>>>(tv, cur, n_level, level_base, offsets, next_base, nodes_tmp, nxt);
            CUB_OK(cudaGetLastError());
            uint32_t last_off = 0, last_cnt = 0;
            CUB_OK(cudaMemcpy(&last_off, offsets + (n_level - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
            CUB_OK(cudaMemcpy(&last_cnt, counts + (n_level - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
            level_base = next_base;
            n_level = last_off + last_cnt;
            std::swap(cur, nxt);
            if (depth > 64) {
                rc = rt_fail(RT_ERR_CUDA, "device BVH build: collapse did not terminate");
                goto done;
            }
        }
        out.n_nodes = level_base;
        out.max_depth = depth;
        stage("collapse to BVH4");
        // ---- 6. final arrays
        CUB_OK(cudaMalloc((void **)&d_nodes, (size_t)out.n_nodes * 128));
        CUB_OK(cudaMemcpy(d_nodes, nodes_tmp, (size_t)out.n_nodes * 128, cudaMemcpyDeviceToDevice));
        CUB_OK(cudaMalloc((void **)&d_prims, (size_t)n * 64));
        CUB_OK(cudaMalloc((void **)&d_shade, (size_t)n * 16));
        CUB_OK(cudaMemset(d_nan, 0xFF, 2 * sizeof(unsigned long long)));
        records_kernel<<<grid_n, TPB>>>(d_sph, d_tri, ns, n, vals_out, d_tm, d_prims, d_shade, d_nan, d_nan + 1);
        CUB_OK(cudaGetLastError());
        unsigned long long hn[2];
        CUB_OK(cudaMemcpy(hn, d_nan, sizeof(hn), cudaMemcpyDeviceToHost));
        out.nan_hit_slot = hn[0] == ~0ull ? -1 : (int32_t)(hn[0] & 0xFFFFFFFFu);
        out.nan_hit_slot_culled = hn[1] == ~0ull ? -1 : (int32_t)(hn[1] & 0xFFFFFFFFu);
        out.n_slots = n;
        stage("primitive records");
        out.nodes = d_nodes;
        out.prims = d_prims;
        out.shade_n = d_shade;
        d_nodes = d_prims = d_shade = nullptr;
    }
done:
    if (d_nodes) cudaFree(d_nodes);
    if (d_prims) cudaFree(d_prims);
    if (d_shade) cudaFree(d_shade);
    return rc;
}

}  // namespace rtb
