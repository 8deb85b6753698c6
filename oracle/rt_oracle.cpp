// rt_oracle.cpp -- CPU ORACLE for the per-pixel render loop of
// kije/HSLU_I.BA_RAYTRACING.F2501_raytracer.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product (librt_b200.so)
// never links, imports or calls anything in oracle/.
//
// It restates the reference algorithm function by function (scalar, LANES = 1 semantics,
// recursion exactly as written, linear scan over every primitive).  Citations are
// reference-relative `file:line`.  Arithmetic is IEEE binary32, compiled with
// -ffp-contract=off; every fused multiply-add is an explicit fmaf() placed where the
// third-party crates (ultraviolet 0.10.0, simba 0.9.0, palette 0.7.6 -- NOT under
// /root/reference, see Cargo.lock) use `mul_add`.
//
// PARITY STATUS: "parity unpinned" at the ulp level -- the reference ships no test, golden
// vector or fixture for any function on this path, its third-party
// arithmetic cannot be read here, and it cannot be built here (no rustc/cargo).  What IS
// pinned: tests/golden/output_png_kats.json holds pixels of the reference's own committed
// render (/root/reference/output.png) that this oracle reproduces (tests/test_oracle_golden.py):
// 13 smooth-region pixels to the LSB, the hit mask on two grids (7505 + 1748 pixels, identical up
// to AA-dilated silhouettes), and a dense 1748-pixel grid under the full default feature set
// (62 % exact, 88 % within 1 LSB, mean |diff| 0.65 LSB -- the residue is the noise of the
// reference's unseeded samplers).
//
// A second mode (mode = 1) swaps the two linear scans for a private median-split BVH; it is
// proven bit-identical to mode 0 by tests/test_oracle_bvh.py and exists only so that
// full-frame oracle images are affordable.  A third mode (mode = 2) is the CPU-baseline stand-in
// for the reference's `simd_render` feature: the same linear scans, 8 rays per AVX2 packet
// against one object at a time (Vec3x8 instantiation, raytracer.rs:48-50,180-182), bit-identical
// to mode 0 per lane (tests/test_oracle_bvh.py); used only for timing by bench.py.
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// flat input structs (byte-compatible with include/rt_api.h so tests can feed both sides the
// same arrays; declared independently on purpose)
struct OMaterial {
    float color[3];
    float metallic, shininess, ior, opacity;
    uint32_t has_opacity;
    float boost;
};
struct OTriangle {
    float v1[3], v2[3], v3[3], normal[3];
    uint32_t material;
};
struct OSphere {
    float center[3];
    float radius;
    uint32_t material;
};
struct OLight {
    float position[3], color[3];
    float intensity;
};
struct OSceneDesc {
    const OTriangle *triangles;
    uint32_t n_triangles;
    const OSphere *spheres;
    uint32_t n_spheres;
    const OMaterial *materials;
    uint32_t n_materials;
    const OLight *lights;
    uint32_t n_lights;
};
struct OConfig {
    uint32_t width, height, flags, aa_samples_per_pixel;
    const float *aa_table;
    uint32_t aa_table_len, light_multiplicator;
    const float *light_clouds;
    uint32_t n_light_clouds, reflection_max_depth, refraction_max_depth;
    uint64_t seed;
    uint32_t tile_rows, tile_rank, tile_world, tables_epoch;
};
enum {
    F_AA = 1u << 0,
    F_AA_ROT = 1u << 1,
    F_REFL = 1u << 2,
    F_REFR = 1u << 3,
    F_CULL = 1u << 4,
    F_PACKET = 1u << 5,
};

constexpr float EPS = 1.1920929e-7f;          // f32::EPSILON (src/float_ext.rs:65-72)
constexpr float AIR_IOR = 1.000293f;          // src/lib.rs:92
const float INF = std::numeric_limits<float>::infinity();

// ------------------------------------------------------------------------------------------
// ultraviolet 0.10 Vec3 semantics (recalled)
struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 v3(const float *p) { return {p[0], p[1], p[2]}; }
inline V3 splat(float s) { return {s, s, s}; }
// dot = x.mul_add(ox, y.mul_add(oy, z * oz))
inline float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
// cross = (y.mul_add(oz, -z * oy), z.mul_add(ox, -x * oz), x.mul_add(oy, -y * ox))
inline V3 cross(V3 a, V3 b) {
    return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}
inline float mag(V3 a) { return sqrtf(dot(a, a)); }
// normalize: r = 1/mag; each component *= r
inline V3 normalized(V3 a) {
    float r = 1.0f / mag(a);
    return {a.x * r, a.y * r, a.z * r};
}
inline V3 mul_add(V3 a, V3 m, V3 add) { return {fmaf(a.x, m.x, add.x), fmaf(a.y, m.y, add.y), fmaf(a.z, m.z, add.z)}; }
// reflect: v -= (2 * v.dot(n)) * n
inline V3 reflected(V3 v, V3 n) {
    float s = 2.0f * dot(v, n);
    return v - n * s;
}
// refracted(i, n, eta): ndi = n.i; k = 1 - eta^2 (1 - ndi^2); k<0 ? 0 : i*eta - (eta*ndi + sqrt k) n
inline V3 refracted(V3 i, V3 n, float eta) {
    float ndi = dot(n, i);
    float k = 1.0f - (eta * eta) * (1.0f - ndi * ndi);
    if (k < 0.0f) return {0.0f, 0.0f, 0.0f};
    return i * eta - n * (eta * ndi + sqrtf(k));
}
inline bool is_nan3(V3 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }
// simd_max / simd_min / simd_clamp with the NaN behaviour of the packet path the reference's own
// render was made with (wide f32x8 -> vmaxps/vminps: `a > b ? a : b`, i.e. a NaN first operand
// yields the second): clamp(NaN, lo, hi) == lo.  Evidence: a NaN distance must attenuate to 0,
// otherwise every TIR refraction child (zero vector -> NaN direction) would turn its
// pixel into a constant colour, and /root/reference/output.png has no such speckles.
inline float smax(float a, float b) { return a > b ? a : b; }
inline float smin(float a, float b) { return a < b ? a : b; }
inline float sclamp(float x, float lo, float hi) { return smin(smax(x, lo), hi); }

// ------------------------------------------------------------------------------------------
struct Mat {  // Material<f32>
    V3 color;
    float metallic, shininess, ior, opacity;
    bool has_opacity;
    float boost;
    // TransmissionProperties::mask (src/raytracing/material.rs:44-50)
    bool transmissive() const { return has_opacity && !(fabsf(opacity - 0.0f) <= EPS); }
};

struct Prim {  // RenderGeometry<Vec3>: kind 0 = sphere, 1 = triangle
    int kind;
    V3 a, e1, e2, n;  // triangle: v1, edge1, edge2, stored normal; sphere: a = centre
    float r_sq;
    uint32_t mat;
};

struct Hit {  // SurfaceInteraction<Vec3> (src/raytracing/surface_interaction.rs:13-32)
    bool valid = false;
    float t = 0;
    V3 p{0, 0, 0}, n{0, 0, 0};
    uint32_t mat = 0;
    uint32_t prim = 0xFFFFFFFFu;
};

struct Counters {
    uint64_t closest = 0, shadow = 0, tri = 0, sph = 0, box = 0;
    uint64_t facing_away = 0;  // shadow queries of a hit whose surface faces away from the light (`light_valid` false: the pair adds nothing)
};

struct BvhNode {
    float lo[3], hi[3];
    int32_t left, right;   // children (internal) ...
    int32_t first, count;  // ... or leaf range into order[]
};

struct Scene {
    std::vector<Prim> prims;  // spheres then triangles (collection order)
    std::vector<Mat> mats;
    std::vector<OLight> lights;
    // private accelerator for mode 1
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> order;
};

// Ray::new_with_mask (src/geometry/ray.rs:43-58): direction is ALWAYS re-normalised
struct Ray {
    V3 o, d;
};
inline Ray make_ray(V3 o, V3 dir) { return {o, normalized(dir)}; }
inline V3 ray_at(const Ray &r, float t) { return mul_add(r.d, splat(t), r.o); }  // ray.rs:61-66

// TriangleData::intersect (src/geometry/basic/triangle.rs:149-212).
// Mat3::inversed / determinant / mat*vec follow ultraviolet 0.10 as recalled:
//   inversed: x = c1 x c2, y = c2 x c0, z = c0 x c1, det = c0 . x, rows scaled by 1/det, transposed
//   determinant: cofactor expansion along the first row, unfused
//   mat * v: a.x*v.x + b.x*v.y + c.x*v.z (unfused, left to right)
inline bool tri_intersect(const Prim &tr, const Mat &m, const Ray &ray, bool cull, Hit &out) {
    if (cull) {
        bool visible = (dot(ray.d, tr.n) < 0.75f) || m.transmissive();
        if (!visible) return false;
    }
    V3 b = tr.a - ray.o;
    V3 c0 = ray.d, c1 = -tr.e1, c2 = -tr.e2;
    V3 x = cross(c1, c2), y = cross(c2, c0), z = cross(c0, c1);
    float det = dot(c0, x);
    float inv_det = 1.0f / det;
    V3 X = x * inv_det, Y = y * inv_det, Z = z * inv_det;
    float t = X.x * b.x + X.y * b.y + X.z * b.z;
    float u = Y.x * b.x + Y.y * b.y + Y.z * b.z;
    float v = Z.x * b.x + Z.y * b.y + Z.z * b.z;
    float det2 = c0.x * (c1.y * c2.z - c2.y * c1.z) - c1.x * (c0.y * c2.z - c2.y * c0.z) +
                 c2.x * (c0.y * c1.z - c1.y * c0.z);
    bool t_invalid = t <= EPS;
    bool uv_invalid = (u < 0.0f) || (v < 0.0f) || ((u + v) >= 1.0f);
    bool det_zero = fabsf(det2 - 0.0f) <= EPS;
    bool valid = !(t_invalid || uv_invalid) && !det_zero;
    if (!valid) return false;
    out.valid = true;
    out.t = t;
    out.p = ray_at(ray, t);
    out.n = tr.n;
    out.mat = tr.mat;
    return true;
}

// SphereData::intersect (src/geometry/basic/sphere.rs:78-162). Returns "Some" (true) whenever
// the discriminant is >= 0; out.valid carries final_t_valid & is_visible_face.
inline bool sph_intersect(const Prim &sp, const Mat &m, const Ray &ray, bool cull, Hit &out) {
    V3 u = ray.d;
    V3 v = ray.o - sp.a;
    float b = 2.0f * dot(u, v);
    float c = dot(v, v) - sp.r_sq;
    float disc = fmaf(b, b, (2.0f * -2.0f) * c);
    if (!(disc >= 0.0f)) return false;
    float s = sqrtf(disc);
    float mb = -b;
    float mba = mb * 0.5f;
    float sa = s * 0.5f;
    float t0 = mba - sa, t1 = mba + sa;
    bool t0v = t0 >= 0.0f, t1v = t1 >= 0.0f;
    bool use0 = t0v && (!t1v || (t0 < t1));
    bool use1 = t1v && !use0;
    float t = INF;
    if (use0) t = t0;
    if (use1) t = t1;
    bool tvalid = use0 || use1;
    V3 p = ray_at(ray, t);
    V3 n = normalized(p - sp.a);
    bool visible = true;
    if (cull) {
        visible = (dot(ray.d, n) < 0.75f) || m.transmissive();
        if (!visible) return false;
    }
    out.valid = visible && tvalid;
    out.t = t;
    out.p = p;
    out.n = n;
    out.mat = sp.mat;
    return true;
}

inline bool prim_intersect(const Scene &sc, uint32_t i, const Ray &ray, bool cull, Hit &h, Counters &cn) {
    const Prim &p = sc.prims[i];
    h.valid = false;
    bool some;
    if (p.kind == 0) {
        cn.sph++;
        some = sph_intersect(p, sc.mats[p.mat], ray, cull, h);
    } else {
        cn.tri++;
        some = tri_intersect(p, sc.mats[p.mat], ray, cull, h);
    }
    h.prim = i;
    return some;
}

// Material::compute_fresnel (src/raytracing/material.rs:468-525)
inline void compute_fresnel(const Mat &m, V3 normal, V3 view, float other_ior, V3 &refl, V3 &trans) {
    if (!m.transmissive()) {
        refl = splat(1.0f) * m.metallic;
        trans = splat(0.0f);
        return;
    }
    bool is_reflective = m.metallic > 0.0f;
    float ior = m.ior;
    float ndv = dot(normal, view);
    float cos_theta = fabsf(ndv);
    bool inside = ndv < 0.0f;
    float eta_t = inside ? (ior / other_ior) : (other_ior / ior);
    float sin2_t = eta_t * eta_t * (1.0f - cos_theta * cos_theta);
    bool tir = (inside && (sin2_t > 1.0f)) || is_reflective;
    float f0s = (other_ior - ior) / (other_ior + ior);
    f0s = f0s * f0s;  // powi(2)
    // ultraviolet Lerp: a * (1 - t) + b * t
    V3 f0 = splat(f0s) * (1.0f - m.metallic) + m.color * m.metallic;
    float omc = 1.0f - cos_theta;
    float p2 = omc * omc;
    float p5 = omc * (p2 * p2);  // powi(5) by squaring
    V3 fres = f0 + (splat(1.0f) - f0) * splat(p5);
    V3 refl_amt = is_reflective ? splat(1.0f) * m.metallic : splat(1.0f);
    V3 f = tir ? refl_amt : fres;
    refl = f;
    trans = splat(1.0f) - f;
}

// Material::absorption (material.rs:213-231)
inline V3 absorption(const Mat &m) {
    float op = m.transmissive() ? m.opacity : 1.0f;
    op = sclamp(op, 0.0f, 1.0f - EPS);
    return m.color * (1.0f - op);
}

struct Shadow {  // IntersectionTest (src/raytracing/raytracer.rs:17-22)
    bool occluded = false;
    float opacity = 1.0f;
    V3 filter{1.0f, 1.0f, 1.0f};
};

// body of the loop in Raytracer::has_any_intersection (raytracer.rs:48-97) for one object
inline void shadow_accumulate(const Scene &sc, const Ray &ray, const Hit &h, float max_distance, Shadow &s) {
    bool has = h.valid && (h.t <= max_distance);
    if (!has) return;
    const Mat &m = sc.mats[h.mat];
    V3 refl, trans;
    compute_fresnel(m, h.n, -ray.d, 1.0f, refl, trans);
    bool tm = m.transmissive();
    if (!tm) trans = splat(0.0f);
    float io = (tm ? m.opacity : 0.0f) * trans.x;
    s.opacity = sclamp(s.opacity - (1.0f - io), 0.0f, 1.0f);
    bool occ = !tm && (fabsf(s.opacity - 0.0f) <= EPS);
    s.occluded = s.occluded || occ;
    s.filter = s.filter - absorption(m);
}

// ------------------------------------------------------------------------------------------
// private BVH (mode 1). Boxes are padded far beyond any float error of the primitive tests
// and the slab test is widened, so it can only ever visit MORE primitives than necessary.
inline void prim_bounds(const Prim &p, float lo[3], float hi[3]) {
    if (p.kind == 0) {
        float r = sqrtf(p.r_sq);
        const float c[3] = {p.a.x, p.a.y, p.a.z};
        for (int k = 0; k < 3; k++) {
            lo[k] = c[k] - r;
            hi[k] = c[k] + r;
        }
    } else {
        V3 b = p.a + p.e1, c = p.a + p.e2;
        const float A[3] = {p.a.x, p.a.y, p.a.z}, B[3] = {b.x, b.y, b.z}, C[3] = {c.x, c.y, c.z};
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(A[k], std::min(B[k], C[k]));
            hi[k] = std::max(A[k], std::max(B[k], C[k]));
        }
    }
    for (int k = 0; k < 3; k++) {
        float pad = 1e-3f * (hi[k] - lo[k]) + 2e-4f * std::max(1.0f, std::max(fabsf(lo[k]), fabsf(hi[k])));
        lo[k] -= pad;
        hi[k] += pad;
    }
}

int build_node(Scene &sc, std::vector<float> &blo, std::vector<float> &bhi, int first, int count) {
    BvhNode node{};
    for (int k = 0; k < 3; k++) {
        node.lo[k] = INF;
        node.hi[k] = -INF;
    }
    float clo[3] = {INF, INF, INF}, chi[3] = {-INF, -INF, -INF};
    for (int i = first; i < first + count; i++) {
        uint32_t p = sc.order[i];
        for (int k = 0; k < 3; k++) {
            node.lo[k] = std::min(node.lo[k], blo[3 * p + k]);
            node.hi[k] = std::max(node.hi[k], bhi[3 * p + k]);
            float c = 0.5f * (blo[3 * p + k] + bhi[3 * p + k]);
            clo[k] = std::min(clo[k], c);
            chi[k] = std::max(chi[k], c);
        }
    }
    node.left = node.right = -1;
    node.first = first;
    node.count = count;
    int idx = (int)sc.nodes.size();
    sc.nodes.push_back(node);
    if (count <= 4) return idx;
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    int mid = first + count / 2;
    std::nth_element(sc.order.begin() + first, sc.order.begin() + mid, sc.order.begin() + first + count,
                     [&](uint32_t a, uint32_t b) {
                         return blo[3 * a + axis] + bhi[3 * a + axis] < blo[3 * b + axis] + bhi[3 * b + axis];
                     });
    int l = build_node(sc, blo, bhi, first, mid - first);
    int r = build_node(sc, blo, bhi, mid, first + count - mid);
    sc.nodes[idx].left = l;
    sc.nodes[idx].right = r;
    sc.nodes[idx].count = 0;
    return idx;
}

void build_bvh(Scene &sc) {
    size_t n = sc.prims.size();
    sc.order.resize(n);
    std::vector<float> blo(3 * n), bhi(3 * n);
    for (size_t i = 0; i < n; i++) {
        sc.order[i] = (uint32_t)i;
        prim_bounds(sc.prims[i], &blo[3 * i], &bhi[3 * i]);
    }
    sc.nodes.clear();
    if (n) build_node(sc, blo, bhi, 0, (int)n);
}

// widened slab test: true if the ray may touch the box anywhere on t in [-inf, tmax]
inline bool slab(const BvhNode &nd, const Ray &r, const float inv[3], float tmax) {
    const float o[3] = {r.o.x, r.o.y, r.o.z};
    float tn = -INF, tf = INF;
    for (int k = 0; k < 3; k++) {
        float t0 = (nd.lo[k] - o[k]) * inv[k];
        float t1 = (nd.hi[k] - o[k]) * inv[k];
        if (std::isnan(t0) || std::isnan(t1)) continue;  // 0 * inf: origin on a slab plane
        if (t0 > t1) std::swap(t0, t1);
        tn = std::max(tn, t0);
        tf = std::min(tf, t1);
    }
    float slack = 1e-5f * (fabsf(tn) + fabsf(tf)) + 1e-6f;
    if (tn - slack > tf + slack) return false;
    if (tf + slack < 0.0f) return false;
    if (tn - slack > tmax) return false;
    return true;
}

// ------------------------------------------------------------------------------------------
struct Tracer {
    const Scene &sc;
    const OConfig &cfg;
    int mode;
    Counters cn;
    bool cull;
    Tracer(const Scene &s, const OConfig &c, int m) : sc(s), cfg(c), mode(m), cull((c.flags & F_CULL) != 0) {}

    // Raytracer::cast_ray (src/raytracing/raytracer.rs:162-220), scalar instantiation
    static inline void nearest_update(Hit &cur, const Hit &h) {
        if (!h.valid) return;  // `if interaction.valid_mask.none() continue`
        if (!cur.valid) {
            cur = h;
            return;
        }
        bool closer = (h.t <= cur.t) && h.valid;
        if (closer) cur = h;  // ties: the LATER object wins
    }

    Hit cast_ray(const Ray &ray) {
        cn.closest++;
        Hit cur;
        bool nanray = is_nan3(ray.o) || is_nan3(ray.d);
        if (mode == 0 || nanray || sc.nodes.empty()) {
            Hit h;
            for (uint32_t i = 0; i < sc.prims.size(); i++) {
                if (prim_intersect(sc, i, ray, cull, h, cn)) nearest_update(cur, h);
            }
            return cur;
        }
        const float inv[3] = {1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
        int stack[128];
        int sp = 0;
        stack[sp++] = 0;
        Hit h;
        while (sp) {
            const BvhNode &nd = sc.nodes[stack[--sp]];
            cn.box++;
            if (!slab(nd, ray, inv, cur.valid ? cur.t : INF)) continue;
            if (nd.left < 0) {
                for (int i = nd.first; i < nd.first + nd.count; i++) {
                    uint32_t pi = sc.order[i];
                    if (prim_intersect(sc, pi, ray, cull, h, cn) && h.valid) {
                        // same decision as the in-order scan: smaller t wins, equal t -> later index
                        if (!cur.valid || h.t < cur.t || (h.t == cur.t && pi > cur.prim)) cur = h;
                    }
                }
            } else {
                stack[sp++] = nd.left;
                stack[sp++] = nd.right;
            }
        }
        return cur;
    }

    // Raytracer::has_any_intersection (raytracer.rs:24-106), scalar instantiation
    Shadow has_any_intersection(V3 from, V3 direction, float max_distance) {
        cn.shadow++;
        Ray ray = make_ray(from, direction);
        Shadow s;
        bool nanray = is_nan3(ray.o) || is_nan3(ray.d);
        if (mode == 0 || nanray || sc.nodes.empty()) {
            Hit h;
            for (uint32_t i = 0; i < sc.prims.size(); i++) {
                if (prim_intersect(sc, i, ray, cull, h, cn)) {
                    shadow_accumulate(sc, ray, h, max_distance, s);
                    if (s.occluded) break;
                }
            }
            return s;
        }
        // BVH mode: collect every hit, then replay them in primitive order so the float
        // accumulation is bit-identical to the linear scan.
        const float inv[3] = {1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
        int stack[128];
        int sp = 0;
        stack[sp++] = 0;
        std::vector<Hit> hits;
        Hit h;
        while (sp) {
            const BvhNode &nd = sc.nodes[stack[--sp]];
            cn.box++;
            if (!slab(nd, ray, inv, max_distance)) continue;
            if (nd.left < 0) {
                for (int i = nd.first; i < nd.first + nd.count; i++) {
                    uint32_t pi = sc.order[i];
                    if (prim_intersect(sc, pi, ray, cull, h, cn) && h.valid && h.t <= max_distance) hits.push_back(h);
                }
            } else {
                stack[sp++] = nd.left;
                stack[sp++] = nd.right;
            }
        }
        std::sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.prim < b.prim; });
        for (const Hit &hh : hits) {
            shadow_accumulate(sc, ray, hh, max_distance, s);
            if (s.occluded) break;
        }
        return s;
    }

    // ------------------------------------------------------------------------------------------
    // mode 2: 8-ray packets (the `simd_render` instantiation of the same two loops)
    struct V8 {
        __m256 x, y, z;
    };
    static inline __m256 dot8(const V8 &a, const V8 &b) { return _mm256_fmadd_ps(a.x, b.x, _mm256_fmadd_ps(a.y, b.y, _mm256_mul_ps(a.z, b.z))); }
    static inline V8 splat8(V3 v) { return {_mm256_set1_ps(v.x), _mm256_set1_ps(v.y), _mm256_set1_ps(v.z)}; }
    static inline V8 cross8(const V8 &a, const V8 &b) {
        const __m256 sign = _mm256_set1_ps(-0.0f);
        return {_mm256_fmadd_ps(a.y, b.z, _mm256_xor_ps(sign, _mm256_mul_ps(a.z, b.y))),
                _mm256_fmadd_ps(a.z, b.x, _mm256_xor_ps(sign, _mm256_mul_ps(a.x, b.z))),
                _mm256_fmadd_ps(a.x, b.y, _mm256_xor_ps(sign, _mm256_mul_ps(a.y, b.x)))};
    }
    static inline V8 scale8(const V8 &a, __m256 s) { return {_mm256_mul_ps(a.x, s), _mm256_mul_ps(a.y, s), _mm256_mul_ps(a.z, s)}; }
    // a.x*b.x + a.y*b.y + a.z*b.z, unfused, left to right (Mat3 * Vec3)
    static inline __m256 mv8(const V8 &a, const V8 &b) {
        return _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(a.x, b.x), _mm256_mul_ps(a.y, b.y)), _mm256_mul_ps(a.z, b.z));
    }

    // TriangleData::intersect for 8 rays: lane mask of valid hits and their t
    inline int tri_intersect8(const Prim &tr, const Mat &m, const V8 &o, const V8 &d, __m256 &t_out) {
        cn.tri += 8;
        const __m256 sign = _mm256_set1_ps(-0.0f), eps = _mm256_set1_ps(EPS), zero = _mm256_setzero_ps(), one = _mm256_set1_ps(1.0f);
        __m256 visible = _mm256_castsi256_ps(_mm256_set1_epi32(-1));
        if (cull) {
            visible = _mm256_cmp_ps(dot8(d, splat8(tr.n)), _mm256_set1_ps(0.75f), _CMP_LT_OQ);
            if (m.transmissive()) visible = _mm256_castsi256_ps(_mm256_set1_epi32(-1));
            if (_mm256_movemask_ps(visible) == 0) return 0;
        }
        const V8 a = splat8(tr.a);
        const V8 b = {_mm256_sub_ps(a.x, o.x), _mm256_sub_ps(a.y, o.y), _mm256_sub_ps(a.z, o.z)};
        const V8 c1 = splat8(-tr.e1), c2 = splat8(-tr.e2);
        const V8 x = cross8(c1, c2), y = cross8(c2, d), z = cross8(d, c1);
        const __m256 det = dot8(d, x);
        const __m256 inv = _mm256_div_ps(one, det);
        const V8 X = scale8(x, inv), Y = scale8(y, inv), Z = scale8(z, inv);
        const __m256 t = mv8(X, b), u = mv8(Y, b), v = mv8(Z, b);
        // determinant: cofactor expansion along the first row, unfused
        const __m256 m0 = _mm256_sub_ps(_mm256_mul_ps(c1.y, c2.z), _mm256_mul_ps(c2.y, c1.z));
        const __m256 m1 = _mm256_sub_ps(_mm256_mul_ps(d.y, c2.z), _mm256_mul_ps(c2.y, d.z));
        const __m256 m2 = _mm256_sub_ps(_mm256_mul_ps(d.y, c1.z), _mm256_mul_ps(c1.y, d.z));
        const __m256 det2 = _mm256_add_ps(_mm256_sub_ps(_mm256_mul_ps(d.x, m0), _mm256_mul_ps(c1.x, m1)), _mm256_mul_ps(c2.x, m2));
        const __m256 t_inv = _mm256_cmp_ps(t, eps, _CMP_LE_OQ);
        const __m256 uv_inv = _mm256_or_ps(_mm256_or_ps(_mm256_cmp_ps(u, zero, _CMP_LT_OQ), _mm256_cmp_ps(v, zero, _CMP_LT_OQ)),
                                           _mm256_cmp_ps(_mm256_add_ps(u, v), one, _CMP_GE_OQ));
        const __m256 det_zero = _mm256_cmp_ps(_mm256_andnot_ps(sign, _mm256_sub_ps(det2, zero)), eps, _CMP_LE_OQ);
        const __m256 valid = _mm256_andnot_ps(det_zero, _mm256_andnot_ps(_mm256_or_ps(t_inv, uv_inv), visible));
        t_out = t;
        return _mm256_movemask_ps(valid);
    }

    // Raytracer::cast_ray, Vec3x8 instantiation: every object against the 8-ray packet
    void cast_ray8(const Ray rays[8], Hit out[8]) {
        cn.closest += 8;
        float buf[6][8];
        for (int l = 0; l < 8; l++) {
            buf[0][l] = rays[l].o.x; buf[1][l] = rays[l].o.y; buf[2][l] = rays[l].o.z;
            buf[3][l] = rays[l].d.x; buf[4][l] = rays[l].d.y; buf[5][l] = rays[l].d.z;
        }
        const V8 o = {_mm256_loadu_ps(buf[0]), _mm256_loadu_ps(buf[1]), _mm256_loadu_ps(buf[2])};
        const V8 d = {_mm256_loadu_ps(buf[3]), _mm256_loadu_ps(buf[4]), _mm256_loadu_ps(buf[5])};
        bool cur_valid[8] = {false};
        float cur_t[8] = {0};
        uint32_t cur_prim[8] = {0};
        Hit h;
        for (uint32_t i = 0; i < sc.prims.size(); i++) {
            const Prim &p = sc.prims[i];
            if (p.kind == 0) {  // few spheres: lane by lane
                for (int l = 0; l < 8; l++) {
                    h.valid = false;
                    cn.sph++;
                    if (sph_intersect(p, sc.mats[p.mat], rays[l], cull, h) && h.valid && (!cur_valid[l] || h.t <= cur_t[l])) {
                        cur_valid[l] = true;
                        cur_t[l] = h.t;
                        cur_prim[l] = i;
                    }
                }
                continue;
            }
            __m256 t8;
            int mask = tri_intersect8(p, sc.mats[p.mat], o, d, t8);
            if (!mask) continue;
            float tt[8];
            _mm256_storeu_ps(tt, t8);
            for (int l = 0; l < 8; l++) {
                if (!(mask >> l & 1)) continue;
                if (!cur_valid[l] || tt[l] <= cur_t[l]) {  // ties: the later object wins; a NaN t never replaces
                    cur_valid[l] = true;
                    cur_t[l] = tt[l];
                    cur_prim[l] = i;
                }
            }
        }
        Counters scratch;
        for (int l = 0; l < 8; l++) {
            out[l] = Hit();
            if (!cur_valid[l]) continue;
            // rebuild the full interaction with the scalar code (same arithmetic, same result)
            std::swap(cn, scratch);
            prim_intersect(sc, cur_prim[l], rays[l], cull, out[l], cn);
            std::swap(cn, scratch);
        }
    }

    // Raytracer::has_any_intersection, Vec3x8 instantiation
    void has_any_intersection8(const V3 from[8], const V3 dir[8], const float maxd[8], Shadow out[8]) {
        cn.shadow += 8;
        Ray rays[8];
        float buf[7][8];
        for (int l = 0; l < 8; l++) {
            rays[l] = make_ray(from[l], dir[l]);
            buf[0][l] = rays[l].o.x; buf[1][l] = rays[l].o.y; buf[2][l] = rays[l].o.z;
            buf[3][l] = rays[l].d.x; buf[4][l] = rays[l].d.y; buf[5][l] = rays[l].d.z;
            buf[6][l] = maxd[l];
            out[l] = Shadow();
        }
        const V8 o = {_mm256_loadu_ps(buf[0]), _mm256_loadu_ps(buf[1]), _mm256_loadu_ps(buf[2])};
        const V8 d = {_mm256_loadu_ps(buf[3]), _mm256_loadu_ps(buf[4]), _mm256_loadu_ps(buf[5])};
        const __m256 md = _mm256_loadu_ps(buf[6]);
        Counters scratch;
        Hit h;
        for (uint32_t i = 0; i < sc.prims.size(); i++) {
            const Prim &p = sc.prims[i];
            int mask = 0;
            if (p.kind == 0) {
                mask = 0xFF;  // resolved lane by lane below
            } else {
                __m256 t8;
                mask = tri_intersect8(p, sc.mats[p.mat], o, d, t8);
                mask &= _mm256_movemask_ps(_mm256_cmp_ps(t8, md, _CMP_LE_OQ));
            }
            if (!mask) continue;
            for (int l = 0; l < 8; l++) {
                if (!(mask >> l & 1) || out[l].occluded) continue;
                std::swap(cn, scratch);
                bool some = prim_intersect(sc, i, rays[l], cull, h, cn);
                std::swap(cn, scratch);
                if (p.kind == 0) cn.sph++;
                if (some) shadow_accumulate(sc, rays[l], h, maxd[l], out[l]);
            }
            // `if completely_occluded.all() && has_intersection.all() { break }` (raytracer.rs:94-96)
            bool all = true;
            for (int l = 0; l < 8; l++) all = all && out[l].occluded;
            if (all) break;
        }
    }
};

// ------------------------------------------------------------------------------------------
// crate constants (src/lib.rs:73-92), evaluated in f32 like the Rust const expressions
struct Consts {
    float aspect, scene_w, scene_h, scene_d, avg_dim, w2s_w, w2s_h, w2s_d;
    V3 focus;
    float eps_d;
    explicit Consts(uint32_t W, uint32_t H) {
        aspect = (float)H / (float)W;
        scene_w = 1.0f;
        scene_h = scene_w * aspect;
        scene_d = (scene_w + scene_h) / 2.0f;
        avg_dim = (scene_w + scene_h + scene_d) / 3.0f;
        w2s_w = scene_w / (float)W;
        w2s_h = scene_h / (float)H;
        w2s_d = scene_d / (float)((W + H) / 2);
        focus = {scene_w / 2.0f, scene_h / 2.0f, -1.9f * scene_d};
        eps_d = EPS * (100.0f * avg_dim);  // Vector3DOperations::default_epsilon_distance (src/vector.rs:697-700)
    }
};

struct PLight {  // PointLight<Vec3> after preprocess()
    V3 pos, color;
    float intensity;
};

// identical integer hash in the product (rt_cloud_index)
inline uint32_t cloud_index(uint64_t seed, uint32_t x, uint32_t y, uint32_t light, uint32_t n) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(x + 1u) + 0xBF58476D1CE4E5B9ull * (uint64_t)(y + 1u) +
                 0x94D049BB133111EBull * (uint64_t)(light + 1u);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(((z >> 32) * (uint64_t)n) >> 32);
}

struct Renderer {
    const Scene &sc;
    const OConfig &cfg;
    Consts k;
    Tracer tr;
    bool refl_on, refr_on;
    Renderer(const Scene &s, const OConfig &c, int mode)
        : sc(s), cfg(c), k(c.width, c.height), tr(s, c, mode), refl_on((c.flags & F_REFL) != 0), refr_on((c.flags & F_REFR) != 0) {}

    // attenuation_factor_based_on_distance (src/renderer/raytracer_renderer.rs:266-277)
    static float attenuation(float distance) {
        float d = fabsf(distance);
        float a = 1.0f / (1.0f + d + 0.1f * d * d);
        return sclamp(a, 0.0f, 1.0f);
    }

    // calculate_lighting (raytracer_renderer.rs:731-874) + PointLight::calculate_contribution_at
    // (src/scene/lighting/light.rs:261-299)
    void calculate_lighting(const Hit &it, V3 view_dir, const std::vector<PLight> &lights, V3 &direct, V3 &specular) {
        const Mat &m = sc.mats[it.mat];
        V3 material_color = m.color;
        V3 ambient = (material_color * splat(1.0f)) * 0.08f;  // AmbientLight::new(one, 0.08), maximize_value(white) = white
        V3 light_color = splat(0.0f);
        V3 specular_color = splat(0.0f);
        bool has_specular = m.shininess > 0.0f;
        for (const PLight &l : lights) {
            V3 light_to_point = l.pos - it.p;
            V3 light_dir = normalized(light_to_point);
            V3 check_point = it.p + light_dir * k.eps_d;
            V3 check_l2p = l.pos - check_point;
            if (!(smax(dot(it.n, light_dir), 0.0f) > 0.0f)) tr.cn.facing_away++;  // (statistics only: the query is made all the same)
            Shadow sh = tr.has_any_intersection(check_point, light_dir, mag(check_l2p));
            if (sh.occluded) continue;
            // PointLight::calculate_contribution_at
            float light_distance = mag(light_to_point) + EPS;
            float cosi = dot(light_to_point, it.n) / light_distance;
            bool pos = cosi > 0.0f;
            float att = 0.95f * (EPS + light_distance + light_distance * light_distance);
            float sig = (tanhf(att) + 1.0f) / 2.0f;
            float light_factor_c = cosi * l.intensity * sclamp(sig, 0.0f, 1.0f);
            V3 ccolor = pos ? (material_color * l.color) : splat(0.0f);
            float cintensity = pos ? light_factor_c : 0.0f;
            // back in calculate_lighting
            V3 light_color_simd = ccolor / sh.filter;
            float diffuse = smax(dot(it.n, light_dir), 0.0f);
            float spec_factor = 0.0f;
            if (has_specular) {
                V3 r = reflected(light_dir, it.n);
                float sd = smax(dot(normalized(r), view_dir), 0.0f);
                spec_factor = powf(sd, smax(m.shininess * 512.0f, 1.0f));
            }
            float lf = diffuse * cintensity * sh.opacity;
            float sf = cintensity * sh.opacity * spec_factor;
            bool light_valid = diffuse > 0.0f;
            V3 diffuse_contribution = material_color * light_color_simd * lf;
            V3 specular_contribution = l.color * sf;
            if (light_valid) light_color = light_color + diffuse_contribution;
            if (has_specular && light_valid) specular_color = specular_color + specular_contribution;
        }
        direct = ambient + light_color;
        specular = specular_color;
    }

    struct Result {
        bool valid = false;
        V3 color{0, 0, 0};
        float t = 0;
    };

    // single_raytrace (raytracer_renderer.rs:147-264). depth < 0 encodes `None`.
    Result single_raytrace(V3 coords, V3 direction, float start_ior, const std::vector<PLight> &lights, int depth) {
        Result res;
        if (depth == 0) return res;
        Ray ray = make_ray(coords, direction);
        Hit it = tr.cast_ray(ray);
        if (!it.valid) return res;
        const Mat &m = sc.mats[it.mat];
        V3 direct, specular;
        calculate_lighting(it, ray.d, lights, direct, specular);
        float df = attenuation(it.t);
        direct = direct * df;
        specular = specular * df;
        bool is_transmissive = m.transmissive();
        bool is_reflective = (m.metallic > 0.0f) || is_transmissive;
        V3 reflection = splat(0.0f), refraction = splat(0.0f);
        if (refl_on && is_reflective) reflection = calculate_reflection(it, ray.d, start_ior, lights, depth);
        if (refr_on && is_transmissive) refraction = calculate_refractions(it, ray.d, start_ior, lights, depth);
        res.valid = true;
        res.t = it.t;
        res.color = is_transmissive ? (reflection + refraction + specular) : (direct + reflection + specular);
        return res;
    }

    // calculate_reflection (raytracer_renderer.rs:526-729)
    V3 calculate_reflection(const Hit &it, V3 view_dir, float start_ior, const std::vector<PLight> &lights, int depth) {
        const Mat &m = sc.mats[it.mat];
        float cos_theta = dot(view_dir, it.n);
        bool inside = cos_theta < 0.0f;
        V3 inormal = inside ? -it.n : it.n;
        float new_ior = inside ? m.ior : AIR_IOR;
        float eta = inside ? (new_ior / start_ior) : (start_ior / new_ior);
        float ci = fabsf(cos_theta);
        float sin2 = eta * eta * (1.0f - ci * ci);
        bool tir = sin2 >= 1.0f;
        bool reflective = (m.metallic > 0.0f) || (m.transmissive() && tir);
        if (!reflective) return splat(0.0f);
        V3 rdir = normalized(reflected(view_dir, it.n));
        bool none = (fabsf(rdir.x - 0.0f) <= EPS) && (fabsf(rdir.y - 0.0f) <= EPS) && (fabsf(rdir.z - 0.0f) <= EPS);
        if (none) return splat(0.0f);
        V3 reflectance, tmp;
        compute_fresnel(m, inormal, -view_dir, start_ior, reflectance, tmp);
        int next = depth < 0 ? (int)cfg.reflection_max_depth : std::max(depth - 1, 0);
        Result c = single_raytrace(it.p + rdir * k.eps_d, rdir, start_ior, lights, next);
        if (!c.valid) return splat(0.0f);
        float df = attenuation(c.t);
        return c.color * df * reflectance;
    }

    // calculate_refractions (raytracer_renderer.rs:279-524)
    V3 calculate_refractions(const Hit &it, V3 view_dir, float start_ior, const std::vector<PLight> &lights, int depth) {
        const Mat &m = sc.mats[it.mat];
        if (!m.transmissive()) return splat(0.0f);
        float cos_theta = dot(view_dir, it.n);
        bool entering = cos_theta <= 0.0f;
        V3 inormal = entering ? -it.n : it.n;
        float new_ior = entering ? m.ior : AIR_IOR;
        float eta = entering ? (new_ior / start_ior) : (start_ior / new_ior);
        float inv_eta = 1.0f / eta;
        V3 tmp, transmittance;
        compute_fresnel(m, inormal, view_dir, inv_eta, tmp, transmittance);
        V3 rdir = normalized(refracted(view_dir, -inormal, inv_eta));
        float op = m.opacity;  // transmissive => opacity().simd_unwrap_or(0) == value
        int step = (op < 0.5f) ? 2 : 1;
        int factor = (op <= 0.3f) ? 3 : ((op < 0.5f) ? 2 : 1);
        int next = depth < 0 ? (int)cfg.refraction_max_depth / factor : std::max(depth - step, 0);
        Result c = single_raytrace(it.p + rdir * k.eps_d, rdir, new_ior, lights, next);
        if (!c.valid) return splat(0.0f);
        float boost = m.boost + 1.0f;
        return c.color * boost * transmittance;
    }

    // single_raytrace + calculate_lighting for one 8-ray packet without bounces (mode 2 timing
    // path). Per lane identical to the scalar functions above; lanes that missed keep tracing
    // shadow packets with non-finite rays like the reference's masked lanes do.
    void single_raytrace8(const Ray rays[8], const std::vector<PLight> &lights, Result out[8]) {
        Hit hit[8];
        tr.cast_ray8(rays, hit);
        bool any = false;
        for (int l = 0; l < 8; l++) any = any || hit[l].valid;
        for (int l = 0; l < 8; l++) out[l] = Result();
        if (!any) return;  // `if nearest_interaction.valid_mask.none() { return None }`
        V3 light_color[8], specular_color[8];
        for (int l = 0; l < 8; l++) light_color[l] = specular_color[l] = splat(0.0f);
        const float nanv = std::numeric_limits<float>::quiet_NaN();
        for (const PLight &lt : lights) {
            V3 from[8], dirs[8], l2p[8], ldir[8];
            float maxd[8];
            for (int l = 0; l < 8; l++) {
                if (!hit[l].valid) {
                    from[l] = dirs[l] = l2p[l] = ldir[l] = splat(nanv);
                    maxd[l] = nanv;
                    continue;
                }
                l2p[l] = lt.pos - hit[l].p;
                ldir[l] = normalized(l2p[l]);
                from[l] = hit[l].p + ldir[l] * k.eps_d;
                dirs[l] = ldir[l];
                maxd[l] = mag(lt.pos - from[l]);
            }
            Shadow sh[8];
            tr.has_any_intersection8(from, dirs, maxd, sh);
            for (int l = 0; l < 8; l++)
                if (hit[l].valid && !(smax(dot(hit[l].n, ldir[l]), 0.0f) > 0.0f)) tr.cn.facing_away++;  // (statistics only)
            for (int l = 0; l < 8; l++) {
                if (!hit[l].valid || sh[l].occluded) continue;
                const Mat &m = sc.mats[hit[l].mat];
                float light_distance = mag(l2p[l]) + EPS;
                float cosi = dot(l2p[l], hit[l].n) / light_distance;
                bool pos = cosi > 0.0f;
                float att = 0.95f * (EPS + light_distance + light_distance * light_distance);
                float sig = (tanhf(att) + 1.0f) / 2.0f;
                float lfc = cosi * lt.intensity * sclamp(sig, 0.0f, 1.0f);
                V3 ccolor = pos ? (m.color * lt.color) : splat(0.0f);
                float cint = pos ? lfc : 0.0f;
                V3 lcs = ccolor / sh[l].filter;
                float diffuse = smax(dot(hit[l].n, ldir[l]), 0.0f);
                bool has_specular = m.shininess > 0.0f;
                float spec = 0.0f;
                if (has_specular) {
                    V3 r = reflected(ldir[l], hit[l].n);
                    spec = powf(smax(dot(normalized(r), rays[l].d), 0.0f), smax(m.shininess * 512.0f, 1.0f));
                }
                float lf = diffuse * cint * sh[l].opacity;
                float sf = cint * sh[l].opacity * spec;
                if (diffuse > 0.0f) {
                    light_color[l] = light_color[l] + m.color * lcs * lf;
                    if (has_specular) specular_color[l] = specular_color[l] + lt.color * sf;
                }
            }
        }
        for (int l = 0; l < 8; l++) {
            if (!hit[l].valid) continue;
            const Mat &m = sc.mats[hit[l].mat];
            V3 ambient = (m.color * splat(1.0f)) * 0.08f;
            float df = attenuation(hit[l].t);
            V3 direct = (ambient + light_color[l]) * df;
            V3 specular = specular_color[l] * df;
            out[l].valid = true;
            out[l].t = hit[l].t;
            out[l].color = m.transmissive() ? (splat(0.0f) + splat(0.0f) + specular) : (direct + splat(0.0f) + specular);
        }
    }

    // SceneLightSource::preprocess -> PointLight::to_point_light_cloud (light.rs:183-225,311-324)
    void pixel_lights(uint32_t px, uint32_t py, std::vector<PLight> &out) {
        uint32_t M = cfg.light_multiplicator ? cfg.light_multiplicator : 1;
        out.clear();
        bool packet_noaa = (cfg.flags & F_PACKET) && !(cfg.flags & F_AA);
        uint32_t kx = packet_noaa ? px / 8u : px;  // one cloud per 8-pixel packet (raytracer_renderer.rs:1274-1280)
        for (uint32_t li = 0; li < sc.lights.size(); li++) {
            const OLight &l = sc.lights[li];
            if (M == 1 || !cfg.light_clouds) {
                out.push_back({v3(l.position), v3(l.color), l.intensity});
                continue;
            }
            uint32_t ci = cloud_index(cfg.seed, kx, py, li, cfg.n_light_clouds);
            const float *cl = cfg.light_clouds + (size_t)ci * M * 3;
            float scale = 1.0f / (float)M;
            V3 w2s = {k.w2s_w, k.w2s_h, k.w2s_d};
            for (uint32_t c = 0; c < M; c++) {
                V3 rp = {cl[3 * c], cl[3 * c + 1], cl[3 * c + 2]};
                out.push_back({v3(l.position) + rp * w2s, v3(l.color), scale * l.intensity});
            }
        }
    }

    // get_antialiasing_simpling_directions + bundle_rays_for_simd_antialiased_raytracing
    // (raytracer_renderer.rs:876-916,1021-1138): per-sample scene-space offsets
    void aa_offsets(std::vector<float> &off) {
        bool rot = (cfg.flags & F_AA_ROT) != 0;
        bool packet = (cfg.flags & F_PACKET) != 0;
        V3 xr = {1, 0, 0}, yr = {0, 1, 0};
        if (rot) {
            float ang = atanf(0.5f);
            float s = sinf(ang), c = cosf(ang);
            V3 ux = {1, 0, 0}, uy = {0, 1, 0};
            xr = mul_add(ux, splat(c), uy * s);
            yr = mul_add(ux, splat(-s), uy * c);
        }
        V3 t = -yr, l = -xr, r = xr, b = yr;
        V3 dirs[8] = {normalized(t),     normalized(l),     normalized(b),     normalized(r),
                      normalized(t + l), normalized(t + r), normalized(b + l), normalized(b + r)};
        float scale = rot ? (sqrtf(5.0f) / 2.05f) : 0.85f;
        uint32_t total = packet ? cfg.aa_table_len : cfg.aa_samples_per_pixel;
        off.resize(2 * total);
        for (uint32_t i = 0; i < total; i++) {
            float rx = cfg.aa_table[2 * i] * k.w2s_w * scale;
            float ry = cfg.aa_table[2 * i + 1] * k.w2s_h * scale;
            const V3 &bias = dirs[packet ? (i % 8) : 0];
            off[2 * i] = rx * bias.x;
            off[2 * i + 1] = ry * bias.y;
        }
    }

    // get_pixel_color / render_pixel_colors (raytracer_renderer.rs:1140-1357)
    bool pixel(uint32_t px, uint32_t py, const std::vector<float> &off, V3 &color) {
        V3 coords = {(float)px * k.w2s_w, (float)py * k.w2s_h, 0.0f};
        V3 dir = coords - k.focus;
        std::vector<PLight> lights;
        pixel_lights(px, py, lights);
        if (!(cfg.flags & F_AA)) {
            Result r = single_raytrace(coords, dir, AIR_IOR, lights, -1);
            color = r.color;
            return r.valid;
        }
        // antialiased_raytrace (raytracer_renderer.rs:918-1016)
        uint32_t S = (uint32_t)off.size() / 2;
        float scale = 1.0f / (float)S;
        bool any = false;
        std::vector<V3> c(S);
        const bool packets = tr.mode == 2 && (cfg.flags & F_PACKET) && !refl_on && !refr_on && S % 8 == 0;
        for (uint32_t i = 0; i < S && packets; i += 8) {  // one Vec3x8 packet = 8 samples of this pixel
            Ray rays[8];
            for (int l = 0; l < 8; l++) rays[l] = make_ray({coords.x + off[2 * (i + l)], coords.y + off[2 * (i + l) + 1], coords.z}, dir);
            Result r8[8];
            single_raytrace8(rays, lights, r8);
            for (int l = 0; l < 8; l++) {
                any = any || r8[l].valid;
                c[i + l] = r8[l].valid ? r8[l].color * scale : splat(0.0f);
            }
        }
        for (uint32_t i = 0; i < S && !packets; i++) {
            V3 o = {coords.x + off[2 * i], coords.y + off[2 * i + 1], coords.z};
            Result r = single_raytrace(o, dir, AIR_IOR, lights, -1);
            any = any || r.valid;
            c[i] = r.valid ? r.color * scale : splat(0.0f);  // misses contribute nothing
        }
        if (!(cfg.flags & F_PACKET)) {
            // scalar path: rest samples folded, then + initial (order of the rayon reduce is
            // unspecified in the reference; fixed here to sample order)
            V3 res = splat(0.0f);
            for (uint32_t i = 1; i < S; i++) res = c[i] + res;
            color = res + c[0];
        } else {
            // packet path: 8 lanes; lane j = (packets 1.. summed) + packet 0, then horizontal sum
            V3 lane[8];
            for (int j = 0; j < 8; j++) {
                V3 res = splat(0.0f);
                for (uint32_t p = 1; p < S / 8; p++) res = c[8 * p + j] + res;
                lane[j] = res + c[j];
            }
            V3 q0 = lane[0] + lane[4], q1 = lane[1] + lane[5], q2 = lane[2] + lane[6], q3 = lane[3] + lane[7];
            color = (q0 + q2) + (q1 + q3);
        }
        return any;
    }
};

// OutputColorEncoder::to_output (src/output/window.rs:105-110): palette f32 -> u8 "magic add"
inline uint32_t encode_channel(float x) {
    float scaled = fminf(x * 255.0f, 255.0f);
    float f = scaled + 8388608.0f;
    uint32_t bits;
    memcpy(&bits, &f, 4);
    uint32_t r = bits < 0x4B000000u ? 0u : bits - 0x4B000000u;
    return r & 0xFFu;
}
inline uint32_t encode(V3 c) { return 0xFF000000u | (encode_channel(c.x) << 16) | (encode_channel(c.y) << 8) | encode_channel(c.z); }

void make_scene(const OSceneDesc *d, Scene &sc) {
    sc.mats.resize(d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const OMaterial &m = d->materials[i];
        sc.mats[i] = {v3(m.color), m.metallic, m.shininess, m.ior, m.opacity, m.has_opacity != 0, m.boost};
    }
    sc.prims.clear();
    for (uint32_t i = 0; i < d->n_spheres; i++) {
        const OSphere &s = d->spheres[i];
        Prim p{};
        p.kind = 0;
        p.a = v3(s.center);
        p.r_sq = s.radius * s.radius;  // SphereData::with_material (sphere.rs:40-48)
        p.mat = s.material;
        sc.prims.push_back(p);
    }
    for (uint32_t i = 0; i < d->n_triangles; i++) {
        const OTriangle &t = d->triangles[i];
        Prim p{};
        p.kind = 1;
        p.a = v3(t.v1);
        p.e1 = v3(t.v2) - v3(t.v1);  // TriangleData::with_material_and_normal (triangle.rs:81-102)
        p.e2 = v3(t.v3) - v3(t.v1);
        p.n = v3(t.normal);
        p.mat = t.material;
        sc.prims.push_back(p);
    }
    sc.lights.assign(d->lights, d->lights + d->n_lights);
}

}  // namespace

extern "C" {

int oracle_version(void) { return 1; }

// Renders either the whole frame (pixel_idx == NULL; row tiles honoured) or the listed pixels.
// out_argb/out_mask are width*height (zero-filled here; misses stay 0). out_rgb (optional) is
// width*height*3 floats. counters[6] = closest rays, shadow rays, triangle tests, sphere tests, box tests, shadow
// queries of hits that face away from the light.
int oracle_render(const OSceneDesc *desc, const OConfig *cfg, int mode, int threads, const uint32_t *pixel_idx,
                  uint32_t n_pixels, uint32_t *out_argb, uint8_t *out_mask, float *out_rgb, uint64_t *counters) {
    if (!desc || !cfg || !out_argb) return -1;
    Scene sc;
    make_scene(desc, sc);
    if (mode == 1) build_bvh(sc);
    const uint32_t W = cfg->width, H = cfg->height;
    std::vector<uint32_t> all;
    if (!pixel_idx) {
        for (uint32_t y = 0; y < H; y++) {
            if (cfg->tile_world > 1 && cfg->tile_rows && ((y / cfg->tile_rows) % cfg->tile_world) != cfg->tile_rank) continue;
            for (uint32_t x = 0; x < W; x++) all.push_back(y * W + x);
        }
        pixel_idx = all.data();
        n_pixels = (uint32_t)all.size();
    }
    memset(out_argb, 0, sizeof(uint32_t) * (size_t)W * H);
    if (out_mask) memset(out_mask, 0, (size_t)W * H);
    if (out_rgb) memset(out_rgb, 0, sizeof(float) * 3 * (size_t)W * H);
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint32_t> next{0};
    std::vector<Counters> cns(threads);
    auto work = [&](int tid) {
        Renderer r(sc, *cfg, mode);
        std::vector<float> off;
        if (cfg->flags & F_AA) r.aa_offsets(off);
        // chunk of the shared pixel queue: small enough that a short pixel list still keeps every thread busy
        // (>= 8 chunks per thread), at most 64 pixels
        const uint32_t CH = std::max<uint32_t>(1u, std::min<uint32_t>(64u, n_pixels / (8u * (uint32_t)threads)));
        for (;;) {
            uint32_t b = next.fetch_add(CH);
            if (b >= n_pixels) break;
            uint32_t e = std::min(n_pixels, b + CH);
            for (uint32_t i = b; i < e; i++) {
                uint32_t idx = pixel_idx[i];
                V3 c;
                bool valid = r.pixel(idx % W, idx / W, off, c);
                if (valid) {
                    out_argb[idx] = encode(c);
                    if (out_mask) out_mask[idx] = 1;
                    if (out_rgb) {
                        out_rgb[3 * (size_t)idx] = c.x;
                        out_rgb[3 * (size_t)idx + 1] = c.y;
                        out_rgb[3 * (size_t)idx + 2] = c.z;
                    }
                }
            }
        }
        cns[tid] = r.tr.cn;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    if (counters) {
        memset(counters, 0, 6 * sizeof(uint64_t));
        for (auto &c : cns) {
            counters[0] += c.closest;
            counters[1] += c.shadow;
            counters[2] += c.tri;
            counters[3] += c.sph;
            counters[4] += c.box;
            counters[5] += c.facing_away;
        }
    }
    return 0;
}

// Raw queries for BVH-vs-linear-scan tests. rays: n x 6 (origin, un-normalised direction).
int oracle_trace_closest(const OSceneDesc *desc, uint32_t flags, int mode, const float *rays, uint32_t n, float *out_t,
                         uint32_t *out_prim) {
    Scene sc;
    make_scene(desc, sc);
    if (mode == 1) build_bvh(sc);
    OConfig cfg{};
    cfg.flags = flags;
    Tracer tr(sc, cfg, mode);
    for (uint32_t i = 0; i < n; i++) {
        Ray r = make_ray(v3(rays + 6 * i), v3(rays + 6 * i + 3));
        Hit h = tr.cast_ray(r);
        out_t[i] = h.valid ? h.t : INF;
        out_prim[i] = h.valid ? h.prim : 0xFFFFFFFFu;
    }
    return 0;
}

// out: n x 5 = {occluded, opacity, filter r, g, b}
int oracle_trace_shadow(const OSceneDesc *desc, uint32_t flags, int mode, const float *rays, const float *max_dist,
                        uint32_t n, float *out) {
    Scene sc;
    make_scene(desc, sc);
    if (mode == 1) build_bvh(sc);
    OConfig cfg{};
    cfg.flags = flags;
    Tracer tr(sc, cfg, mode);
    for (uint32_t i = 0; i < n; i++) {
        Shadow s = tr.has_any_intersection(v3(rays + 6 * i), v3(rays + 6 * i + 3), max_dist[i]);
        out[5 * i] = s.occluded ? 1.0f : 0.0f;
        out[5 * i + 1] = s.opacity;
        out[5 * i + 2] = s.filter.x;
        out[5 * i + 3] = s.filter.y;
        out[5 * i + 4] = s.filter.z;
    }
    return 0;
}

uint32_t oracle_cloud_index(uint64_t seed, uint32_t x, uint32_t y, uint32_t light, uint32_t n) {
    return cloud_index(seed, x, y, light, n);
}

uint32_t oracle_encode(float r, float g, float b) { return encode({r, g, b}); }

}  // extern "C"
