// rt_api.cu -- the C ABI of librt_b200.so (include/rt_api.h): device scene lifetime, render
// calls, row-tile bookkeeping and the raw-query parity entry points.
//
// Replaces RaytracerRenderer::render (reference src/renderer/raytracer_renderer.rs:1360-1379)
// and the tiling driver Renderer::render_to_buffer[_chunked_inplace] (src/renderer/mod.rs:96-209)
// behind it.  There is no CPU path: without a CUDA device every compute entry point fails with
// RT_ERR_NO_DEVICE.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "rt_bvh.h"
#include "rt_bvh_device.h"
#include "rt_internal.h"
#include "rt_kernels.cuh"

using namespace rtd;

#define CU_CHECK(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            int _code = (_e == cudaErrorMemoryAllocation) ? RT_ERR_OUT_OF_MEMORY                           \
                        : (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? RT_ERR_NO_DEVICE \
                                                                                         : RT_ERR_CUDA;     \
            return rt_fail(_code, "%s failed: %s", #expr, cudaGetErrorString(_e));                         \
        }                                                                                                  \
    } while (0)

constexpr uint32_t MAX_PASSES = 64;  // light groups of the direct-lighting path

struct rt_scene {
    int device = 0;
    DScene d{};
    void *buf_nodes = nullptr, *buf_prims = nullptr, *buf_shade = nullptr, *buf_mats = nullptr, *buf_lights = nullptr;
    uint32_t n_nodes = 0, n_prims = 0;
    int sm_count = 0, ctas_per_sm = 0, ctas_per_sm_count = 0;
    // per-render scratch, grown on demand
    float4 *samples = nullptr;
    size_t samples_cap = 0;
    uint4 *primary_hits = nullptr;
    size_t primary_cap = 0;
    DirectState *direct_state = nullptr;  // running light sums between the light groups of direct_kernel (bounded: windows)
    size_t direct_cap = 0;
    // wavefront form of the bounce configs: the path queue of the chunk being rendered
    PathPoint *wave_points = nullptr;
    size_t wave_points_cap = 0;
    float4 *wave_contrib = nullptr;
    size_t wave_contrib_cap = 0;
    int32_t *wave_heads = nullptr;
    size_t wave_heads_cap = 0;
    unsigned int *wave_count = nullptr, *h_wave_count = nullptr;
    int paths_ctas_per_sm[2] = {0, 0};
    bool l2_policy_set = false;
    cudaStream_t l2_policy_stream = nullptr;
    uint64_t last_wave_points = 0;
    unsigned int *chain_done = nullptr;   // per batch of a window: tag of the window's last unit (slot reuse across windows)
    size_t chain_done_cap = 0;
    unsigned int *pass_counters = nullptr;  // work counter of direct_kernel
    HeavyQueue *heavy_queue = nullptr;      // list of deferred (very expensive) units of direct_kernel
    RenderParams *d_params = nullptr;       // global-memory copy of the kernel parameters for direct_kernel's out-of-line paths
    RenderParams *h_params = nullptr;       // pinned staging buffers of it (ring: a copy must outlive its async upload)
    uint32_t h_params_next = 0;
    RenderParams d_params_host{};           // what d_params currently holds
    bool d_params_valid = false;
    unsigned int *d_watchdog = nullptr;     // sticky device word: != 0 when a bounded wait loop of direct_kernel gave up
    unsigned int *h_watchdog = nullptr;     // pinned: its value after the last fetched frame
    bool frame_has_watchdog = false;        // the frame being enqueued launches direct_kernel
    bool watchdog_pending = false;          // a fetch has been enqueued and not been looked at yet
    cudaEvent_t ev_watchdog = nullptr;      // recorded after that fetch
    cudaStream_t last_stream = nullptr;     // stream of the previous call + event recorded at its end: a call on ANOTHER
    cudaEvent_t ev_last = nullptr;          // stream first waits for it (the scratch buffers are shared)
    bool has_last = false;
    float4 *heavy_scratch = nullptr;        // per deferred unit: M copies x 32 samples x 2 float4
    size_t heavy_scratch_cap = 0;
    uint32_t frame_serial = 0;              // base of the chain tags of the current frame
    int ctas_per_sm_direct = 0, ctas_per_sm_direct_count = 0;
    int direct_ctas_per_sm[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // per direct_kernel variant, filled on first use
    float *clouds_raw = nullptr;  // K x M x 3 cloud points, as handed over
    float4 *clouds = nullptr;     // the same points scaled to scene units (scale_clouds_kernel)
    size_t clouds_raw_cap = 0, clouds_cap = 0;
    uint32_t clouds_epoch = 0, clouds_k = 0, clouds_m = 0, clouds_w = 0, clouds_h = 0;  // what the device copy currently holds
    float cloud_lo[3] = {0, 0, 0}, cloud_hi[3] = {0, 0, 0};  // bounds of the points in it
    const float *clouds_src = nullptr;
    unsigned int *work_counter = nullptr;
    unsigned long long *counters = nullptr;
    uint32_t *slab_argb = nullptr;
    size_t slab_cap = 0;
    uint32_t *pinned_argb = nullptr;
    size_t pinned_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // ring of event pairs around the dominant kernel (direct_kernel / render_kernel) of the last renders
    static constexpr uint32_t KERNEL_EVENT_RING = 64;
    cudaEvent_t kev[KERNEL_EVENT_RING][2] = {};
    uint64_t kev_count = 0;
    // rt_render_progressive: two-deep band pipeline
    uint32_t *band_slab[2] = {nullptr, nullptr};
    size_t band_slab_cap[2] = {0, 0};
    uint32_t *band_host[2] = {nullptr, nullptr};
    size_t band_host_cap[2] = {0, 0};
    cudaEvent_t band_rendered[2] = {nullptr, nullptr}, band_copied[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    uint32_t last_progressive_overlapped = 0;
    char last_kernel[64] = "";
    rt_scene_info info{};
};

static cudaEvent_t kernel_event(rt_scene *s, int which) {
    cudaEvent_t &e = s->kev[s->kev_count % rt_scene::KERNEL_EVENT_RING][which];
    if (!e && cudaEventCreate(&e) != cudaSuccess) e = nullptr;
    return e;
}

static uint32_t local_rows_of(const rt_config *c) {
    if (c->tile_world <= 1 || c->tile_rows == 0) return c->height;
    uint32_t rows = 0;
    uint32_t n_tiles = (c->height + c->tile_rows - 1) / c->tile_rows;
    for (uint32_t t = c->tile_rank; t < n_tiles; t += c->tile_world) {
        uint32_t r0 = t * c->tile_rows;
        rows += (r0 + c->tile_rows <= c->height) ? c->tile_rows : (c->height - r0);
    }
    return rows;
}

static uint32_t image_row_of(const rt_config *c, uint32_t lr) {
    if (c->tile_world <= 1 || c->tile_rows == 0) return lr;
    uint32_t t = lr / c->tile_rows, r = lr - t * c->tile_rows;
    return (t * c->tile_world + c->tile_rank) * c->tile_rows + r;
}

static int validate_config(const rt_config *c) {
    if (!c) return rt_fail(RT_ERR_INVALID_ARGUMENT, "null rt_config");
    if (c->width == 0 || c->height == 0 || c->width > 65536 || c->height > 65536)
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "invalid image size %ux%u", c->width, c->height);
    if (c->tile_world > 1 && (c->tile_rows == 0 || c->tile_rank >= c->tile_world))
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "invalid row tiling (rows %u, rank %u, world %u)", c->tile_rows, c->tile_rank, c->tile_world);
    if (c->flags & RT_FLAG_ANTI_ALIASING) {
        uint32_t S = (c->flags & RT_FLAG_PACKET_SEMANTICS) ? c->aa_table_len : c->aa_samples_per_pixel;
        if (!c->aa_table || S == 0 || S > (uint32_t)MAX_AA || c->aa_table_len < S)
            return rt_fail(RT_ERR_INVALID_ARGUMENT, "anti-aliasing needs aa_table with 1..%d samples (got %u, table %u)", MAX_AA, S, c->aa_table_len);
        if ((c->flags & RT_FLAG_PACKET_SEMANTICS) && (S % 8))
            return rt_fail(RT_ERR_INVALID_ARGUMENT, "packet semantics need a multiple of 8 samples (got %u)", S);
    }
    {
        // work items, sample indices and unit numbers are 32-bit on the device: reject frames that would wrap
        const uint32_t S = (c->flags & RT_FLAG_ANTI_ALIASING) ? ((c->flags & RT_FLAG_PACKET_SEMANTICS) ? c->aa_table_len : c->aa_samples_per_pixel) : 1u;
        const uint64_t tiles_x = (c->width + TILE_W - 1) / TILE_W, tiles_y = ((uint64_t)local_rows_of(c) + TILE_H - 1) / TILE_H;
        const uint64_t items = tiles_x * tiles_y * (TILE_W * TILE_H) * std::max<uint64_t>(S, 1);
        if (items >= (1ull << 31))
            return rt_fail(RT_ERR_INVALID_ARGUMENT, "frame too large for one call: %llu primary samples (limit 2^31); render it in row tiles",
                           (unsigned long long)items);
    }
    if (c->light_multiplicator > 1 && (!c->light_clouds || c->n_light_clouds == 0))
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "light_multiplicator %u needs light_clouds", c->light_multiplicator);
    if (c->reflection_max_depth > 0xFFFF || c->refraction_max_depth > 0xFFFF || c->reflection_max_depth >= (uint32_t)PATH_STACK - 1 ||
        c->refraction_max_depth >= (uint32_t)PATH_STACK - 1)
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "bounce depth %u/%u exceeds the path stack (%d)", c->reflection_max_depth, c->refraction_max_depth, PATH_STACK - 2);
    return RT_OK;
}

// crate constants (reference src/lib.rs:73-92) and AA offsets
// (raytracer_renderer.rs:876-916,1021-1138), evaluated in f32 on the host
static void fill_params(const rt_scene *s, const rt_config *c, RenderParams &p) {
    memset(&p, 0, sizeof(p));
    p.sc = s->d;
    const uint32_t W = c->width, H = c->height;
    p.width = W;
    p.height = H;
    p.flags = c->flags;
    const float aspect = (float)H / (float)W;
    const float sw = 1.0f, sh = sw * aspect, sd = (sw + sh) / 2.0f;
    const float avg = (sw + sh + sd) / 3.0f;
    p.w2s_w = sw / (float)W;
    p.w2s_h = sh / (float)H;
    p.w2s_d = sd / (float)((W + H) / 2);
    p.focus_x = sw / 2.0f;
    p.focus_y = sh / 2.0f;
    p.focus_z = -1.9f * sd;
    p.eps_d = 1.1920929e-7f * (100.0f * avg);
    const bool aa = (c->flags & RT_FLAG_ANTI_ALIASING) != 0, packet = (c->flags & RT_FLAG_PACKET_SEMANTICS) != 0;
    p.S = aa ? (packet ? c->aa_table_len : c->aa_samples_per_pixel) : 1;
    p.sample_scale = 1.0f / (float)p.S;
    if (aa) {
        const bool rot = (c->flags & RT_FLAG_AA_ROTATION_SCALE) != 0;
        float xr[3] = {1, 0, 0}, yr[3] = {0, 1, 0};
        if (rot) {
            float ang = atanf(0.5f), sn = sinf(ang), cs = cosf(ang);
            // x.mul_add(cos, y * sin), x.mul_add(-sin, y * cos) on the unit vectors
            xr[0] = fmaf(1.0f, cs, 0.0f * sn); xr[1] = fmaf(0.0f, cs, 1.0f * sn); xr[2] = fmaf(0.0f, cs, 0.0f * sn);
            yr[0] = fmaf(1.0f, -sn, 0.0f * cs); yr[1] = fmaf(0.0f, -sn, 1.0f * cs); yr[2] = fmaf(0.0f, -sn, 0.0f * cs);
        }
        float dirs[8][3];
        for (int k = 0; k < 3; k++) {
            float t = -yr[k], l = -xr[k], r = xr[k], b = yr[k];
            dirs[0][k] = t; dirs[1][k] = l; dirs[2][k] = b; dirs[3][k] = r;
            dirs[4][k] = t + l; dirs[5][k] = t + r; dirs[6][k] = b + l; dirs[7][k] = b + r;
        }
        for (auto &v : dirs) {
            float m = sqrtf(fmaf(v[0], v[0], fmaf(v[1], v[1], v[2] * v[2])));
            float inv = 1.0f / m;
            v[0] *= inv; v[1] *= inv; v[2] *= inv;
        }
        const float scale = rot ? (sqrtf(5.0f) / 2.05f) : 0.85f;
        for (uint32_t i = 0; i < p.S; i++) {
            float rx = c->aa_table[2 * i] * p.w2s_w * scale;
            float ry = c->aa_table[2 * i + 1] * p.w2s_h * scale;
            const float *bias = dirs[packet ? (i % 8) : 0];
            p.aa_off[2 * i] = rx * bias[0];
            p.aa_off[2 * i + 1] = ry * bias[1];
        }
    }
    p.M = c->light_multiplicator ? c->light_multiplicator : 1;
    p.LM = p.M * s->d.n_lights;
    p.inv_M = 1.0f / (float)p.M;
    p.K = c->n_light_clouds;
    p.seed = c->seed;
    p.refl_depth = c->reflection_max_depth;
    p.refr_depth = c->refraction_max_depth;
    p.tile_rows = c->tile_rows;
    p.tile_rank = c->tile_rank;
    p.tile_world = c->tile_world;
    p.local_rows = local_rows_of(c);
    p.tiles_x = (W + TILE_W - 1) / TILE_W;
    uint32_t tiles_y = (p.local_rows + TILE_H - 1) / TILE_H;
    p.n_items = p.tiles_x * tiles_y * (TILE_W * TILE_H) * p.S;
}

template <typename T>
static int grow(T *&ptr, size_t &cap, size_t need_bytes) {
    if (need_bytes <= cap) return RT_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    CU_CHECK(cudaMalloc((void **)&ptr, need_bytes));
    cap = need_bytes;
    return RT_OK;
}

static int upload(void **dst, const std::vector<float> &src) {
    size_t bytes = std::max<size_t>(src.size() * sizeof(float), 16);
    CU_CHECK(cudaMalloc(dst, bytes));
    if (!src.empty()) CU_CHECK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice));
    return RT_OK;
}

static int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return rt_fail(RT_ERR_NO_DEVICE, "no CUDA device available (%s); librt_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return rt_fail(RT_ERR_INVALID_ARGUMENT, "device %d out of range (0..%d)", device, n - 1);
    CU_CHECK(cudaSetDevice(device));
    return RT_OK;
}

// global-memory copy of the kernel parameters for the out-of-line device paths (uploaded only when it changed)
static int upload_params(rt_scene *s, const RenderParams &p, cudaStream_t stream) {
    constexpr uint32_t RING = 16;  // pinned staging ring: a copy must outlive its async upload
    if (!s->d_params) CU_CHECK(cudaMalloc((void **)&s->d_params, sizeof(RenderParams)));
    if (!s->h_params) CU_CHECK(cudaMallocHost((void **)&s->h_params, RING * sizeof(RenderParams)));
    if (s->d_params_valid && memcmp(&s->d_params_host, &p, sizeof(RenderParams)) == 0) return RT_OK;
    RenderParams *stage = s->h_params + (s->h_params_next++ % RING);
    memcpy(stage, &p, sizeof(RenderParams));
    CU_CHECK(cudaMemcpyAsync(s->d_params, stage, sizeof(RenderParams), cudaMemcpyHostToDevice, stream));
    memcpy(&s->d_params_host, &p, sizeof(RenderParams));
    s->d_params_valid = true;
    return RT_OK;
}

// RT_L2_PERSIST=1 (A/B switch, default off): asks the L2 to keep the BVH node array resident
// (cudaAccessPolicyWindow on the launching stream).  Only matters for scenes whose nodes + primitives exceed the
// 126 MB L2 (BASELINE configs[4]); measured in profiles/README.md.
static void apply_l2_policy(rt_scene *s, cudaStream_t stream) {
    static const int want = getenv("RT_L2_PERSIST") ? atoi(getenv("RT_L2_PERSIST")) : 0;
    if (!want || (s->l2_policy_set && s->l2_policy_stream == stream)) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, s->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0) return;
    const size_t bytes = (size_t)s->n_nodes * 128;
    const size_t window = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
    const size_t carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, window);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = s->buf_nodes;
    attr.accessPolicyWindow.num_bytes = window;
    attr.accessPolicyWindow.hitRatio = window ? std::min(1.0f, (float)carve / (float)window) : 0.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    s->l2_policy_set = true;
    s->l2_policy_stream = stream;
}

// Direct lighting of `p.n_items` work items -- primary samples (points == false: the frames without bounces) or the
// shading points of a path queue (points == true: the wavefront form of the bounce configs) -- with direct_kernel.
static int enqueue_direct_lighting(rt_scene *s, const rt_config *cfg, const RenderParams &p, uint32_t warps_needed, bool count, bool points,
                                   cudaStream_t stream, uint64_t *launches, bool record_events) {
    int rc = RT_OK;
    {
        // direct lighting only: one pass per light group over all samples
        uint32_t n_groups = p.M > 1 ? s->d.n_lights : (p.LM ? 1u : 0u);  // (soft shadows: per light; regrouped below)
        uint32_t passes = std::max(n_groups, 1u);  // a scene without lights still writes its samples
        if (passes > MAX_PASSES) return rt_fail(RT_ERR_INVALID_ARGUMENT, "too many lights for the soft-shadow path (%u > %u)", passes, MAX_PASSES);
        if ((uint64_t)warps_needed * passes >= (1ull << 32))
            return rt_fail(RT_ERR_INVALID_ARGUMENT, "frame too large for one call: %llu work units (limit 2^32); render it in row tiles",
                           (unsigned long long)warps_needed * passes);
        RenderParams pd = p;
        if (n_groups == 0) {  // no lights: ambient term only
            pd.LM = 0;
            pd.M = 1;
        }
        // shadow beams need M copies per light and usable bounds of the cloud points
        const bool beams = pd.M > 1 && !(cfg->flags & RT_FLAG_PER_RAY_SHADOWS) && std::isfinite(pd.cloud_lo[0] + pd.cloud_lo[1] + pd.cloud_lo[2]) &&
                           std::isfinite(pd.cloud_hi[0] + pd.cloud_hi[1] + pd.cloud_hi[2]);
        // few units per warp (a rank's share of a multi-GPU frame): the tail of the launch matters more than the
        // last few per cent of steady-state throughput -> the variant that defers very expensive units
        const uint32_t resident_warps = (uint32_t)s->sm_count * (uint32_t)std::max(s->ctas_per_sm_direct, 1) * CTA_WARPS;
        const uint32_t units_per_warp = (uint32_t)(((uint64_t)warps_needed * passes) / resident_warps);
        uint32_t defer_below = 50;  // (measured with all lights in a unit: the plain variant wins a rank's quarter of the C3 frame -- 74 -- the deferring one its eighth -- 37)
        if (const char *e = getenv("RT_DEFER_BELOW")) defer_below = (uint32_t)std::max(0, atoi(e));  // tuning experiments
        const bool defer = beams && units_per_warp < defer_below;
        // Soft shadows: how many lights a unit takes in a row.  A unit lights its 32 samples with up to 8 lights in a row
        // -- all lights of every shipped scene --, the running sums staying in registers: no chained records, no waits,
        // a fifth of the unit bookkeeping (measured: profiles/README.md).  A scene with more lights is lit in groups of
        // 8, chained through the records; RT_LIGHTS_PER_UNIT (A/B runs, tests) selects any other group size.  The
        // deferring variant keeps its records, tags and list entries per LIGHT whatever the unit size (a unit that defers
        // at one of its lights goes on as the per-light units of its batch).
        const uint32_t lights_pl = std::max(n_groups, 1u);  // per-light count (soft shadows), 1 otherwise
        uint32_t lights_per_unit = 1;
        if (pd.M > 1 && n_groups > 0) {
            lights_per_unit = std::min(n_groups, 8u);
            if (const char *e = getenv("RT_LIGHTS_PER_UNIT")) lights_per_unit = (uint32_t)std::max(1, std::min<int>(atoi(e), (int)n_groups));
            n_groups = (n_groups + lights_per_unit - 1) / lights_per_unit;
            passes = n_groups;
        }
        pd.lights_per_unit = lights_per_unit;
        // Chained records between the light groups of a batch (direct_kernel): windows of `chain_window` batches with
        // two generations each; the DEFER variant keeps one window with every generation.
        // The window must be several times the number of resident warps: a unit's predecessor was handed out one window
        // earlier, and only if (nearly) all of it has been worked off by then does the unit find its record waiting
        // (measured on B200, profiles/README.md: 16384 batches = 4 x the resident warps costs 8 % of the frame in waits,
        // 65536 nothing).
        uint32_t chain_window = 65536;
        if (const char *e = getenv("RT_CHAIN_WINDOW")) chain_window = (uint32_t)std::max(1, atoi(e));  // A/B runs, tests
        chain_window = std::max(chain_window, 2u * resident_warps);  // (below: one working warp per batch column, the rest waiting)
        chain_window = defer ? warps_needed : std::min(chain_window, warps_needed);
        const uint32_t n_windows = (warps_needed + chain_window - 1) / chain_window;
        const uint32_t gens = defer ? std::max(lights_pl, 2u) - 1 : passes > 1 ? std::min(2u, passes - 1) : 1u;
        {
            const size_t before = s->direct_cap, before_done = s->chain_done_cap;
            const size_t n_records = (defer ? lights_pl > 1 : passes > 1) ? (size_t)chain_window * 32 * gens : 1;  // (one group: nothing is chained)
            rc = grow(s->direct_state, s->direct_cap, n_records * sizeof(DirectState));
            if (rc) return rc;
            if ((rc = grow(s->chain_done, s->chain_done_cap, (size_t)chain_window * sizeof(unsigned int)))) return rc;
            if (s->direct_cap != before) CU_CHECK(cudaMemsetAsync(s->direct_state, 0, s->direct_cap, stream));  // no stale tags
            if (s->chain_done_cap != before_done) CU_CHECK(cudaMemsetAsync(s->chain_done, 0, s->chain_done_cap, stream));
        }
        // tags of this frame's chained records, one per (window, group): never equal to a tag an earlier frame left behind
        const uint64_t tags = (defer ? (uint64_t)lights_pl : (uint64_t)n_windows * passes) + 1;
        if (tags >= 0x7FFF0000ull) return rt_fail(RT_ERR_INVALID_ARGUMENT, "frame too large for one call (%llu chain tags)", (unsigned long long)tags);
        if (s->frame_serial == 0 || (uint64_t)s->frame_serial + tags > 0x7FFF0000ull) {  // (the top bit of a chain tag is CHAIN_DEFERRED)
            CU_CHECK(cudaMemsetAsync(s->direct_state, 0, s->direct_cap, stream));
            CU_CHECK(cudaMemsetAsync(s->chain_done, 0, s->chain_done_cap, stream));
            s->frame_serial = 1;
        }
        const uint32_t frame_tag = s->frame_serial;
        s->frame_serial += (uint32_t)tags;
        pd.chain_window = chain_window;
        pd.chain_done = s->chain_done;
        using kernel_t = void (*)(const RenderParams, uint32_t, uint32_t, unsigned int *, DirectState *, HeavyQueue *, float4 *, const RenderParams *,
                                  unsigned int *);
        kernel_t kernel;
        if (points) {
            if (count)
                kernel = beams ? (defer ? direct_kernel<true, true, true, true> : direct_kernel<true, true, false, true>) : direct_kernel<true, false, false, true>;
            else
                kernel = beams ? (defer ? direct_kernel<false, true, true, true> : direct_kernel<false, true, false, true>) : direct_kernel<false, false, false, true>;
        } else if (count)
            kernel = beams ? (defer ? direct_kernel<true, true, true> : direct_kernel<true, true, false>) : direct_kernel<true, false, false>;
        else
            kernel = beams ? (defer ? direct_kernel<false, true, true> : direct_kernel<false, true, false>) : direct_kernel<false, false, false>;
        // persistent grid: SMs x resident CTAs of the selected variant (they differ in register budget)
        const int variant = (points ? 6 : 0) + (count ? 3 : 0) + (beams ? (defer ? 2 : 1) : 0);
        if (s->direct_ctas_per_sm[variant] == 0) {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, CTA_THREADS, 0) != cudaSuccess || n < 1) n = 1;
            s->direct_ctas_per_sm[variant] = n;
        }
        int dgrid = s->sm_count * s->direct_ctas_per_sm[variant];
        dgrid = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)dgrid, (warps_needed * passes + CTA_WARPS - 1) / CTA_WARPS));
        HeavyQueue *hq = nullptr;
        if (defer) {
            if (!s->heavy_queue) CU_CHECK(cudaMalloc((void **)&s->heavy_queue, sizeof(HeavyQueue)));
            pd.heavy_cap = HEAVY_CAP;
            if (const char *e = getenv("RT_HEAVY_CAP")) pd.heavy_cap = (uint32_t)std::min<long>(HEAVY_CAP, std::max<long>(0, atol(e)));  // tests: the list-full path
            if (const char *e = getenv("RT_DEFER_ALL")) pd.defer_all = atoi(e) != 0;  // tests: every unit asks for deferral
            pd.defer_cand = 32;  // (B200, a rank's eighth of the C3 frame, max over the 8 ranks: 48 -> 1.56 ms, 32 -> 1.38, 24 -> 1.42, 16 -> 1.50)
            if (const char *e = getenv("RT_DEFER_CAND")) pd.defer_cand = (uint32_t)std::max(0, atoi(e));  // A/B runs
            rc = grow(s->heavy_scratch, s->heavy_scratch_cap, (size_t)HEAVY_CAP * pd.M * 32 * 2 * sizeof(float4));
            if (rc) return rc;
            CU_CHECK(cudaMemsetAsync(s->heavy_queue, 0, sizeof(HeavyQueue), stream));  // list, tickets, watchdog and the unit counter
            hq = s->heavy_queue;
            if ((rc = upload_params(s, pd, stream))) return rc;
        }
        unsigned int *unit_counter = hq ? &hq->unit_counter : s->pass_counters;
        if (!hq) CU_CHECK(cudaMemsetAsync(s->pass_counters, 0, sizeof(unsigned int), stream));
        if (record_events)
            if (cudaEvent_t e = kernel_event(s, 0)) cudaEventRecord(e, stream);
        snprintf(s->last_kernel, sizeof(s->last_kernel), points ? "paths_kernel + direct_kernel<%s,%s,%s,points> + gather_paths_kernel" : "direct_kernel<%s,%s,%s>",
                 count ? "true" : "false", beams ? "true" : "false", defer ? "true" : "false");
        s->frame_has_watchdog = true;  // every variant has bounded wait loops (chained records)
        kernel<<<dgrid, CTA_THREADS, 0, stream>>>(pd, passes, frame_tag, unit_counter, s->direct_state, hq, s->heavy_scratch, s->d_params,
                                                  s->d_watchdog);
        CU_CHECK(cudaGetLastError());
        if (record_events) {
            if (cudaEvent_t e = kernel_event(s, 1)) cudaEventRecord(e, stream);
            s->kev_count++;
        }
        if (launches) *launches += 1;
    }
    return RT_OK;
}

// Bounce configs, wavefront form (rt_kernels.cuh): per chunk of primary samples, paths_kernel appends the shading
// points of the reflection / refraction trees to a queue, direct_kernel<.., POINTS> lights the queue, and
// gather_paths_kernel sums every sample's contributions in path order.  The host reads the number of points after
// paths_kernel (it sizes the lighting launch), so this path synchronises `stream` once per chunk; a chunk whose
// points do not fit the queue is rendered again in two halves.
static int enqueue_wavefront(rt_scene *s, const rt_config *cfg, RenderParams &p, bool count, cudaStream_t stream, uint64_t *launches) {
    int rc = RT_OK;
    uint32_t chunk_items = 1u << 22;
    if (const char *e = getenv("RT_WAVE_CHUNK")) chunk_items = (uint32_t)std::max(32, atoi(e));  // tests: small chunks
    chunk_items = std::min((chunk_items + 31u) & ~31u, (p.n_items + 31u) & ~31u);
    uint64_t capacity = std::min<uint64_t>(1ull << 25, std::max<uint64_t>(1ull << 16, 12ull * chunk_items));
    if (const char *e = getenv("RT_WAVE_POINTS")) capacity = (uint64_t)std::max(64, atoi(e));  // tests: the overflow path
    if ((rc = grow(s->wave_points, s->wave_points_cap, capacity * sizeof(PathPoint)))) return rc;
    if ((rc = grow(s->wave_contrib, s->wave_contrib_cap, capacity * sizeof(float4)))) return rc;
    if ((rc = grow(s->wave_heads, s->wave_heads_cap, (size_t)chunk_items * sizeof(int32_t)))) return rc;
    if (!s->wave_count) CU_CHECK(cudaMalloc((void **)&s->wave_count, sizeof(unsigned int)));
    if (!s->h_wave_count) CU_CHECK(cudaMallocHost((void **)&s->h_wave_count, sizeof(unsigned int)));
    if (s->paths_ctas_per_sm[count ? 1 : 0] == 0) {
        int n = 0;
        cudaError_t e = count ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, paths_kernel<true>, CTA_THREADS, 0)
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, paths_kernel<false>, CTA_THREADS, 0);
        s->paths_ctas_per_sm[count ? 1 : 0] = (e == cudaSuccess && n >= 1) ? n : 1;
    }
    p.pq.points = s->wave_points;
    p.pq.contrib = s->wave_contrib;
    p.pq.sample_head = s->wave_heads;
    p.pq.n_points = s->wave_count;
    p.pq.capacity = (uint32_t)capacity;
    if (cudaEvent_t e = kernel_event(s, 0)) cudaEventRecord(e, stream);
    uint64_t total_points = 0;
    for (uint32_t first = 0; first < p.n_items;) {
        const uint32_t n = std::min(chunk_items, p.n_items - first);
        p.pq.first_item = first;
        p.pq.n_chunk_items = n;
        CU_CHECK(cudaMemsetAsync(s->wave_count, 0, sizeof(unsigned int), stream));
        CU_CHECK(cudaMemsetAsync(s->work_counter, 0, sizeof(unsigned int), stream));
        int grid = s->sm_count * s->paths_ctas_per_sm[count ? 1 : 0];
        grid = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)grid, ((n + 31) / 32 + CTA_WARPS - 1) / CTA_WARPS));
        p.tail_div = 2u * (uint32_t)grid * CTA_WARPS;
        if (count)
            paths_kernel<true><<<grid, CTA_THREADS, 0, stream>>>(p);
        else
            paths_kernel<false><<<grid, CTA_THREADS, 0, stream>>>(p);
        CU_CHECK(cudaGetLastError());
        if (launches) *launches += 1;
        CU_CHECK(cudaMemcpyAsync(s->h_wave_count, s->wave_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
        CU_CHECK(cudaStreamSynchronize(stream));
        const unsigned int n_points = *s->h_wave_count;
        if (n_points > capacity) {  // the queue is too small for this chunk: two halves
            if (n <= 32) return rt_fail(RT_ERR_OUT_OF_MEMORY, "wavefront: %u shading points of 32 samples do not fit the path queue (%llu)", n_points,
                                        (unsigned long long)capacity);
            chunk_items = std::max(32u, ((n / 2) + 31u) & ~31u);
            continue;
        }
        if (n_points) {
            RenderParams pl = p;
            pl.n_items = n_points;
            if ((rc = enqueue_direct_lighting(s, cfg, pl, (n_points + 31u) / 32u, count, true, stream, launches, false))) return rc;
        }
        gather_paths_kernel<<<(n + 255) / 256, 256, 0, stream>>>(p);
        CU_CHECK(cudaGetLastError());
        if (launches) *launches += 1;
        total_points += n_points;
        first += n;
    }
    if (total_points == 0) snprintf(s->last_kernel, sizeof(s->last_kernel), "paths_kernel + gather_paths_kernel");
    if (cudaEvent_t e = kernel_event(s, 1)) cudaEventRecord(e, stream);
    s->kev_count++;
    s->last_wave_points = total_points;
    return RT_OK;
}

// enqueue render + resolve on `stream`; slab outputs are device pointers
static int enqueue_render(rt_scene *s, const rt_config *cfg, RenderParams &p, uint32_t *d_argb, uint8_t *d_mask, cudaStream_t stream,
                          uint64_t *launches, bool to_frame = false) {
    fill_params(s, cfg, p);
    s->frame_has_watchdog = false;
    apply_l2_policy(s, stream);
    // the scene's scratch buffers and cached tables are shared by all calls: a call on another stream than the
    // previous one is ordered after it
    if (s->has_last && s->last_stream != stream) CU_CHECK(cudaStreamWaitEvent(stream, s->ev_last, 0));
    size_t n_samples = (size_t)p.local_rows * p.width * p.S;
    int rc = grow(s->samples, s->samples_cap, std::max<size_t>(n_samples, 1) * sizeof(float4));
    if (rc) return rc;
    p.samples = s->samples;
    rc = grow(s->primary_hits, s->primary_cap, std::max<size_t>(p.n_items, 1) * sizeof(uint4));
    if (rc) return rc;
    p.primary_hits = s->primary_hits;
    p.work_counter = s->work_counter;
    p.counters = s->counters;
    if (p.local_rows == 0) return RT_OK;
    const bool count = (cfg->flags & RT_FLAG_COUNT_WORK) != 0;
    if (count) CU_CHECK(cudaMemsetAsync(s->counters, 0, 16 * sizeof(unsigned long long), stream));
    int grid = s->sm_count * (count ? s->ctas_per_sm_count : s->ctas_per_sm);
    uint32_t warps_needed = (p.n_items + 31) / 32;
    grid = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)grid, (warps_needed + CTA_WARPS - 1) / CTA_WARPS));
    p.tail_div = 2u * (uint32_t)grid * CTA_WARPS;
    if (const char *e = getenv("RT_TAIL_DIV")) p.tail_div = (uint32_t)std::max(1, atoi(e));  // tuning experiments
    const unsigned pgrid = (p.n_items + 127u) / 128u;
    const bool bounces = (cfg->flags & (RT_FLAG_REFLECTIONS | RT_FLAG_REFRACTIONS)) != 0;
    if (count)
        primary_kernel<true><<<pgrid, 128, 0, stream>>>(p);
    else
        primary_kernel<false><<<pgrid, 128, 0, stream>>>(p);
    CU_CHECK(cudaGetLastError());
    if (launches) *launches += 1;
    // the soft-shadow cloud table is needed from here on: its host-side work (bounds scan, staging of the
    // host-to-device copy) overlaps the primary pre-pass kernel that is already running
    if (p.LM > 0 && p.M > 1) {
        const size_t n_pts = (size_t)cfg->n_light_clouds * p.M;
        const bool cached = cfg->tables_epoch != 0 && cfg->tables_epoch == s->clouds_epoch && cfg->light_clouds == s->clouds_src &&
                            cfg->n_light_clouds == s->clouds_k && p.M == s->clouds_m && cfg->width == s->clouds_w && cfg->height == s->clouds_h &&
                            s->clouds;
        if (!cached) {
            rc = grow(s->clouds_raw, s->clouds_raw_cap, n_pts * 3 * sizeof(float));
            if (rc) return rc;
            if ((rc = grow(s->clouds, s->clouds_cap, n_pts * sizeof(float4)))) return rc;
            CU_CHECK(cudaMemcpyAsync(s->clouds_raw, cfg->light_clouds, n_pts * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
            scale_clouds_kernel<<<(unsigned)((n_pts + 255) / 256), 256, 0, stream>>>(s->clouds_raw, s->clouds, (uint32_t)n_pts, p.w2s_w, p.w2s_h, p.w2s_d);
            CU_CHECK(cudaGetLastError());
            if (launches) *launches += 1;
            {  // bounds of the points (one pass; a non-finite point leaves no usable bound: beams are switched off)
                float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
                bool finite = true;
                const float *pt = cfg->light_clouds;
                for (size_t i = 0; i < n_pts; i++, pt += 3) {
                    lo[0] = std::min(lo[0], pt[0]); hi[0] = std::max(hi[0], pt[0]);
                    lo[1] = std::min(lo[1], pt[1]); hi[1] = std::max(hi[1], pt[1]);
                    lo[2] = std::min(lo[2], pt[2]); hi[2] = std::max(hi[2], pt[2]);
                    finite &= (pt[0] - pt[0] == 0.0f) & (pt[1] - pt[1] == 0.0f) & (pt[2] - pt[2] == 0.0f);
                }
                for (int k = 0; k < 3; k++) {
                    s->cloud_lo[k] = finite ? lo[k] : -INFINITY;
                    s->cloud_hi[k] = finite ? hi[k] : INFINITY;
                }
            }
            s->clouds_epoch = cfg->tables_epoch;
            s->clouds_src = cfg->light_clouds;
            s->clouds_k = cfg->n_light_clouds;
            s->clouds_m = p.M;
            s->clouds_w = cfg->width;
            s->clouds_h = cfg->height;
        }
        p.clouds = s->clouds;
        for (int k = 0; k < 3; k++) {
            p.cloud_lo[k] = s->cloud_lo[k];
            p.cloud_hi[k] = s->cloud_hi[k];
        }
    }
    bool wavefront = bounces;
    if (const char *e = getenv("RT_WAVEFRONT")) wavefront = wavefront && atoi(e) != 0;  // 0: the round-1 megakernel (A/B runs, tests)
    if (wavefront) {
        if ((rc = enqueue_wavefront(s, cfg, p, count, stream, launches))) return rc;
    } else if (bounces) {
        // per-sample reflection / refraction trees: persistent megakernel
        CU_CHECK(cudaMemsetAsync(s->work_counter, 0, sizeof(unsigned int), stream));
        // shadow beams (lanes lighting their own hit) need M copies per light and usable bounds of the cloud points
        const RenderParams *gp = nullptr;
        if (p.M > 1 && p.LM > 0 && !(cfg->flags & RT_FLAG_PER_RAY_SHADOWS) && std::isfinite(p.cloud_lo[0] + p.cloud_lo[1] + p.cloud_lo[2]) &&
            std::isfinite(p.cloud_hi[0] + p.cloud_hi[1] + p.cloud_hi[2])) {
            if ((rc = upload_params(s, p, stream))) return rc;
            gp = s->d_params;
        }
        if (cudaEvent_t e = kernel_event(s, 0)) cudaEventRecord(e, stream);
        if (count)
            render_kernel<true><<<grid, CTA_THREADS, 0, stream>>>(p, gp);
        else
            render_kernel<false><<<grid, CTA_THREADS, 0, stream>>>(p, gp);
        CU_CHECK(cudaGetLastError());
        snprintf(s->last_kernel, sizeof(s->last_kernel), "render_kernel<%s>", count ? "true" : "false");
        if (cudaEvent_t e = kernel_event(s, 1)) cudaEventRecord(e, stream);
        s->kev_count++;
        if (launches) *launches += 1;
    } else {
        if ((rc = enqueue_direct_lighting(s, cfg, p, warps_needed, count, false, stream, launches, true))) return rc;
    }
    uint32_t n_pix = p.local_rows * p.width;
    resolve_kernel<<<(n_pix + 255) / 256, 256, 0, stream>>>(p, d_argb, d_mask, to_frame ? 1 : 0);
    CU_CHECK(cudaGetLastError());
    if (launches) *launches += 1;
    return RT_OK;
}

// after the frame's kernels: fetch the sticky device watchdog word (enqueue) / look at it (after the fetch has completed)
static int enqueue_watchdog_fetch(rt_scene *s, cudaStream_t stream) {
    if (s->frame_has_watchdog) {
        CU_CHECK(cudaMemcpyAsync(s->h_watchdog, s->d_watchdog, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
        CU_CHECK(cudaEventRecord(s->ev_watchdog, stream));
        s->watchdog_pending = true;
        s->frame_has_watchdog = false;
    }
    // end of this call's work on `stream`: what a later call on another stream waits for
    CU_CHECK(cudaEventRecord(s->ev_last, stream));
    s->last_stream = stream;
    s->has_last = true;
    return RT_OK;
}
// block = false: only looks when the fetch has already completed (cheap poll at the start of the next call)
static int check_watchdog(rt_scene *s, bool block) {
    if (!s->watchdog_pending) return RT_OK;
    if (block) {
        CU_CHECK(cudaEventSynchronize(s->ev_watchdog));
    } else {
        cudaError_t q = cudaEventQuery(s->ev_watchdog);
        if (q == cudaErrorNotReady) return RT_OK;
        if (q != cudaSuccess) return rt_fail(RT_ERR_CUDA, "cudaEventQuery failed: %s", cudaGetErrorString(q));
    }
    s->watchdog_pending = false;
    const unsigned int code = *s->h_watchdog;
    if (code) {
        *s->h_watchdog = 0;
        cudaMemset(s->d_watchdog, 0, sizeof(unsigned int));  // reported: re-arm
        return rt_fail(RT_ERR_CUDA, "direct_kernel: a wait loop gave up (code %u); a frame rendered since the last check is invalid", code);
    }
    return RT_OK;
}

static int fetch_counters(rt_scene *s, rt_stats *st) {
    unsigned long long c[16];
    CU_CHECK(cudaMemcpy(c, s->counters, sizeof(c), cudaMemcpyDeviceToHost));
    st->closest_rays = c[0];
    st->shadow_rays = c[1];
    st->box_tests = c[2];
    st->triangle_tests = c[3];
    st->sphere_tests = c[4];
    st->beam_triangle_tests = c[5];
    st->prepass_box_tests = c[8];
    st->prepass_triangle_tests = c[9];
    st->prepass_sphere_tests = c[10];
    return RT_OK;
}

// FP32 FMA throughput microbenchmark: the denominator of the FP32 roofline (MEASURED_PEAKS.json has
// no FP32 figure). 8 independent register-to-register FFMA chains per thread, all SMs busy.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f, x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 12345.678f) out[0] = r;  // keeps the chains alive
}

namespace {
struct DevBufs {  // frees its device buffers on every exit path
    std::vector<void *> ptrs;
    ~DevBufs() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <typename T>
    cudaError_t alloc(T **out, size_t bytes) {
        cudaError_t e = cudaMalloc((void **)out, bytes);
        if (e == cudaSuccess) ptrs.push_back(*out);
        return e;
    }
};
}  // namespace

static int scene_from_flat(const rtb::FlatScene &flat, int device, double build_ms, rt_scene **out_scene);
static int scene_from_device_build(const rt_scene_desc *desc, int device, rt_scene **out_scene);
static bool use_device_builder(const rt_scene_desc *desc);
constexpr int RT_BUILD_FALLBACK = 1;  // scene_from_device_build: use the host builder instead

extern "C" {

int rt_measure_fp32_peak(int device, double *out_tflops) {
    if (!out_tflops) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_measure_fp32_peak: null output");
    int rc = select_device(device);
    if (rc) return rc;
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, device));
    float *d_out = nullptr;
    CU_CHECK(cudaMalloc((void **)&d_out, 16));
    cudaEvent_t e0, e1;
    CU_CHECK(cudaEventCreate(&e0));
    CU_CHECK(cudaEventCreate(&e1));
    const int iters = 4096, grid = prop.multiProcessorCount * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CU_CHECK(cudaEventRecord(e0, 0));
        fp32_peak_kernel<<<grid, 256>>>(d_out, iters, 1.0000001f, 1e-9f);
        CU_CHECK(cudaEventRecord(e1, 0));
        CU_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        CU_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 8.0 * 16.0 * (double)iters * 256.0 * (double)grid;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *out_tflops = best;
    return RT_OK;
}

#ifdef RT_DBG_REJECT
int rt_dbg_counters(unsigned long long *out8) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, rtd::g_dbg, sizeof(z));
    cudaMemcpyToSymbol(rtd::g_dbg, z, sizeof(z));
    return 0;
}
#endif

#ifdef RT_DBG_TAIL
int rt_dbg_tail(unsigned long long *out, unsigned n_warps) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, rtd::g_dbg_tail, sizeof(unsigned long long) * 12 * n_warps);
}
#endif

int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rt_scene_create(const rt_scene_desc *desc, int device, rt_scene **out_scene) {
    if (!out_scene) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_create: null out_scene");
    *out_scene = nullptr;
    int rc = rtb::check_scene_desc(desc);
    if (rc) return rc;
    if (use_device_builder(desc)) {
        rc = scene_from_device_build(desc, device, out_scene);
        if (rc != RT_BUILD_FALLBACK) return rc;  // (tree too deep for the traversal stack: the host builder bounds its depth)
    }
    rtb::FlatScene flat;
    const auto t_build0 = std::chrono::steady_clock::now();
    rc = rtb::build_flat_scene(desc, flat);
    if (rc) return rc;
    const double build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build0).count();
    return scene_from_flat(flat, device, build_ms, out_scene);
}

}  // extern "C"

// Which builder: the host's binned-SAH build gives the better tree and costs milliseconds for the reference's own
// scenes (14.6 k primitives); from DEVICE_BUILD_MIN primitives on the build runs on the GPU (rt_bvh_device.cu).
// RT_BVH_BUILDER=host|device overrides (A/B runs, tests).
constexpr uint32_t DEVICE_BUILD_MIN = 1u << 17;
static bool use_device_builder(const rt_scene_desc *desc) {
    const uint64_t n = (uint64_t)desc->n_triangles + desc->n_spheres;
    if (const char *e = getenv("RT_BVH_BUILDER")) {
        if (!strcmp(e, "host")) return false;
        if (!strcmp(e, "device")) return n >= 16;
    }
    return n >= DEVICE_BUILD_MIN;
}

static int scene_finish(rt_scene *s);

static int scene_from_device_build(const rt_scene_desc *desc, int device, rt_scene **out_scene) {
    *out_scene = nullptr;
    int rc = select_device(device);
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    rtb::FlatScene small;  // materials and lights only
    std::vector<uint8_t> transmissive;
    rtb::pack_materials_and_lights(desc, small, transmissive);
    rtb::DeviceBvh bvh;
    if ((rc = rtb::build_scene_on_device(desc, transmissive, bvh))) return rc;
    constexpr uint32_t MAX_BVH4_DEPTH = 20;  // the traversal stack holds 3 * depth + 1 references (rt_device.cuh, rt_bvh.cpp)
    if (bvh.max_depth > MAX_BVH4_DEPTH) {
        cudaFree(bvh.nodes);
        cudaFree(bvh.prims);
        cudaFree(bvh.shade_n);
        return RT_BUILD_FALLBACK;
    }
    rt_scene *s = new (std::nothrow) rt_scene();
    if (!s) {
        cudaFree(bvh.nodes);
        cudaFree(bvh.prims);
        cudaFree(bvh.shade_n);
        return rt_fail(RT_ERR_OUT_OF_MEMORY, "out of host memory");
    }
    s->device = device;
    s->buf_nodes = bvh.nodes;
    s->buf_prims = bvh.prims;
    s->buf_shade = bvh.shade_n;
    auto fail = [&](int code) {
        rt_scene_destroy(s);
        return code;
    };
    if ((rc = upload(&s->buf_mats, small.mats))) return fail(rc);
    if ((rc = upload(&s->buf_lights, small.lights))) return fail(rc);
    s->d.nodes = (const float4 *)s->buf_nodes;
    s->d.prims = (const float4 *)s->buf_prims;
    s->d.shade_n = (const float4 *)s->buf_shade;
    s->d.mats = (const float4 *)s->buf_mats;
    s->d.lights = (const float4 *)s->buf_lights;
    s->d.n_prims = bvh.n_slots;
    s->d.n_nodes = bvh.n_nodes;
    s->d.n_lights = desc->n_lights;
    s->d.nan_hit_slot = bvh.nan_hit_slot;
    s->d.nan_hit_slot_culled = bvh.nan_hit_slot_culled;
    s->d.box_pad = bvh.box_pad;
    s->n_nodes = bvh.n_nodes;
    s->n_prims = bvh.n_slots;
    s->info.n_nodes = bvh.n_nodes;
    s->info.n_slots = bvh.n_slots;
    s->info.bvh_depth = bvh.max_depth;
    s->info.builder = 1;
    CU_CHECK(cudaDeviceSynchronize());
    s->info.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();  // incl. the upload of the primitives
    s->info.upload_ms = 0.0;
    s->info.device_bytes = (uint64_t)bvh.n_nodes * 128 + (uint64_t)bvh.n_slots * 80 + (small.mats.size() + small.lights.size()) * sizeof(float);
    if ((rc = scene_finish(s))) return fail(rc);
    *out_scene = s;
    return RT_OK;
}

// uploads a flattened scene to `device` and allocates the per-scene scratch (the second half of rt_scene_create;
// rt_multi_create builds the BVH once and runs this once per device)
static int scene_from_flat(const rtb::FlatScene &flat, int device, double build_ms, rt_scene **out_scene) {
    *out_scene = nullptr;
    const auto t_build1 = std::chrono::steady_clock::now();
    int rc = select_device(device);
    if (rc) return rc;
    rt_scene *s = new (std::nothrow) rt_scene();
    if (!s) return rt_fail(RT_ERR_OUT_OF_MEMORY, "out of host memory");
    s->device = device;
    auto fail = [&](int code) {
        rt_scene_destroy(s);
        return code;
    };
    if ((rc = upload(&s->buf_nodes, flat.nodes))) return fail(rc);
    if ((rc = upload(&s->buf_prims, flat.prims))) return fail(rc);
    if ((rc = upload(&s->buf_shade, flat.shade_n))) return fail(rc);
    if ((rc = upload(&s->buf_mats, flat.mats))) return fail(rc);
    if ((rc = upload(&s->buf_lights, flat.lights))) return fail(rc);
    s->d.nodes = (const float4 *)s->buf_nodes;
    s->d.prims = (const float4 *)s->buf_prims;
    s->d.shade_n = (const float4 *)s->buf_shade;
    s->d.mats = (const float4 *)s->buf_mats;
    s->d.lights = (const float4 *)s->buf_lights;
    s->d.n_prims = flat.n_prims;
    s->d.n_nodes = flat.n_nodes;
    s->d.n_lights = flat.n_lights;
    s->d.nan_hit_slot = flat.nan_hit_slot;
    s->d.nan_hit_slot_culled = flat.nan_hit_slot_culled;
    s->d.box_pad = flat.box_pad;
    s->n_nodes = flat.n_nodes;
    s->n_prims = flat.n_prims;
    s->info.n_nodes = flat.n_nodes;
    s->info.n_slots = flat.n_prims;
    s->info.bvh_depth = flat.max_depth;
    s->info.builder = 0;
    s->info.build_ms = build_ms;
    s->info.upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build1).count();
    s->info.device_bytes = (flat.nodes.size() + flat.prims.size() + flat.shade_n.size() + flat.mats.size() + flat.lights.size()) * sizeof(float);
    if ((rc = scene_finish(s))) return fail(rc);
    *out_scene = s;
    return RT_OK;
}

// occupancy of the render kernels on this device + the per-scene scratch every render call uses
static int scene_finish(rt_scene *s) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, s->device) != cudaSuccess) return rt_fail(RT_ERR_CUDA, "cudaGetDeviceProperties failed");
    s->sm_count = prop.multiProcessorCount;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->ctas_per_sm, render_kernel<false>, CTA_THREADS, 0) != cudaSuccess || s->ctas_per_sm < 1)
        return rt_fail(RT_ERR_CUDA, "render kernel cannot be launched on this device (built for sm_100a): %s", cudaGetErrorString(cudaGetLastError()));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->ctas_per_sm_count, render_kernel<true>, CTA_THREADS, 0) != cudaSuccess || s->ctas_per_sm_count < 1)
        s->ctas_per_sm_count = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->ctas_per_sm_direct, direct_kernel<false, true, false>, CTA_THREADS, 0) != cudaSuccess || s->ctas_per_sm_direct < 1)
        s->ctas_per_sm_direct = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->ctas_per_sm_direct_count, direct_kernel<true, true, false>, CTA_THREADS, 0) != cudaSuccess || s->ctas_per_sm_direct_count < 1)
        s->ctas_per_sm_direct_count = 1;
    if (cudaMalloc((void **)&s->work_counter, sizeof(unsigned int)) != cudaSuccess || cudaMalloc((void **)&s->pass_counters, MAX_PASSES * sizeof(unsigned int)) != cudaSuccess || cudaMalloc((void **)&s->counters, 16 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&s->ev0) != cudaSuccess ||
        cudaEventCreate(&s->ev1) != cudaSuccess || cudaMalloc((void **)&s->d_watchdog, sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(s->d_watchdog, 0, sizeof(unsigned int)) != cudaSuccess || cudaMallocHost((void **)&s->h_watchdog, sizeof(unsigned int)) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_watchdog, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_last, cudaEventDisableTiming) != cudaSuccess)
        return rt_fail(RT_ERR_CUDA, "scene scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return RT_OK;
}

extern "C" {

void rt_scene_destroy(rt_scene *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    void *bufs[] = {s->buf_nodes, s->buf_prims, s->buf_shade, s->buf_mats, s->buf_lights, s->samples, s->primary_hits, s->direct_state, s->pass_counters, s->heavy_queue, s->heavy_scratch, s->d_params, s->clouds, s->clouds_raw, s->work_counter, s->counters, s->slab_argb, s->d_watchdog, s->chain_done, s->wave_points, s->wave_contrib, s->wave_heads, s->wave_count};
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (s->h_params) cudaFreeHost(s->h_params);
    if (s->h_watchdog) cudaFreeHost(s->h_watchdog);
    if (s->h_wave_count) cudaFreeHost(s->h_wave_count);
    if (s->pinned_argb) cudaFreeHost(s->pinned_argb);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev_watchdog) cudaEventDestroy(s->ev_watchdog);
    if (s->ev_last) cudaEventDestroy(s->ev_last);
    for (int k = 0; k < 2; k++) {
        if (s->band_slab[k]) cudaFree(s->band_slab[k]);
        if (s->band_host[k]) cudaFreeHost(s->band_host[k]);
        if (s->band_rendered[k]) cudaEventDestroy(s->band_rendered[k]);
        if (s->band_copied[k]) cudaEventDestroy(s->band_copied[k]);
    }
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    for (auto &pair : s->kev)
        for (cudaEvent_t e : pair)
            if (e) cudaEventDestroy(e);
    delete s;
}

uint32_t rt_tile_local_rows(const rt_config *cfg) { return cfg ? local_rows_of(cfg) : 0; }
uint32_t rt_tile_row(const rt_config *cfg, uint32_t local_row) { return cfg ? image_row_of(cfg, local_row) : 0; }

int rt_render_device(rt_scene *s, const rt_config *cfg, void *d_out_argb, void *d_out_hit_mask, void *cuda_stream, rt_stats *stats) {
    if (!s || !d_out_argb) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_render_device: null scene or output");
    int rc = validate_config(cfg);
    if (rc) return rc;
    CU_CHECK(cudaSetDevice(s->device));
    if ((rc = check_watchdog(s, false))) return rc;  // an earlier asynchronous frame that went wrong surfaces here at the latest
    cudaStream_t stream = (cudaStream_t)cuda_stream;  // NULL = the legacy default stream, as the caller asked
    auto t0 = std::chrono::steady_clock::now();
    RenderParams p;
    uint64_t launches = 0;
    if (stats) CU_CHECK(cudaEventRecord(s->ev0, stream));
    rc = enqueue_render(s, cfg, p, (uint32_t *)d_out_argb, (uint8_t *)d_out_hit_mask, stream, &launches);
    if (rc) return rc;
    if (stats) CU_CHECK(cudaEventRecord(s->ev1, stream));
    if ((rc = enqueue_watchdog_fetch(s, stream))) return rc;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        CU_CHECK(cudaStreamSynchronize(stream));
        if ((rc = check_watchdog(s, true))) return rc;
        CU_CHECK(cudaEventSynchronize(s->ev1));
        float ms = 0;
        CU_CHECK(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        stats->kernel_ms = ms;
        stats->kernel_launches = launches;
        if (cfg->flags & RT_FLAG_COUNT_WORK) {
            rc = fetch_counters(s, stats);
            if (rc) return rc;
        }
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return RT_OK;
}

const char *rt_last_kernel(const rt_scene *s) { return s ? s->last_kernel : ""; }

int rt_scene_get_info(const rt_scene *s, rt_scene_info *out) {
    if (!s || !out) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_get_info: null argument");
    *out = s->info;
    out->scratch_bytes = s->samples_cap + s->primary_cap + s->direct_cap + s->chain_done_cap + s->wave_points_cap + s->wave_contrib_cap + s->wave_heads_cap +
                         s->heavy_scratch_cap + (s->heavy_queue ? sizeof(HeavyQueue) : 0) + s->clouds_cap + s->clouds_raw_cap + s->slab_cap;
    return RT_OK;
}

int rt_check_frame(rt_scene *s) {
    if (!s) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_check_frame: null scene");
    CU_CHECK(cudaSetDevice(s->device));
    return check_watchdog(s, true);
}

int rt_render(rt_scene *s, const rt_config *cfg, uint32_t *out_argb, uint8_t *out_hit_mask, rt_stats *stats) {
    if (!s || !out_argb) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_render: null scene or output");
    int rc = validate_config(cfg);
    if (rc) return rc;
    CU_CHECK(cudaSetDevice(s->device));
    if ((rc = check_watchdog(s, false))) return rc;
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t rows = local_rows_of(cfg);
    const size_t n_pix = (size_t)rows * cfg->width;
    if ((rc = grow(s->slab_argb, s->slab_cap, std::max<size_t>(n_pix, 1) * sizeof(uint32_t)))) return rc;
    if (s->pinned_cap < n_pix) {
        if (s->pinned_argb) cudaFreeHost(s->pinned_argb);
        s->pinned_argb = nullptr;
        s->pinned_cap = 0;
        CU_CHECK(cudaMallocHost((void **)&s->pinned_argb, std::max<size_t>(n_pix, 1) * sizeof(uint32_t)));
        s->pinned_cap = n_pix;
    }
    RenderParams p;
    uint64_t launches = 0;
    CU_CHECK(cudaEventRecord(s->ev0, s->stream));
    rc = enqueue_render(s, cfg, p, s->slab_argb, nullptr, s->stream, &launches);  // (the hit mask is slab != 0)
    if (rc) return rc;
    CU_CHECK(cudaEventRecord(s->ev1, s->stream));
    if (n_pix) CU_CHECK(cudaMemcpyAsync(s->pinned_argb, s->slab_argb, n_pix * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    if ((rc = enqueue_watchdog_fetch(s, s->stream))) return rc;
    CU_CHECK(cudaStreamSynchronize(s->stream));
    if ((rc = check_watchdog(s, true))) return rc;
    // ImageBuffer semantics (reference src/renderer/mod.rs:116-125): only hit pixels are stored.  A slab pixel is
    // 0xFFrrggbb for a hit and 0 for a miss (resolve_kernel), so the hit mask is `pixel != 0`: no second copy.
    // The merge is memory-bound on one core (0.3 ms for the 768x640 frame): large frames are split over a few threads.
    const bool zero_misses = (cfg->flags & RT_FLAG_ZERO_MISSES) != 0;
    auto merge_rows = [&](uint32_t lr0, uint32_t lr1) {
        for (uint32_t lr = lr0; lr < lr1; lr++) {
            const size_t row = image_row_of(cfg, lr);
            const uint32_t *src = s->pinned_argb + (size_t)lr * cfg->width;
            uint32_t *dst = out_argb + row * cfg->width;
            if (zero_misses)
                memcpy(dst, src, (size_t)cfg->width * sizeof(uint32_t));
            else
                for (uint32_t x = 0; x < cfg->width; x++) dst[x] = src[x] ? src[x] : dst[x];
            if (out_hit_mask) {
                uint8_t *dmask = out_hit_mask + row * cfg->width;
                for (uint32_t x = 0; x < cfg->width; x++) dmask[x] = src[x] != 0u;
            }
        }
    };
    const uint32_t merge_threads = n_pix >= (1u << 18) ? std::min(4u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
    if (merge_threads > 1) {
        std::vector<std::thread> pool;
        const uint32_t per = (rows + merge_threads - 1) / merge_threads;
        for (uint32_t t = 1; t < merge_threads; t++) pool.emplace_back(merge_rows, std::min(rows, t * per), std::min(rows, (t + 1) * per));
        merge_rows(0, std::min(rows, per));
        for (auto &th : pool) th.join();
    } else {
        merge_rows(0, rows);
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        float ms = 0;
        CU_CHECK(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        stats->kernel_ms = ms;
        stats->kernel_launches = launches;
        if (cfg->flags & RT_FLAG_COUNT_WORK) {
            rc = fetch_counters(s, stats);
            if (rc) return rc;
        }
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return RT_OK;
}

// one band of a progressive render, serial form: render, copy, merge, in order (the counter build uses it: the
// device counters are per call)
static int progressive_serial(rt_scene *s, const rt_config *cfg, uint32_t *out_argb, uint8_t *out_hit_mask, uint32_t band_rows, uint32_t n_bands,
                              rt_band_callback on_band, void *user, rt_stats *stats) {
    rt_stats total;
    memset(&total, 0, sizeof(total));
    for (uint32_t b = 0; b < n_bands; b++) {
        rt_config band = *cfg;
        band.tile_rows = band_rows;
        band.tile_rank = b;
        band.tile_world = n_bands;
        rt_stats st;
        int rc = rt_render(s, &band, out_argb, out_hit_mask, &st);
        if (rc) return rc;
        total.kernel_ms += st.kernel_ms;
        total.total_ms += st.total_ms;
        total.closest_rays += st.closest_rays;
        total.shadow_rays += st.shadow_rays;
        total.box_tests += st.box_tests;
        total.triangle_tests += st.triangle_tests;
        total.sphere_tests += st.sphere_tests;
        total.kernel_launches += st.kernel_launches;
        total.beam_triangle_tests += st.beam_triangle_tests;
        total.prepass_box_tests += st.prepass_box_tests;
        total.prepass_triangle_tests += st.prepass_triangle_tests;
        total.prepass_sphere_tests += st.prepass_sphere_tests;
        const uint32_t first = b * band_rows, n_rows = std::min(band_rows, cfg->height - first);
        if (on_band && on_band(user, first, n_rows)) break;
    }
    if (stats) *stats = total;
    return RT_OK;
}

// The reference's window re-blits the shared ImageBuffer at 60 fps while the render thread fills it
// (src/output/window.rs:73-84, src/main.rs:342-347).  Here the frame is rendered in horizontal bands with a
// two-deep pipeline: band k+1 is rendering on the scene's stream while band k's pixels travel device-to-host on a
// copy stream (two device slabs, two pinned buffers) and the host merges band k into the caller's buffer and runs
// the callback.  The GPU never waits for the host: the next band is enqueued before the host blocks on band k's copy.
int rt_render_progressive(rt_scene *s, const rt_config *cfg, uint32_t *out_argb, uint8_t *out_hit_mask, uint32_t band_rows, rt_band_callback on_band,
                          void *user, rt_stats *stats) {
    if (!s || !out_argb || band_rows == 0) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_render_progressive: null scene / output or band_rows == 0");
    int rc = validate_config(cfg);
    if (rc) return rc;
    if (cfg->tile_world > 1) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_render_progressive: cfg must not use row tiles");
    // a band = one row tile: tile b of `n_bands` tiles of band_rows rows, owned by "rank" b
    const uint32_t n_bands = (cfg->height + band_rows - 1) / band_rows, W = cfg->width;
    if (cfg->flags & RT_FLAG_COUNT_WORK) return progressive_serial(s, cfg, out_argb, out_hit_mask, band_rows, n_bands, on_band, user, stats);
    CU_CHECK(cudaSetDevice(s->device));
    if ((rc = check_watchdog(s, false))) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    const size_t band_pix = (size_t)std::min(band_rows, cfg->height) * W;
    for (int k = 0; k < 2; k++) {
        if ((rc = grow(s->band_slab[k], s->band_slab_cap[k], band_pix * sizeof(uint32_t)))) return rc;
        if (s->band_host_cap[k] < band_pix) {
            if (s->band_host[k]) cudaFreeHost(s->band_host[k]);
            s->band_host[k] = nullptr;
            s->band_host_cap[k] = 0;
            CU_CHECK(cudaMallocHost((void **)&s->band_host[k], band_pix * sizeof(uint32_t)));
            s->band_host_cap[k] = band_pix;
        }
        if (!s->band_rendered[k]) CU_CHECK(cudaEventCreateWithFlags(&s->band_rendered[k], cudaEventDisableTiming));
        if (!s->band_copied[k]) CU_CHECK(cudaEventCreateWithFlags(&s->band_copied[k], cudaEventDisableTiming));
    }
    if (!s->copy_stream) CU_CHECK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    const bool zero_misses = (cfg->flags & RT_FLAG_ZERO_MISSES) != 0;
    uint64_t launches = 0;
    uint32_t overlapped = 0;  // bands whose successor was still rendering when their host merge began
    auto enqueue_band = [&](uint32_t b) -> int {
        rt_config band = *cfg;
        band.tile_rows = band_rows;
        band.tile_rank = b;
        band.tile_world = n_bands;
        const int k = (int)(b & 1u);
        // slab k was last read by the copy of band b - 2
        if (b >= 2) CU_CHECK(cudaStreamWaitEvent(s->stream, s->band_copied[k], 0));
        RenderParams p;
        int rc2 = enqueue_render(s, &band, p, s->band_slab[k], nullptr, s->stream, &launches);
        if (rc2) return rc2;
        CU_CHECK(cudaEventRecord(s->band_rendered[k], s->stream));
        CU_CHECK(cudaStreamWaitEvent(s->copy_stream, s->band_rendered[k], 0));
        const uint32_t first = b * band_rows, n_rows = std::min(band_rows, cfg->height - first);
        CU_CHECK(cudaMemcpyAsync(s->band_host[k], s->band_slab[k], (size_t)n_rows * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->copy_stream));
        CU_CHECK(cudaEventRecord(s->band_copied[k], s->copy_stream));
        return RT_OK;
    };
    CU_CHECK(cudaEventRecord(s->ev0, s->stream));
    if ((rc = enqueue_band(0))) return rc;
    bool stopped = false;
    for (uint32_t b = 0; b < n_bands && !stopped; b++) {
        if (b + 1 < n_bands && (rc = enqueue_band(b + 1))) return rc;  // keep the GPU busy before blocking on band b
        const int k = (int)(b & 1u);
        CU_CHECK(cudaEventSynchronize(s->band_copied[k]));
        if (b + 1 < n_bands && cudaEventQuery(s->band_rendered[(b + 1) & 1u]) == cudaErrorNotReady) overlapped++;
        const uint32_t first = b * band_rows, n_rows = std::min(band_rows, cfg->height - first);
        for (uint32_t r = 0; r < n_rows; r++) {
            const uint32_t *src = s->band_host[k] + (size_t)r * W;
            uint32_t *dst = out_argb + (size_t)(first + r) * W;
            if (zero_misses)
                memcpy(dst, src, (size_t)W * sizeof(uint32_t));
            else
                for (uint32_t x = 0; x < W; x++) dst[x] = src[x] ? src[x] : dst[x];
            if (out_hit_mask) {
                uint8_t *dm = out_hit_mask + (size_t)(first + r) * W;
                for (uint32_t x = 0; x < W; x++) dm[x] = src[x] != 0u;
            }
        }
        if (on_band && on_band(user, first, n_rows)) stopped = true;  // (a band already in flight finishes; its rows stay untouched)
    }
    CU_CHECK(cudaEventRecord(s->ev1, s->stream));
    if ((rc = enqueue_watchdog_fetch(s, s->stream))) return rc;
    CU_CHECK(cudaStreamSynchronize(s->stream));
    CU_CHECK(cudaStreamSynchronize(s->copy_stream));
    if ((rc = check_watchdog(s, true))) return rc;
    s->last_progressive_overlapped = overlapped;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        float ms = 0;
        CU_CHECK(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        stats->kernel_ms = ms;
        stats->kernel_launches = launches;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return RT_OK;
}

uint32_t rt_progressive_overlapped_bands(const rt_scene *s) { return s ? s->last_progressive_overlapped : 0; }

int rt_frame_constants(const rt_config *cfg, float *out_constants, float *out_offsets, uint32_t *out_samples) {
    int rc = validate_config(cfg);
    if (rc) return rc;
    rt_scene none;  // fill_params only copies the (empty) device scene description
    RenderParams p;
    fill_params(&none, cfg, p);
    if (out_constants) {
        const float k[8] = {p.w2s_w, p.w2s_h, p.w2s_d, p.focus_x, p.focus_y, p.focus_z, p.eps_d, p.sample_scale};
        memcpy(out_constants, k, sizeof(k));
    }
    if (out_offsets) memcpy(out_offsets, p.aa_off, sizeof(float) * 2 * p.S);
    if (out_samples) *out_samples = p.S;
    return RT_OK;
}

int rt_kernel_times(rt_scene *s, uint32_t n, float *out_ms, uint32_t *out_n) {
    if (!s || !out_ms || !out_n) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_kernel_times: null argument");
    CU_CHECK(cudaSetDevice(s->device));
    const uint64_t have = std::min<uint64_t>(s->kev_count, rt_scene::KERNEL_EVENT_RING);
    const uint32_t m = (uint32_t)std::min<uint64_t>(n, have);
    for (uint32_t i = 0; i < m; i++) {  // oldest first
        const uint64_t idx = (s->kev_count - m + i) % rt_scene::KERNEL_EVENT_RING;
        if (!s->kev[idx][0] || !s->kev[idx][1]) return rt_fail(RT_ERR_CUDA, "rt_kernel_times: events unavailable");
        CU_CHECK(cudaEventSynchronize(s->kev[idx][1]));
        CU_CHECK(cudaEventElapsedTime(&out_ms[i], s->kev[idx][0], s->kev[idx][1]));
    }
    *out_n = m;
    return RT_OK;
}

int rt_scatter_tile_rows(const rt_config *cfg, const void *d_slab, void *d_frame, uint32_t elem_size, void *cuda_stream) {
    if (!cfg || !d_slab || !d_frame || (elem_size != 1 && elem_size != 4))
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_scatter_tile_rows: bad argument");
    uint32_t rows = local_rows_of(cfg);
    if (!rows) return RT_OK;
    uint32_t row_bytes = cfg->width * elem_size;
    size_t total = (size_t)row_bytes * rows;
    scatter_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(
        (const uint8_t *)d_slab, (uint8_t *)d_frame, row_bytes, rows, cfg->tile_rows, cfg->tile_rank, cfg->tile_world);
    CU_CHECK(cudaGetLastError());
    return RT_OK;
}

int rt_place_tiles(const rt_config *cfg, const void *d_slabs, uint64_t slab_stride_pixels, void *d_frame, void *d_hit_mask, void *cuda_stream) {
    if (!cfg || !d_slabs || !d_frame || cfg->tile_world == 0 || cfg->tile_rows == 0 || cfg->width == 0 || cfg->height == 0)
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_place_tiles: bad argument");
    const size_t total = (size_t)cfg->width * cfg->height;
    place_tiles_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(
        (const uint32_t *)d_slabs, (size_t)slab_stride_pixels, (uint32_t *)d_frame, (uint8_t *)d_hit_mask, cfg->width, cfg->height, cfg->tile_rows,
        cfg->tile_world);
    CU_CHECK(cudaGetLastError());
    return RT_OK;
}

int rt_trace_closest(rt_scene *s, uint32_t flags, const float *rays, uint32_t n, float *out_t, uint32_t *out_prim) {
    if (!s || (n && (!rays || !out_t || !out_prim))) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_trace_closest: null argument");
    if (!n) return RT_OK;
    CU_CHECK(cudaSetDevice(s->device));
    DevBufs bufs;
    float *d_rays = nullptr, *d_t = nullptr;
    uint32_t *d_prim = nullptr;
    CU_CHECK(bufs.alloc(&d_rays, (size_t)n * 6 * sizeof(float)));
    CU_CHECK(bufs.alloc(&d_t, (size_t)n * sizeof(float)));
    CU_CHECK(bufs.alloc(&d_prim, (size_t)n * sizeof(uint32_t)));
    CU_CHECK(cudaMemcpy(d_rays, rays, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice));
    trace_closest_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>(s->d, flags, d_rays, n, d_t, d_prim);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaStreamSynchronize(s->stream));
    CU_CHECK(cudaMemcpy(out_t, d_t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    CU_CHECK(cudaMemcpy(out_prim, d_prim, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_trace_shadow(rt_scene *s, uint32_t flags, const float *rays, const float *max_dist, uint32_t n, float *out) {
    if (!s || (n && (!rays || !max_dist || !out))) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_trace_shadow: null argument");
    if (!n) return RT_OK;
    CU_CHECK(cudaSetDevice(s->device));
    DevBufs bufs;
    float *d_rays = nullptr, *d_md = nullptr, *d_out = nullptr;
    CU_CHECK(bufs.alloc(&d_rays, (size_t)n * 6 * sizeof(float)));
    CU_CHECK(bufs.alloc(&d_md, (size_t)n * sizeof(float)));
    CU_CHECK(bufs.alloc(&d_out, (size_t)n * 5 * sizeof(float)));
    CU_CHECK(cudaMemcpy(d_rays, rays, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice));
    CU_CHECK(cudaMemcpy(d_md, max_dist, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    trace_shadow_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>(s->d, flags, d_rays, d_md, n, d_out);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaStreamSynchronize(s->stream));
    CU_CHECK(cudaMemcpy(out, d_out, (size_t)n * 5 * sizeof(float), cudaMemcpyDeviceToHost));
    return RT_OK;
}

}  // extern "C"

// =============================================================================================
// Multi-GPU inside ONE process (rt_multi_*): what a Rust host that binds this library gets when the
// box has several B200s.  The reference partitions the frame inside render() over a rayon pool
// (src/image_buffer.rs:48-97); here the partition is interleaved row tiles over the devices (tile t ->
// device t % N, scene + BVH replicated, BVH built once), one host worker thread per device so that
// all devices are launched concurrently, and the "gather" is fused into the last kernel of the
// frame: every device's resolve_kernel stores its pixels straight into the full frame on device 0
// through NVLink peer memory (P2P stores) -- no slab, no collective call, no placement pass.  Device
// 0's stream waits for the other devices' events and copies the frame to the host.  A device without
// peer access to device 0 falls back to its own slab + device-to-host copy.
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

namespace {
class WorkerPool {  // n persistent threads; run(job) executes job(r) on thread r for every r and waits
   public:
    explicit WorkerPool(int n) : n_(n) {
        for (int r = 0; r < n; r++) threads_.emplace_back([this, r] { loop(r); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void run(const std::function<void(int)> &job) {
        std::unique_lock<std::mutex> lk(m_);
        job_ = &job;
        pending_ = n_;
        gen_++;
        cv_work_.notify_all();
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
    }

   private:
    void loop(int r) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)> *job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                job = job_;
            }
            (*job)(r);
            std::lock_guard<std::mutex> lk(m_);
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)> *job_ = nullptr;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};
}  // namespace

struct rt_multi {
    std::vector<int> devices;
    std::vector<rt_scene *> scenes;     // one replica per device
    std::vector<char> peer;             // device r can store into device 0's memory (r = 0: trivially)
    std::vector<cudaEvent_t> done;      // device r's share of the frame has been written
    std::vector<uint32_t *> slab;       // fallback without peer access: compact slab on device r ...
    std::vector<size_t> slab_cap;
    std::vector<uint32_t *> slab_host;  // ... and its pinned host copy
    std::vector<size_t> slab_host_cap;
    uint32_t *frame = nullptr;          // device 0: the full frame
    size_t frame_cap = 0;
    uint32_t *frame_host = nullptr;     // pinned
    size_t frame_host_cap = 0;
    cudaStream_t stream0 = nullptr;
    std::unique_ptr<WorkerPool> pool;
    double build_ms = 0, upload_ms = 0;
};

static void multi_free(rt_multi *m) {
    if (!m) return;
    m->pool.reset();
    for (size_t r = 0; r < m->devices.size(); r++) {
        cudaSetDevice(m->devices[r]);
        if (r < m->done.size() && m->done[r]) cudaEventDestroy(m->done[r]);
        if (r < m->slab.size() && m->slab[r]) cudaFree(m->slab[r]);
        if (r < m->slab_host.size() && m->slab_host[r]) cudaFreeHost(m->slab_host[r]);
        if (r < m->scenes.size() && m->scenes[r]) rt_scene_destroy(m->scenes[r]);
    }
    if (!m->devices.empty()) cudaSetDevice(m->devices[0]);
    if (m->frame) cudaFree(m->frame);
    if (m->frame_host) cudaFreeHost(m->frame_host);
    if (m->stream0) cudaStreamDestroy(m->stream0);
    delete m;
}

extern "C" {

int rt_multi_create(const rt_scene_desc *desc, const int *devices, uint32_t n_devices, rt_multi **out) {
    if (!out) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_multi_create: null out");
    *out = nullptr;
    const int visible = rt_device_count();
    if (visible == 0) return rt_fail(RT_ERR_NO_DEVICE, "no CUDA device available; librt_b200 has no CPU fallback");
    if (n_devices == 0) n_devices = (uint32_t)visible;  // 0 = every visible device
    if (n_devices > 64) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_multi_create: %u devices", n_devices);
    rt_multi *m = new (std::nothrow) rt_multi();
    if (!m) return rt_fail(RT_ERR_OUT_OF_MEMORY, "out of host memory");
    for (uint32_t r = 0; r < n_devices; r++) {
        const int d = devices ? devices[r] : (int)r;
        if (d < 0 || d >= visible) {
            multi_free(m);
            return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_multi_create: device %d out of range (0..%d)", d, visible - 1);
        }
        for (int seen : m->devices)
            if (seen == d) {
                multi_free(m);
                return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_multi_create: device %d listed twice", d);
            }
        m->devices.push_back(d);
    }
    // Small scenes: the BVH is built once, on the host, and uploaded to every device by its own worker thread.
    // Large scenes: every device builds its own copy (rt_bvh_device.cu; deterministic, so the replicas are identical).
    int rc = rtb::check_scene_desc(desc);
    if (rc) {
        multi_free(m);
        return rc;
    }
    const bool device_build = use_device_builder(desc);
    rtb::FlatScene flat;
    bool have_flat = false;
    auto host_build = [&]() -> int {
        if (have_flat) return RT_OK;
        const auto t0 = std::chrono::steady_clock::now();
        const int rc2 = rtb::build_flat_scene(desc, flat);
        m->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        have_flat = rc2 == RT_OK;
        return rc2;
    };
    if (!device_build && (rc = host_build())) {
        multi_free(m);
        return rc;
    }
    const uint32_t n = n_devices;
    m->scenes.assign(n, nullptr);
    m->peer.assign(n, 0);
    m->done.assign(n, nullptr);
    m->slab.assign(n, nullptr);
    m->slab_cap.assign(n, 0);
    m->slab_host.assign(n, nullptr);
    m->slab_host_cap.assign(n, 0);
    m->pool.reset(new WorkerPool((int)n));
    std::vector<int> rcs(n, RT_OK);
    std::vector<std::string> errs(n);
    const auto t1 = std::chrono::steady_clock::now();
    auto create_on = [&](int r, bool from_flat) {
        int rc2 = from_flat ? scene_from_flat(flat, m->devices[r], m->build_ms, &m->scenes[r]) : scene_from_device_build(desc, m->devices[r], &m->scenes[r]);
        if (rc2 == RT_BUILD_FALLBACK) {  // tree too deep for the traversal stack: host build below
            rcs[r] = RT_BUILD_FALLBACK;
            return;
        }
        if (!rc2 && cudaEventCreateWithFlags(&m->done[r], cudaEventDisableTiming) != cudaSuccess) rc2 = rt_fail(RT_ERR_CUDA, "cudaEventCreate failed");
        if (!rc2) {
            if (r == 0) {
                m->peer[r] = 1;
            } else {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, m->devices[r], m->devices[0]) == cudaSuccess && can) {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[0], 0);
                    if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) m->peer[r] = 1;
                }
                cudaGetLastError();  // (an "already enabled" answer must not linger as the thread's last error)
                if (getenv("RT_MULTI_NO_PEER")) m->peer[r] = 0;  // tests: the fallback path
            }
        }
        rcs[r] = rc2;
        if (rc2) errs[r] = rt_last_error();
    };
    m->pool->run([&](int r) { create_on(r, !device_build); });
    bool fallback = false;
    for (uint32_t r = 0; r < n; r++) fallback |= rcs[r] == RT_BUILD_FALLBACK;
    if (fallback) {
        if ((rc = host_build())) {
            multi_free(m);
            return rc;
        }
        m->pool->run([&](int r) {
            if (rcs[r] == RT_BUILD_FALLBACK) create_on(r, true);
        });
    }
    if (device_build && !fallback && m->scenes[0]) m->build_ms = m->scenes[0]->info.build_ms;
    m->upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    for (uint32_t r = 0; r < n; r++)
        if (rcs[r]) {
            const int code = rcs[r];
            const std::string msg = errs[r];
            multi_free(m);
            return rt_fail(code, "rt_multi_create: device %u: %s", r, msg.c_str());
        }
    cudaSetDevice(m->devices[0]);
    if (cudaStreamCreateWithFlags(&m->stream0, cudaStreamNonBlocking) != cudaSuccess) {
        multi_free(m);
        return rt_fail(RT_ERR_CUDA, "rt_multi_create: stream creation failed");
    }
    *out = m;
    return RT_OK;
}

void rt_multi_destroy(rt_multi *m) { multi_free(m); }

uint32_t rt_multi_device_count(const rt_multi *m) { return m ? (uint32_t)m->devices.size() : 0; }

int rt_multi_get_info(const rt_multi *m, rt_scene_info *out, uint32_t *out_peer_devices) {
    if (!m || !out) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_multi_get_info: null argument");
    *out = m->scenes[0]->info;
    out->build_ms = m->build_ms;
    out->upload_ms = m->upload_ms;  // all devices, in parallel
    if (out_peer_devices) {
        uint32_t k = 0;
        for (char p : m->peer) k += p ? 1u : 0u;
        *out_peer_devices = k;
    }
    return RT_OK;
}

int rt_multi_render(rt_multi *m, const rt_config *cfg, uint32_t *out_argb, uint8_t *out_hit_mask, rt_stats *stats) {
    if (!m || !out_argb) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_multi_render: null handle or output");
    int rc = validate_config(cfg);
    if (rc) return rc;
    if (cfg->tile_world > 1) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_multi_render: cfg must not use row tiles (the call tiles the frame itself)");
    const auto t0 = std::chrono::steady_clock::now();
    const uint32_t n = (uint32_t)m->devices.size(), W = cfg->width, H = cfg->height;
    const uint32_t tile_rows = cfg->tile_rows ? cfg->tile_rows : 2u;
    const size_t n_pix = (size_t)W * H;
    CU_CHECK(cudaSetDevice(m->devices[0]));
    if ((rc = grow(m->frame, m->frame_cap, n_pix * sizeof(uint32_t)))) return rc;
    if (m->frame_host_cap < n_pix) {
        if (m->frame_host) cudaFreeHost(m->frame_host);
        m->frame_host = nullptr;
        m->frame_host_cap = 0;
        CU_CHECK(cudaHostAlloc((void **)&m->frame_host, n_pix * sizeof(uint32_t), cudaHostAllocPortable));
        m->frame_host_cap = n_pix;
    }
    std::vector<int> rcs(n, RT_OK);
    std::vector<std::string> errs(n);
    std::vector<rt_config> cfgs(n, *cfg);
    std::vector<uint64_t> launches(n, 0);
    m->pool->run([&](int r) {
        auto body = [&]() -> int {
            rt_scene *s = m->scenes[r];
            CU_CHECK(cudaSetDevice(m->devices[r]));
            int rc2 = check_watchdog(s, false);
            if (rc2) return rc2;
            rt_config &c = cfgs[r];
            if (n > 1) {
                c.tile_rows = tile_rows;
                c.tile_rank = (uint32_t)r;
                c.tile_world = n;
            } else {
                c.tile_rows = c.tile_rank = 0;
                c.tile_world = 1;
            }
            RenderParams p;
            CU_CHECK(cudaEventRecord(s->ev0, s->stream));
            if (m->peer[r]) {
                if ((rc2 = enqueue_render(s, &c, p, m->frame, nullptr, s->stream, &launches[r], true))) return rc2;
            } else {
                const size_t rows = local_rows_of(&c), bytes = std::max<size_t>(rows * W, 1) * sizeof(uint32_t);
                if ((rc2 = grow(m->slab[r], m->slab_cap[r], bytes))) return rc2;
                if (m->slab_host_cap[r] < bytes) {
                    if (m->slab_host[r]) cudaFreeHost(m->slab_host[r]);
                    m->slab_host[r] = nullptr;
                    m->slab_host_cap[r] = 0;
                    CU_CHECK(cudaHostAlloc((void **)&m->slab_host[r], bytes, cudaHostAllocPortable));
                    m->slab_host_cap[r] = bytes;
                }
                if ((rc2 = enqueue_render(s, &c, p, m->slab[r], nullptr, s->stream, &launches[r], false))) return rc2;
                if (rows) CU_CHECK(cudaMemcpyAsync(m->slab_host[r], m->slab[r], rows * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
            }
            CU_CHECK(cudaEventRecord(s->ev1, s->stream));
            if ((rc2 = enqueue_watchdog_fetch(s, s->stream))) return rc2;
            CU_CHECK(cudaEventRecord(m->done[r], s->stream));
            return RT_OK;
        };
        rcs[r] = body();
        if (rcs[r]) errs[r] = rt_last_error();
    });
    for (uint32_t r = 0; r < n; r++)
        if (rcs[r]) return rt_fail(rcs[r], "rt_multi_render: device %u: %s", r, errs[r].c_str());
    // device 0: once every peer has stored its rows, one device-to-host copy of the frame
    CU_CHECK(cudaSetDevice(m->devices[0]));
    bool any_peer = false;
    for (uint32_t r = 0; r < n; r++)
        if (m->peer[r]) {
            CU_CHECK(cudaStreamWaitEvent(m->stream0, m->done[r], 0));
            any_peer = true;
        }
    if (any_peer) CU_CHECK(cudaMemcpyAsync(m->frame_host, m->frame, n_pix * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream0));
    CU_CHECK(cudaStreamSynchronize(m->stream0));
    // every worker: its device's frames are valid (watchdog), then its rows go into the caller's buffer -- the merge
    // (ImageBuffer semantics: only hit pixels are stored; pixel != 0 <=> hit, see rt_render) runs on all workers at once
    const bool zero_misses = (cfg->flags & RT_FLAG_ZERO_MISSES) != 0;
    m->pool->run([&](int r) {
        auto body = [&]() -> int {
            CU_CHECK(cudaSetDevice(m->devices[r]));
            CU_CHECK(cudaEventSynchronize(m->done[r]));
            int rc2 = check_watchdog(m->scenes[r], true);
            if (rc2) return rc2;
            const uint32_t rows = local_rows_of(&cfgs[r]);
            for (uint32_t lr = 0; lr < rows; lr++) {
                const size_t row = image_row_of(&cfgs[r], lr);
                const uint32_t *src = m->peer[r] ? m->frame_host + row * W : m->slab_host[r] + (size_t)lr * W;
                uint32_t *dst = out_argb + row * W;
                if (zero_misses)
                    memcpy(dst, src, (size_t)W * sizeof(uint32_t));
                else
                    for (uint32_t x = 0; x < W; x++) dst[x] = src[x] ? src[x] : dst[x];
                if (out_hit_mask) {
                    uint8_t *dm = out_hit_mask + row * W;
                    for (uint32_t x = 0; x < W; x++) dm[x] = src[x] != 0u;
                }
            }
            return RT_OK;
        };
        rcs[r] = body();
        if (rcs[r]) errs[r] = rt_last_error();
    });
    for (uint32_t r = 0; r < n; r++)
        if (rcs[r]) return rt_fail(rcs[r], "rt_multi_render: device %u: %s", r, errs[r].c_str());
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (uint32_t r = 0; r < n; r++) {
            CU_CHECK(cudaSetDevice(m->devices[r]));
            float ms = 0;
            CU_CHECK(cudaEventElapsedTime(&ms, m->scenes[r]->ev0, m->scenes[r]->ev1));
            stats->kernel_ms = std::max(stats->kernel_ms, (double)ms);  // max over devices
            stats->kernel_launches += launches[r];
            if (cfg->flags & RT_FLAG_COUNT_WORK) {
                rt_stats st{};
                if ((rc = fetch_counters(m->scenes[r], &st))) return rc;
                stats->closest_rays += st.closest_rays;
                stats->shadow_rays += st.shadow_rays;
                stats->box_tests += st.box_tests;
                stats->triangle_tests += st.triangle_tests;
                stats->sphere_tests += st.sphere_tests;
                stats->beam_triangle_tests += st.beam_triangle_tests;
                stats->prepass_box_tests += st.prepass_box_tests;
                stats->prepass_triangle_tests += st.prepass_triangle_tests;
                stats->prepass_sphere_tests += st.prepass_sphere_tests;
            }
        }
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return RT_OK;
}

}  // extern "C"
