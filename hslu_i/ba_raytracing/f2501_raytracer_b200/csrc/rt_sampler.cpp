// rt_sampler.cpp -- seeded stand-ins for the reference's un-vendored, OS-entropy-seeded
// `fast_poisson 1.0.2` samplers (Cargo.lock; call sites src/renderer/raytracer_renderer.rs:105-127
// and src/scene/lighting/light.rs:197-206).  The reference is not reproducible run to run on
// these paths; this file produces "the first n points of a Bridson Poisson-disk
// process" with the same parameters from a documented counter-based PRNG, so that the oracle and
// the GPU consume identical tables.  Host code; runs once per scene/config, not per pixel.
#include <cmath>
#include <cstdint>
#include <vector>

#include "rt_internal.h"

namespace {

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    // uniform in [0,1) with 24 random bits
    float uniform() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    // standard normal (Box-Muller, evaluated in double, rounded once)
    float normal() {
        double u1 = ((double)(next() >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        double u2 = (double)(next() >> 11) * (1.0 / 9007199254740992.0);
        return (float)(std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2));
    }
};

// First `want` points of Bridson's algorithm in [0,dim)^N with minimum distance `radius` and
// `attempts` candidates per active point; fewer may be returned if the process dies out.
template <int N>
void bridson_first_points(SplitMix64 &rng, float dim, float radius, uint32_t attempts, uint32_t want, std::vector<float> &pts) {
    pts.clear();
    if (!want) return;
    std::vector<uint32_t> active;
    float p[N];
    for (int k = 0; k < N; k++) p[k] = rng.uniform() * dim;
    pts.insert(pts.end(), p, p + N);
    active.push_back(0);
    const float r2 = radius * radius;
    while (!active.empty() && pts.size() / N < want) {
        uint32_t ai = rng.below((uint32_t)active.size());
        const uint32_t base = active[ai];
        bool placed = false;
        for (uint32_t a = 0; a < attempts && !placed; a++) {
            float dist = radius * (1.0f + rng.uniform());  // annulus [r, 2r)
            float v[N];
            float m2 = 0.0f;
            for (int k = 0; k < N; k++) {
                v[k] = rng.normal();
                m2 += v[k] * v[k];
            }
            if (!(m2 > 0.0f)) continue;
            float f = dist / std::sqrt(m2);
            bool inside = true;
            for (int k = 0; k < N; k++) {
                p[k] = pts[(size_t)base * N + k] + v[k] * f;
                inside = inside && p[k] >= 0.0f && p[k] < dim;
            }
            if (!inside) continue;
            bool ok = true;
            for (size_t j = 0; j < pts.size() / N && ok; j++) {
                float d2 = 0.0f;
                for (int k = 0; k < N; k++) {
                    float d = pts[j * N + k] - p[k];
                    d2 += d * d;
                }
                ok = d2 >= r2;
            }
            if (!ok) continue;
            active.push_back((uint32_t)(pts.size() / N));
            pts.insert(pts.end(), p, p + N);
            placed = true;
        }
        if (!placed) {
            active[ai] = active.back();
            active.pop_back();
        }
    }
}

}  // namespace

extern "C" {

int rt_sampler_aa_table(uint64_t seed, uint32_t samples_per_pixel, int randomness, float *out_table, uint32_t table_len) {
    if (!out_table || table_len == 0 || samples_per_pixel == 0)
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_sampler_aa_table: null/empty table");
    const uint32_t total_rays = ((samples_per_pixel + 7u) / 8u) * 8u;  // next_multiple_of(8)
    std::vector<float> samples = {0.0f, 0.0f};
    for (int i = 0; i < 8; i++) {
        samples.push_back(1.0f);
        samples.push_back(1.0f);
    }
    if (randomness) {
        SplitMix64 rng(seed ^ 0xA5A5F00DCAFE1234ull);
        std::vector<float> pts;
        bridson_first_points<2>(rng, 1.2f, 3.0f / (float)total_rays, total_rays, total_rays - 1, pts);
        samples.insert(samples.end(), pts.begin(), pts.end());
    } else {
        for (uint32_t i = 0; i < total_rays; i++) {
            samples.push_back(1.0f);
            samples.push_back(1.0f);
        }
    }
    for (uint32_t i = 0; i < table_len; i++) {
        bool have = 2 * (size_t)i + 1 < samples.size() && i < total_rays;
        out_table[2 * i] = have ? samples[2 * i] : 1.0f;
        out_table[2 * i + 1] = have ? samples[2 * i + 1] : 1.0f;
    }
    return RT_OK;
}

int rt_sampler_light_clouds(uint64_t seed, uint32_t m, uint32_t n_clouds, float *out_points) {
    if (!out_points || m == 0 || n_clouds == 0) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rt_sampler_light_clouds: null/empty output");
    const float cloud_radius = 1.725f + ((float)m / 20.0f);
    const float radius = 4.0f / (float)m;
    std::vector<float> pts;
    for (uint32_t c = 0; c < n_clouds; c++) {
        SplitMix64 rng(seed * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull * (uint64_t)(c + 1));
        rng.next();
        bridson_first_points<3>(rng, cloud_radius, radius, m, m, pts);
        float *dst = out_points + (size_t)c * m * 3;
        for (uint32_t i = 0; i < m; i++) {
            if ((size_t)i * 3 + 2 < pts.size()) {
                dst[3 * i] = pts[3 * i];
                dst[3 * i + 1] = pts[3 * i + 1];
                dst[3 * i + 2] = pts[3 * i + 2];
            } else {  // pad_using(N, |_| sample_random() * cloud_radius)  (light.rs:213)
                dst[3 * i] = rng.uniform() * cloud_radius;
                dst[3 * i + 1] = rng.uniform() * cloud_radius;
                dst[3 * i + 2] = rng.uniform() * cloud_radius;
            }
        }
    }
    return RT_OK;
}

uint32_t rt_cloud_index(uint64_t seed, uint32_t x, uint32_t y, uint32_t light, uint32_t n_clouds) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(x + 1u) + 0xBF58476D1CE4E5B9ull * (uint64_t)(y + 1u) +
                 0x94D049BB133111EBull * (uint64_t)(light + 1u);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(((z >> 32) * (uint64_t)n_clouds) >> 32);
}

}  // extern "C"
