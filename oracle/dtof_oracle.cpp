// TEST INFRASTRUCTURE -- CPU oracle (see dtof_oracle.h for scope and pinning status).
//
// Scalar float32 restatement of the reference's JIT-variant algorithm for the hot path
// `dopplertofpath` + `correlated`. File:line citations refer to /root/reference.
// Compile with -ffp-contract=off: every fused multiply-add below is an explicit fmaf()
// placed where the reference writes dr::fmadd / where Dr.Jit's dot/transform helpers fuse.

#include "dtof_oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// small vector / matrix helpers (arithmetic order of Dr.Jit's static-array helpers in JIT mode)
// ------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{ x, y, z }; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
// dr::fmadd(a, s, b) on vectors
inline V3 fma3(V3 a, float s, V3 b) { return v3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }
// dr::dot, ext/drjit/include/drjit/array_router.h (fmadd chain)
inline float dot3(V3 a, V3 b) {
    float r = a.x * b.x;
    r = fmaf(a.y, b.y, r);
    r = fmaf(a.z, b.z, r);
    return r;
}
// dr::cross, array_router.h:649-659 (fmsub form)
inline V3 cross3(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
inline float rsqrt_f(float x) { return 1.f / sqrtf(x); }
// dr::normalize = v * rsqrt(squared_norm(v)), array_router.h:644-646
inline V3 normalize3(V3 a) { return a * rsqrt_f(dot3(a, a)); }
inline float mulsign(float v, float s) { return std::signbit(s) ? -v : v; }
inline float max3(V3 a) { return std::max(std::max(a.x, a.y), a.z); }

struct M34 {
    float m[12];
}; // row-major [R|t]
inline M34 load_m34(const float *p) {
    M34 r;
    memcpy(r.m, p, sizeof(r.m));
    return r;
}
// Transform::transform_affine(Point), include/mitsuba/core/transform.h:96-104
inline V3 xf_point(const M34 &M, V3 p) {
    float r[3];
    for (int i = 0; i < 3; ++i) {
        float a = M.m[4 * i + 3];
        a = fmaf(M.m[4 * i + 0], p.x, a);
        a = fmaf(M.m[4 * i + 1], p.y, a);
        a = fmaf(M.m[4 * i + 2], p.z, a);
        r[i] = a;
    }
    return v3(r[0], r[1], r[2]);
}
// Transform::operator*(Vector), transform.h (entry(0)*x, then fmadd)
inline V3 xf_vector(const M34 &M, V3 v) {
    float r[3];
    for (int i = 0; i < 3; ++i) {
        float a = M.m[4 * i + 0] * v.x;
        a = fmaf(M.m[4 * i + 1], v.y, a);
        a = fmaf(M.m[4 * i + 2], v.z, a);
        r[i] = a;
    }
    return v3(r[0], r[1], r[2]);
}
// Transform::operator*(Normal): inverse_transpose * n  ==  (M^-1)^T n; `Minv` is the inverse of M
inline V3 xf_normal(const M34 &Minv, V3 n) {
    float r[3];
    for (int i = 0; i < 3; ++i) {
        float a = Minv.m[0 + i] * n.x;
        a = fmaf(Minv.m[4 + i], n.y, a);
        a = fmaf(Minv.m[8 + i], n.z, a);
        r[i] = a;
    }
    return v3(r[0], r[1], r[2]);
}
// Inverse of an affine map. The reference inverts the full 4x4 (transform.h:56-58); for a last row
// of (0,0,0,1) this is the same map up to rounding.
inline M34 inverse_m34(const M34 &M) {
    const float *a = M.m;
    float c00 = fmaf(a[5], a[10], -(a[6] * a[9])), c01 = fmaf(a[6], a[8], -(a[4] * a[10])),
          c02 = fmaf(a[4], a[9], -(a[5] * a[8]));
    float det = fmaf(a[0], c00, fmaf(a[1], c01, a[2] * c02));
    float id = 1.f / det;
    M34 r;
    r.m[0] = c00 * id;
    r.m[1] = fmaf(a[2], a[9], -(a[1] * a[10])) * id;
    r.m[2] = fmaf(a[1], a[6], -(a[2] * a[5])) * id;
    r.m[4] = c01 * id;
    r.m[5] = fmaf(a[0], a[10], -(a[2] * a[8])) * id;
    r.m[6] = fmaf(a[2], a[4], -(a[0] * a[6])) * id;
    r.m[8] = c02 * id;
    r.m[9] = fmaf(a[1], a[8], -(a[0] * a[9])) * id;
    r.m[10] = fmaf(a[0], a[5], -(a[1] * a[4])) * id;
    for (int i = 0; i < 3; ++i) {
        float t = r.m[4 * i + 0] * a[3];
        t = fmaf(r.m[4 * i + 1], a[7], t);
        t = fmaf(r.m[4 * i + 2], a[11], t);
        r.m[4 * i + 3] = -t;
    }
    return r;
}
// AnimatedTransform::eval, transform.h:455-460:  M0 * (1 - t) + M1 * t
inline M34 lerp_m34(const M34 &A, const M34 &B, float t) {
    M34 r;
    float s = 1.f - t;
    for (int i = 0; i < 12; ++i)
        r.m[i] = fmaf(B.m[i], t, A.m[i] * s);
    return r;
}

// ------------------------------------------------------------------------------------------
// RNG: sample_tea_32 (include/mitsuba/core/random.h:77-90), PCG32 (ext/drjit/include/drjit/random.h:55-137),
// permute_kensler (random.h:235-292)
// ------------------------------------------------------------------------------------------
inline void tea32(uint32_t v0, uint32_t v1, int rounds, uint32_t &o0, uint32_t &o1) {
    uint32_t sum = 0;
    for (int i = 0; i < rounds; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    o0 = v0;
    o1 = v1;
}

struct Pcg {
    uint64_t state, inc;
    inline uint32_t next_u32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dull + inc;
        uint32_t xorshifted = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t) (old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((-(int32_t) rot) & 31));
    }
    inline float next_f32() {
        uint32_t u = (next_u32() >> 9) | 0x3f800000u;
        float f;
        memcpy(&f, &u, 4);
        return f - 1.f;
    }
    // dr::PCG32::seed(size=1, initstate, initseq), random.h:55-63
    inline void seed(uint64_t initstate, uint64_t initseq) {
        state = 0;
        inc = (initseq << 1) | 1u;
        next_u32();
        state += initstate;
        next_u32();
    }
};

inline uint32_t permute_kensler(uint32_t index, uint32_t sample_count, uint32_t seed) {
    if (sample_count == 1)
        return 0;
    uint32_t w = sample_count - 1;
    w |= w >> 1;
    w |= w >> 2;
    w |= w >> 4;
    w |= w >> 8;
    w |= w >> 16;
    do {
        uint32_t tmp = index;
        tmp ^= seed;
        tmp *= 0xe170893d;
        tmp ^= seed >> 16;
        tmp ^= (tmp & w) >> 4;
        tmp ^= seed >> 8;
        tmp *= 0x0929eb3f;
        tmp ^= seed >> 23;
        tmp ^= (tmp & w) >> 1;
        tmp *= 1 | seed >> 27;
        tmp *= 0x6935fa69;
        tmp ^= (tmp & w) >> 11;
        tmp *= 0x74dcb303;
        tmp ^= (tmp & w) >> 2;
        tmp *= 0x9e501cc3;
        tmp ^= (tmp & w) >> 2;
        tmp *= 0xc860a3df;
        tmp &= w;
        tmp ^= tmp >> 5;
        index = tmp;
    } while (index >= sample_count);
    return (index + seed) % sample_count;
}

// ------------------------------------------------------------------------------------------
// CorrelatedSampler, JIT branch (src/samplers/correlated.cpp:38-167; src/render/sampler.cpp:85-134)
// ------------------------------------------------------------------------------------------
struct LaneSampler {
    Pcg rng, rng_time, rng_path;
    uint32_t perm_seed, dim, pass, idx_mod_spp, spp_pp, sample_count, tcn, pcn;
    uint32_t draws; // draws from `rng` (diagnostic)

    void seed(const dtof_params &p, uint64_t idx64, uint32_t spp_per_pass) {
        uint32_t idx = (uint32_t) idx64;
        uint32_t S = p.base_seed + p.seed; // sampler.cpp:119, correlated.cpp:42
        uint32_t a, b;
        tea32(S, idx, 4, a, b);
        rng.seed(a, b); // sampler.cpp:126-128
        tea32(S + 1, idx / p.time_correlate_number, 4, a, b);
        rng_time.seed(a, b); // correlated.cpp:54-57
        tea32(S + 2, idx / p.path_correlate_number, 4, a, b);
        rng_path.seed(a, b);
        // compute_per_sequence_seed, sampler.cpp:85-92
        uint32_t sequence_idx = spp_per_pass * (idx / spp_per_pass);
        tea32(p.base_seed, sequence_idx + p.seed, 4, a, b);
        perm_seed = a;
        dim = 0;
        pass = 0;
        spp_pp = spp_per_pass;
        idx_mod_spp = spp_per_pass > 1 ? idx % spp_per_pass : 0;
        sample_count = p.sample_count;
        tcn = p.time_correlate_number;
        pcn = p.path_correlate_number;
        draws = 0;
    }
    void advance() { // sampler.cpp:52-55
        dim = 0;
        pass++;
    }
    uint32_t sample_index() const { return pass * spp_pp + idx_mod_spp; } // sampler.cpp:94-103

    bool stock = false; // true: the integrator calls Sampler::next_1d / next_2d (path, velocity), not *_correlate
    // next_1d_correlate, correlated.cpp:156-161: BOTH streams always advance;
    // Sampler::next_1d, correlated.cpp:78-83: the independent stream only
    float next_1d(bool correlate) {
        if (stock) {
            draws++;
            return rng.next_f32();
        }
        float r1 = rng_path.next_f32();
        float r2 = rng.next_f32();
        draws++;
        return correlate ? r1 : r2;
    }
    // next_1d_time, correlated.cpp:92-153
    float next_time(uint32_t strategy, float shift, bool strat) {
        if (strategy == DTOF_TIME_UNIFORM) {
            draws++;
            return rng.next_f32();
        }
        uint32_t si = sample_index();
        float r;
        if (strategy == DTOF_TIME_STRATIFIED) {
            r = rng.next_f32();
            draws++;
        } else {
            r = rng_time.next_f32();
        }
        if (strat) {
            int n_stratum = (int) (sample_count / tcn);
            if (strategy == DTOF_TIME_STRATIFIED) {
                uint32_t ps = perm_seed + dim++;
                uint32_t p1 = permute_kensler(si / tcn, (uint32_t) n_stratum, ps);
                ps = perm_seed + dim++;
                uint32_t p2 = permute_kensler(si / tcn, (uint32_t) n_stratum, ps);
                uint32_t pp = (si % tcn != 0) ? p1 : p2;
                r = ((float) pp + r) / (float) n_stratum;
            } else {
                uint32_t pp = si / tcn;
                r = ((float) pp + r) / (float) n_stratum;
            }
        }
        if (strategy == DTOF_TIME_STRATIFIED) {
            uint32_t pp = si % tcn;
            return ((float) pp + r) * (1.f / (float) tcn);
        } else if (strategy == DTOF_TIME_ANTITHETIC) {
            uint32_t rem = si % tcn;
            if (tcn == 2)
                return rem != 1 ? r : r + shift;
            return r + (float) rem / (float) tcn;
        } else { // ANTITHETIC_MIRROR
            uint32_t rem = si % tcn;
            float r2 = 1.0f - r + shift;
            return rem != 1 ? r : r2;
        }
    }
};

// ------------------------------------------------------------------------------------------
// Math: Dr.Jit sincos (Cephes), ext/drjit/include/drjit/math.h:76-176; fmod, array_router.h:484-486
// ------------------------------------------------------------------------------------------
inline void dr_sincos(float x, float &s_out, float &c_out) {
    float xa = fabsf(x);
    int32_t j = (int32_t) (xa * 1.2732395447351626862f);
    j = (j + 1) & ~1;
    float y = (float) j;
    uint32_t xbits;
    memcpy(&xbits, &x, 4);
    uint32_t sign_sin = ((uint32_t) j << 29) ^ xbits;
    uint32_t sign_cos = (uint32_t) (~(j - 2)) << 29;
    y = xa - y * 0.78515625f - y * 2.4187564849853515625e-4f - y * 3.77489497744594108e-8f;
    float z = y * y;
    if (xa == std::numeric_limits<float>::infinity())
        z = std::numeric_limits<float>::quiet_NaN();
    // estrin(z, c0, c1, c2) = fmadd(z*z, c2, fmadd(z, c1, c0))
    float z2 = z * z;
    float s = fmaf(z2, -1.9515295891e-4f, fmaf(z, 8.3321608736e-3f, -1.6666654611e-1f)) * z;
    float c = fmaf(z2, 2.443315711809948e-5f, fmaf(z, -1.388731625493765e-3f, 4.166664568298827e-2f)) * z;
    s = fmaf(s, y, y);
    c = fmaf(c, z, fmaf(z, -0.5f, 1.f));
    bool polymask = (j & 2) == 0;
    float rs = polymask ? s : c, rc = polymask ? c : s;
    uint32_t b;
    memcpy(&b, &rs, 4);
    b ^= (sign_sin & 0x80000000u);
    memcpy(&s_out, &b, 4);
    memcpy(&b, &rc, 4);
    b ^= (sign_cos & 0x80000000u);
    memcpy(&c_out, &b, 4);
}
inline float dr_cos(float x) {
    float s, c;
    dr_sincos(x, s, c);
    return c;
}
// dr::fmod(x, y) = fnmadd(trunc(x / y), y, x)
inline float dr_fmod(float x, float y) { return fmaf(-truncf(x / y), y, x); }

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvPi = 0.31830988618379067154f;

// eval_modulation_function_value_low_pass, include/mitsuba/render/waveform_utils.h:36-62 (JIT: float math)
inline float waveform_lowpass(float t_, uint32_t type) {
    float t = dr_fmod(t_, kTwoPi);
    switch (type) {
        case DTOF_WAVE_SINUSOIDAL: return dr_cos(t);
        case DTOF_WAVE_RECTANGULAR: {
            float a = t / kPi, b = 2.f - a, c = a < b ? a : b;
            return 2.f - 4.f * c;
        }
        case DTOF_WAVE_TRIANGULAR: {
            float a = t / kPi, b = 2.f - a, c = a < b ? a : b;
            return (4.f * c * c * c - 6.f * c * c + 1.f) * 2.f / 3.f;
        }
        case DTOF_WAVE_TRAPEZOIDAL: {
            float a = t / kPi, b = 2.f - a, c = a < b ? a : b;
            float r = 2.f - 4.f * c;
            return std::min(std::max(2.f * r, -2.f), 2.f);
        }
    }
    return dr_cos(t);
}
// eval_modulation_function_value, waveform_utils.h:24-33 (trapezoidal falls through to cos)
inline float waveform_full(float t_, uint32_t type) {
    float t = dr_fmod(t_, kTwoPi);
    switch (type) {
        case DTOF_WAVE_SINUSOIDAL: return dr_cos(t);
        case DTOF_WAVE_RECTANGULAR: return fabsf(t - kPi) > 0.5f * kPi ? 1.f : -1.f;
        case DTOF_WAVE_TRIANGULAR: return t < kPi ? 1.f - 2.f * t / kPi : -3.f + 2.f * t / kPi;
    }
    return dr_cos(t);
}

// Scalars derived once per render (the reference computes them in double, then narrows to Float:
// dopplertofpath.cpp:60-65)
struct Modulation {
    float w_g, w_d, k_phi, phase, half_g1, g1, g0, w_gd;
    uint32_t type;
    bool lowpass;
    explicit Modulation(const dtof_params &p) {
        double mhz = (double) p.w_g;
        w_g = (float) (2.0 * M_PI * mhz * 1e6);
        w_d = (float) (2.0 * M_PI / (double) p.time * (double) p.hetero_frequency);
        k_phi = (float) ((2.0 * M_PI * mhz) / 300.0);
        phase = p.sensor_phase_offset;
        half_g1 = (float) (0.5 * (double) p.g_1);
        g1 = p.g_1;
        g0 = p.g_0;
        w_gd = w_g + w_d;
        type = p.wave_function_type;
        lowpass = p.low_frequency_component_only != 0;
    }
    // eval_modulation_weight, dopplertofpath.cpp:60-77
    float eval(float ray_time, float path_length) const {
        float phi = k_phi * path_length;
        if (lowpass) {
            float t = w_d * ray_time + phase + phi;
            return half_g1 * waveform_lowpass(t, type);
        }
        // Full-waveform mode is ill-conditioned in float32 (w_g * t ~ 1e5 rad, ulp ~ 0.016 rad). The reference's
        // compiled code contracts these two statements into fused multiply-adds (-ffp-contract=fast); the oracle
        // does the same so that it can be pinned against the reference run.
        float t1 = fmaf(w_g, ray_time, -phi);
        float t2 = fmaf(w_gd, ray_time, phase);
        float g_t = g1 * waveform_full(t1, type) + g0;
        float s_t = waveform_full(t2, type);
        return s_t * g_t;
    }
};

// ------------------------------------------------------------------------------------------
// Warps / frames
// ------------------------------------------------------------------------------------------
// fresnel(cos_theta_i, eta) of a dielectric interface, include/mitsuba/render/fresnel.h:35-71
inline void fresnel_dielectric(float cos_theta_i, float eta, float &r, float &cos_theta_t, float &eta_it, float &eta_ti) {
    bool outside = cos_theta_i >= 0.f;
    float rcp_eta = 1.f / eta;
    eta_it = outside ? eta : rcp_eta;
    eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t_sqr = fmaf(-fmaf(-cos_theta_i, cos_theta_i, 1.f), eta_ti * eta_ti, 1.f);
    float ci = fabsf(cos_theta_i), ct = sqrtf(std::max(cos_theta_t_sqr, 0.f));
    bool index_matched = eta == 1.f, special = index_matched || ci == 0.f;
    float a_s = fmaf(-eta_it, ct, ci) / fmaf(eta_it, ct, ci);
    float a_p = fmaf(-eta_it, ci, ct) / fmaf(eta_it, ci, ct);
    r = 0.5f * (a_s * a_s + a_p * a_p);
    if (special)
        r = index_matched ? 0.f : 1.f;
    cos_theta_t = cos_theta_i >= 0.f ? -ct : ct; // mulsign_neg = select(v2 >= 0, -v1, v1), array_router.h:389-396
}

// ---- MicrofacetDistribution with visible-normal sampling, include/mitsuba/render/microfacet.h:185-428 -------------
// dr::erfinv (ext/drjit/include/drjit/math.h:1527-1550, after M. Giles)
inline float dr_erfinv(float x) {
    float w = -logf((1.f - x) * (1.f + x));
    float w1 = w - 2.5f, w2 = sqrtf(w) - 3.f;
    const float a[9] = { 1.50140941f, 0.246640727f, -0.00417768164f, -0.00125372503f, 0.00021858087f, -4.39150654e-06f,
                         -3.5233877e-06f, 3.43273939e-07f, 2.81022636e-08f };
    const float b[9] = { 2.83297682f, 1.00167406f, 0.00943887047f, -0.0076224613f, 0.00573950773f, -0.00367342844f,
                         0.00134934322f, 0.000100950558f, -0.000200214257f };
    float p1 = a[8], p2 = b[8];
    for (int i = 7; i >= 0; --i) {
        p1 = fmaf(p1, w1, a[i]);
        p2 = fmaf(p2, w2, b[i]);
    }
    return (w < 5.f ? p1 : p2) * x;
}
struct Microfacet {
    bool ggx;
    float au, av;
    Microfacet(bool ggx_, float au_, float av_) : ggx(ggx_), au(std::max(au_, 1e-4f)), av(std::max(av_, 1e-4f)) {}
    static float sqr(float x) { return x * x; }
    float eval(V3 m) const { // :185-210
        float alpha_uv = au * av, cos_theta = m.z, cos_theta_2 = cos_theta * cos_theta, result;
        const float pi = 3.14159265358979323846f;
        if (!ggx)
            result = expf(-(sqr(m.x / au) + sqr(m.y / av)) / cos_theta_2) / (pi * alpha_uv * sqr(cos_theta_2));
        else
            result = 1.f / (pi * alpha_uv * sqr(sqr(m.x / au) + sqr(m.y / av) + sqr(m.z)));
        return result * cos_theta > 1e-20f ? result : 0.f;
    }
    float smith_g1(V3 v, V3 m) const { // :341-365
        float xy_alpha_2 = sqr(au * v.x) + sqr(av * v.y), tan_theta_alpha_2 = xy_alpha_2 / sqr(v.z), result;
        if (!ggx) {
            float a = 1.f / sqrtf(tan_theta_alpha_2), a_sqr = a * a;
            result = a >= 1.6f ? 1.f : (3.535f * a + 2.181f * a_sqr) / (1.f + 2.276f * a + 2.577f * a_sqr);
        } else {
            result = 2.f / (1.f + sqrtf(1.f + tan_theta_alpha_2));
        }
        if (xy_alpha_2 == 0.f)
            result = 1.f;
        if (dot3(v, m) * v.z <= 0.f)
            result = 0.f;
        return result;
    }
    void sample_visible_11(float cos_theta_i, float sx, float sy, float &slope_x, float &slope_y) const { // :368-418
        if (!ggx) {
            const float inv_sqrt_pi = 0.56418958354775628695f;
            float tan_theta_i = sqrtf(std::max(fmaf(-cos_theta_i, cos_theta_i, 1.f), 0.f)) / cos_theta_i;
            float cot_theta_i = 1.f / tan_theta_i;
            float maxval = erff(cot_theta_i);
            sx = std::max(std::min(sx, 1.f - 1e-6f), 1e-6f);
            sy = std::max(std::min(sy, 1.f - 1e-6f), 1e-6f);
            float x = maxval - (maxval + 1.f) * erff(sqrtf(-logf(sx)));
            sx *= 1.f + maxval + inv_sqrt_pi * tan_theta_i * expf(-sqr(cot_theta_i));
            for (int i = 0; i < 3; ++i) {
                float slope = dr_erfinv(x);
                float value = 1.f + x + inv_sqrt_pi * tan_theta_i * expf(-sqr(slope)) - sx;
                float derivative = 1.f - slope * tan_theta_i;
                x -= value / derivative;
            }
            slope_x = dr_erfinv(x);
            slope_y = dr_erfinv(fmaf(2.f, sy, -1.f));
        } else {
            // warp::square_to_uniform_disk_concentric (warp.h:54-89)
            float px, py;
            {
                float x = fmaf(2.f, sx, -1.f), y = fmaf(2.f, sy, -1.f);
                bool is_zero = x == 0.f && y == 0.f, quadrant_1_or_3 = fabsf(x) < fabsf(y);
                float r = quadrant_1_or_3 ? y : x, rp = quadrant_1_or_3 ? x : y;
                float phi = 0.25f * 3.14159265358979323846f * rp / r;
                if (quadrant_1_or_3)
                    phi = 0.5f * 3.14159265358979323846f - phi;
                if (is_zero)
                    phi = 0.f;
                float s, c;
                dr_sincos(phi, s, c);
                px = r * c, py = r * s;
            }
            float s = 0.5f * (1.f + cos_theta_i);
            float a = sqrtf(std::max(1.f - px * px, 0.f));
            py = fmaf(py, s, fmaf(-a, s, a)); // dr::lerp(a, b, t) = fmadd(b, t, fnmadd(a, t, a))
            float x = px, y = py, z = sqrtf(std::max(1.f - (px * px + py * py), 0.f));
            float sin_theta_i = sqrtf(std::max(1.f - cos_theta_i * cos_theta_i, 0.f));
            float norm = 1.f / fmaf(sin_theta_i, y, cos_theta_i * z);
            slope_x = fmaf(cos_theta_i, y, -(sin_theta_i * z)) * norm;
            slope_y = x * norm;
        }
    }
    // visible normal sampling, :296-326; returns m and its density
    V3 sample(V3 wi, float sx, float sy, float &pdf) const {
        V3 wi_p = normalize3(v3(au * wi.x, av * wi.y, wi.z));
        // Frame3f::sincos_phi (frame.h): sin_theta_2 = x^2 + y^2, inv = rsqrt, result = (y, x) * inv, (0, 1) when tiny
        float sin_theta_2 = fmaf(wi_p.x, wi_p.x, wi_p.y * wi_p.y), inv = 1.f / sqrtf(sin_theta_2);
        float sin_phi = wi_p.y * inv, cos_phi = wi_p.x * inv;
        if (fabsf(sin_theta_2) <= 4.f * 1.1920929e-07f)
            sin_phi = 0.f, cos_phi = 1.f;
        float slx, sly;
        sample_visible_11(wi_p.z, sx, sy, slx, sly);
        float rx = fmaf(cos_phi, slx, -(sin_phi * sly)) * au, ry = fmaf(sin_phi, slx, cos_phi * sly) * av;
        V3 m = normalize3(v3(-rx, -ry, 1.f));
        pdf = eval(m) * smith_g1(wi, m) * fabsf(dot3(wi, m)) / wi.z;
        return m;
    }
};

// fresnel_diffuse_reflectance, include/mitsuba/render/fresnel.h:328-355
inline float fresnel_diffuse_reflectance(float eta) {
    float inv_eta = 1.f / eta;
    float approx_1 = fmaf(0.0636f, inv_eta, fmaf(eta, fmaf(eta, -1.4399f, 0.7099f), 0.6681f));
    const float c[6] = { 0.919317f, -3.4793f, 6.75335f, -7.80989f, 4.98554f, -1.36881f };
    float approx_2 = c[5];
    for (int i = 4; i >= 0; --i)
        approx_2 = fmaf(inv_eta, approx_2, c[i]); // dr::horner
    return eta < 1.f ? approx_1 : approx_2;
}
inline float fresnel_r(float cos_theta_i, float eta) {
    float r, ct, eit, eti;
    fresnel_dielectric(cos_theta_i, eta, r, ct, eit, eti);
    return r;
}
// constants of SmoothPlastic::parameters_changed (plastic.cpp:193-208)
struct PlasticParams {
    float eta, inv_eta_2, fdr_int, ssw;
    bool nonlinear;
    explicit PlasticParams(const dtof_bsdf &b) {
        eta = b.eta[0];
        nonlinear = b.eta[1] != 0.f;
        inv_eta_2 = 1.f / (eta * eta);
        fdr_int = fresnel_diffuse_reflectance(1.f / eta);
        float d_mean = (b.reflectance[0] + b.reflectance[1] + b.reflectance[2]) * (1.f / 3.f);
        float s_mean = (b.k[0] + b.k[1] + b.k[2]) * (1.f / 3.f);
        ssw = s_mean / (d_mean + s_mean);
    }
    V3 diff(const dtof_bsdf &b) const { // diffuse_reflectance / (1 - fdr_int [* diffuse_reflectance])
        V3 d = v3(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
        return nonlinear ? v3(d.x / (1.f - d.x * fdr_int), d.y / (1.f - d.y * fdr_int), d.z / (1.f - d.z * fdr_int))
                         : v3(d.x / (1.f - fdr_int), d.y / (1.f - fdr_int), d.z / (1.f - fdr_int));
    }
};

// fresnel_conductor, include/mitsuba/render/fresnel.h:93-117 (one colour channel)
inline float fresnel_conductor(float cos_theta_i, float eta_r, float eta_i) {
    float cos2 = cos_theta_i * cos_theta_i, sin2 = 1.f - cos2, sin4 = sin2 * sin2;
    float temp_1 = eta_r * eta_r - eta_i * eta_i - sin2;
    float a_2_pb_2 = sqrtf(std::max(temp_1 * temp_1 + 4.f * eta_i * eta_i * eta_r * eta_r, 0.f));
    float a = sqrtf(std::max(.5f * (a_2_pb_2 + temp_1), 0.f));
    float term_1 = a_2_pb_2 + cos2, term_2 = 2.f * cos_theta_i * a;
    float r_s = (term_1 - term_2) / (term_1 + term_2);
    float term_3 = a_2_pb_2 * cos2 + sin4, term_4 = term_2 * sin2;
    float r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return .5f * (r_s + r_p);
}

// warp::square_to_uniform_sphere, include/mitsuba/core/warp.h:250-255
constexpr float kInvFourPi = 0.07957747154594766788f;
inline V3 square_to_uniform_sphere(float sx, float sy) {
    float z = fmaf(-2.f, sy, 1.f);                       // fnmadd(2, y, 1)
    float r = sqrtf(std::max(fmaf(-z, z, 1.f), 0.f));    // circ(z) = safe_sqrt(fnmadd(z, z, 1))
    float s, c;
    dr_sincos(2.f * 3.14159265358979323846f * sx, s, c);
    return v3(r * c, r * s, z);
}

// square_to_uniform_disk_concentric + square_to_cosine_hemisphere, include/mitsuba/core/warp.h:54-89,320-330
inline V3 square_to_cosine_hemisphere(float sx, float sy) {
    float x = fmaf(2.f, sx, -1.f), y = fmaf(2.f, sy, -1.f);
    bool is_zero = x == 0.f && y == 0.f, q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = 0.25f * kPi * rp / r;
    if (q13)
        phi = 0.5f * kPi - phi;
    if (is_zero)
        phi = 0.f;
    float s, c;
    dr_sincos(phi, s, c);
    float px = r * c, py = r * s;
    float z = sqrtf(std::max(1.f - fmaf(py, py, px * px), 0.f));
    return v3(px, py, z);
}
// coordinate_system, include/mitsuba/core/vector.h:116-136
inline void coordinate_system(V3 n, V3 &s, V3 &t) {
    float sign = std::signbit(n.z) ? -1.f : 1.f; // dr::sign: copysign(1, x)
    float a = -1.f / (sign + n.z), b = n.x * n.y * a;
    s = v3(mulsign(n.x * n.x * a, n.z) + 1.f, mulsign(b, n.z), mulsign(-n.x, n.z));
    t = v3(b, fmaf(n.y, n.y * a, sign), -n.y);
}
// square_to_uniform_triangle, warp.h:153-156
inline void square_to_uniform_triangle(float sx, float sy, float &bx, float &by) {
    float t = sqrtf(std::max(1.f - sx, 0.f));
    bx = 1.f - t;
    by = t * sy;
}

// ------------------------------------------------------------------------------------------
// Scene
// ------------------------------------------------------------------------------------------
struct OMesh {
    std::vector<V3> pos, nrm;
    std::vector<float> uv;
    std::vector<uint32_t> faces;
    uint32_t bsdf, flip, kind;
    int32_t emitter;
    // area-emitter sampling data
    std::vector<float> area_pmf, area_cdf; // Mesh::build_pmf, src/render/mesh.cpp:361-393
    float area_sum = 0, area_norm = 0;
    uint32_t valid_lo = 0, valid_hi = 0;
    M34 rect_to_world;
    V3 rect_n, rect_s, rect_t;
    float rect_inv_area = 0;
};

struct OTri {
    V3 p0, p1, p2;
    uint32_t mesh, face;
    uint32_t gid; // position in the order of the scene description (stable under the BVH's reordering)
};

struct ONode {
    float bmin[3], bmax[3];
    uint32_t left, count; // count > 0: leaf [left, left+count) into group tri order; else children left, left+1
};

struct OGroup {
    uint32_t first_tri, n_tris;
    std::vector<ONode> nodes; // empty -> brute force
    bool animated;
    float t0, t1;
    M34 m0, m1;
};

struct Hit {
    float t, u, v;
    uint32_t inst, tri; // tri = global triangle id
};

struct Counters {
    uint64_t rays_closest = 0, rays_shadow = 0, nodes = 0, tris = 0, inst = 0, samples = 0;
};

} // namespace

struct dtof_oracle_scene {
    std::vector<OMesh> meshes;
    std::vector<OTri> tris;
    std::vector<OGroup> groups;
    std::vector<dtof_bsdf> bsdfs;
    std::vector<dtof_emitter> emitters;
    dtof_camera cam;
    dtof_film film;
    mutable Counters stats;
    // constant environment emitter (src/emitters/constant.cpp): index into `emitters` (-1: none) and its bounding
    // sphere (ConstantBackgroundEmitter::set_scene, :73-82)
    int env_emitter = -1;
    V3 env_center = v3(0, 0, 0);
    float env_radius = 1.f;
};

namespace {

void build_bvh(std::vector<OTri> &tris, OGroup &g) {
    // plain binned-SAH BVH2 over the group's triangles (oracle-only, independent of the product's builder)
    struct Ref {
        float bmin[3], bmax[3], c[3];
        OTri tri;
    };
    std::vector<Ref> refs(g.n_tris);
    for (uint32_t i = 0; i < g.n_tris; ++i) {
        const OTri &t = tris[g.first_tri + i];
        const float *p[3] = { &t.p0.x, &t.p1.x, &t.p2.x };
        for (int a = 0; a < 3; ++a) {
            refs[i].bmin[a] = std::min(p[0][a], std::min(p[1][a], p[2][a]));
            refs[i].bmax[a] = std::max(p[0][a], std::max(p[1][a], p[2][a]));
            refs[i].c[a] = 0.5f * (refs[i].bmin[a] + refs[i].bmax[a]);
        }
        refs[i].tri = t;
    }
    g.nodes.clear();
    g.nodes.reserve(2 * g.n_tris);
    g.nodes.push_back(ONode{});
    struct Job {
        uint32_t node, lo, hi;
    };
    std::vector<Job> stack{ { 0, 0, g.n_tris } };
    while (!stack.empty()) {
        Job j = stack.back();
        stack.pop_back();
        float bmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, bmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        float cmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, cmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        for (uint32_t i = j.lo; i < j.hi; ++i)
            for (int a = 0; a < 3; ++a) {
                bmin[a] = std::min(bmin[a], refs[i].bmin[a]);
                bmax[a] = std::max(bmax[a], refs[i].bmax[a]);
                cmin[a] = std::min(cmin[a], refs[i].c[a]);
                cmax[a] = std::max(cmax[a], refs[i].c[a]);
            }
        ONode &n = g.nodes[j.node];
        for (int a = 0; a < 3; ++a) {
            float pad = 1e-6f * std::max(fabsf(bmin[a]), fabsf(bmax[a])) + 1e-7f * (bmax[a] - bmin[a]) + 1e-30f;
            n.bmin[a] = bmin[a] - pad;
            n.bmax[a] = bmax[a] + pad;
        }
        uint32_t cnt = j.hi - j.lo;
        int axis = 0;
        for (int a = 1; a < 3; ++a)
            if (cmax[a] - cmin[a] > cmax[axis] - cmin[axis])
                axis = a;
        if (cnt <= 4 || cmax[axis] <= cmin[axis]) {
            n.left = j.lo;
            n.count = cnt;
            continue;
        }
        float mid = 0.5f * (cmin[axis] + cmax[axis]);
        auto it = std::partition(refs.begin() + j.lo, refs.begin() + j.hi, [&](const Ref &r) { return r.c[axis] < mid; });
        uint32_t m = (uint32_t) (it - refs.begin());
        if (m == j.lo || m == j.hi)
            m = (j.lo + j.hi) / 2;
        uint32_t l = (uint32_t) g.nodes.size();
        g.nodes.push_back(ONode{});
        g.nodes.push_back(ONode{});
        g.nodes[j.node].left = l;
        g.nodes[j.node].count = 0;
        stack.push_back({ l, j.lo, m });
        stack.push_back({ l + 1, m, j.hi });
    }
    for (uint32_t i = 0; i < g.n_tris; ++i)
        tris[g.first_tri + i] = refs[i].tri;
}

// moeller_trumbore, include/mitsuba/render/mesh.h:342-365
inline bool tri_test(V3 o, V3 d, float maxt, const OTri &tr, float &t, float &u, float &v) {
    V3 e1 = tr.p1 - tr.p0, e2 = tr.p2 - tr.p0;
    V3 pvec = cross3(d, e2);
    float inv_det = 1.f / dot3(e1, pvec);
    V3 tvec = o - tr.p0;
    u = dot3(tvec, pvec) * inv_det;
    if (!(u >= 0.f && u <= 1.f))
        return false;
    V3 qvec = cross3(tvec, e1);
    v = dot3(d, qvec) * inv_det;
    if (!(v >= 0.f && u + v <= 1.f))
        return false;
    t = dot3(e2, qvec) * inv_det;
    return t >= 0.f && t <= maxt;
}

inline bool box_test(const ONode &n, V3 o, V3 id, float maxt) {
    // slab test; fminf/fmaxf drop the NaNs of 0*inf, the far side is widened (Ize 2013) so the BVH never
    // rejects a triangle the brute-force loop accepts (boxes are also padded at build time)
    float t0 = 0.f, t1 = maxt;
    const float oo[3] = { o.x, o.y, o.z }, ii[3] = { id.x, id.y, id.z };
    for (int a = 0; a < 3; ++a) {
        float ta = (n.bmin[a] - oo[a]) * ii[a], tb = (n.bmax[a] - oo[a]) * ii[a];
        float tn = fminf(ta, tb), tf = fmaxf(ta, tb);
        t0 = fmaxf(t0, tn);
        t1 = fminf(t1, tf);
    }
    // maxt (the best hit so far) is widened too: far from the origin one ulp of t exceeds the padding of the boxes
    return t0 <= t1 * 1.000001f;
}

// Embree-style instance entry (ext/embree/kernels/common/scene_instance.h:133-138,186-206):
// world->local = rcp(lerp(M0, M1, f)), f from the UNCLAMPED relative time (default.h:225-231)
inline void enter_group(const OGroup &g, V3 o, V3 d, float time, V3 &oo, V3 &dd) {
    if (!g.animated) {
        oo = o;
        dd = d;
        return;
    }
    float f = (time - g.t0) / (g.t1 - g.t0);
    M34 inv = inverse_m34(lerp_m34(g.m0, g.m1, f));
    oo = xf_point(inv, o);
    dd = xf_vector(inv, d);
}

// Scene::ray_intersect_preliminary, src/render/scene.cpp:136-144 (closest hit; ties -> lowest triangle id)
bool intersect_closest(const dtof_oracle_scene &sc, V3 o, V3 d, float maxt, float time, Hit &hit, Counters &st) {
    hit.t = std::numeric_limits<float>::infinity();
    float best = maxt;
    bool found = false;
    st.rays_closest++;
    for (uint32_t gi = 0; gi < sc.groups.size(); ++gi) {
        const OGroup &g = sc.groups[gi];
        V3 oo, dd;
        enter_group(g, o, d, time, oo, dd);
        if (g.animated)
            st.inst++;
        auto test = [&](uint32_t ti) {
            float t, u, v;
            st.tris++;
            if (tri_test(oo, dd, best, sc.tris[ti], t, u, v)) {
                if (t < best || !found || (t == best && sc.tris[ti].gid < sc.tris[hit.tri].gid)) {
                    best = t;
                    hit = Hit{ t, u, v, gi, ti };
                    found = true;
                }
            }
        };
        if (g.nodes.empty()) {
            for (uint32_t i = 0; i < g.n_tris; ++i)
                test(g.first_tri + i);
        } else {
            V3 id = v3(1.f / dd.x, 1.f / dd.y, 1.f / dd.z);
            uint32_t stack[64], sp = 0;
            stack[sp++] = 0;
            while (sp) {
                const ONode &n = g.nodes[stack[--sp]];
                st.nodes++;
                if (!box_test(n, oo, id, best))
                    continue;
                if (n.count) {
                    for (uint32_t i = 0; i < n.count; ++i)
                        test(g.first_tri + n.left + i);
                } else {
                    stack[sp++] = n.left;
                    stack[sp++] = n.left + 1;
                }
            }
        }
    }
    return found;
}

// Scene::ray_test, src/render/scene.cpp:146-154 (any hit)
bool intersect_any(const dtof_oracle_scene &sc, V3 o, V3 d, float maxt, float time, Counters &st) {
    st.rays_shadow++;
    for (uint32_t gi = 0; gi < sc.groups.size(); ++gi) {
        const OGroup &g = sc.groups[gi];
        V3 oo, dd;
        enter_group(g, o, d, time, oo, dd);
        if (g.animated)
            st.inst++;
        float t, u, v;
        if (g.nodes.empty()) {
            for (uint32_t i = 0; i < g.n_tris; ++i) {
                st.tris++;
                if (tri_test(oo, dd, maxt, sc.tris[g.first_tri + i], t, u, v))
                    return true;
            }
        } else {
            V3 id = v3(1.f / dd.x, 1.f / dd.y, 1.f / dd.z);
            uint32_t stack[64], sp = 0;
            stack[sp++] = 0;
            while (sp) {
                const ONode &n = g.nodes[stack[--sp]];
                st.nodes++;
                if (!box_test(n, oo, id, maxt))
                    continue;
                if (n.count) {
                    for (uint32_t i = 0; i < n.count; ++i) {
                        st.tris++;
                        if (tri_test(oo, dd, maxt, sc.tris[g.first_tri + n.left + i], t, u, v))
                            return true;
                    }
                } else {
                    stack[sp++] = n.left;
                    stack[sp++] = n.left + 1;
                }
            }
        }
    }
    return false;
}

// SurfaceInteraction (the fields the path uses)
struct SI {
    bool valid;
    float t, time;
    V3 p, n, sh_n, sh_s, sh_t, dp_du, wi;
    uint32_t mesh;
};

// Mesh::compute_surface_interaction (src/render/mesh.cpp:633-789) in the group's space,
// Instance::compute_surface_interaction (src/shapes/instance.cpp:155-250) for animated groups,
// finalize_surface_interaction / initialize_sh_frame (include/mitsuba/render/interaction.h:258-268,493-516)
SI compute_si(const dtof_oracle_scene &sc, const Hit &h, V3 ray_o, V3 ray_d, float ray_time) {
    (void) ray_o;
    SI si;
    const OTri &tr = sc.tris[h.tri];
    const OMesh &m = sc.meshes[tr.mesh];
    const OGroup &g = sc.groups[h.inst];
    uint32_t f0 = m.faces[3 * tr.face], f1 = m.faces[3 * tr.face + 1], f2 = m.faces[3 * tr.face + 2];
    V3 p0 = m.pos[f0], p1 = m.pos[f1], p2 = m.pos[f2];
    float b1 = h.u, b2 = h.v, b0 = 1.f - b1 - b2;
    V3 dp0 = p1 - p0, dp1 = p2 - p0;
    // si.p = fmadd(p0, b0, fmadd(p1, b1, p2 * b2))
    si.p = fma3(p0, b0, fma3(p1, b1, p2 * b2));
    si.t = h.t;
    si.n = normalize3(cross3(dp0, dp1));
    V3 dp_du, dp_dv;
    coordinate_system(si.n, dp_du, dp_dv);
    if (!m.uv.empty()) {
        float u0x = m.uv[2 * f0], u0y = m.uv[2 * f0 + 1], u1x = m.uv[2 * f1], u1y = m.uv[2 * f1 + 1],
              u2x = m.uv[2 * f2], u2y = m.uv[2 * f2 + 1];
        float d0x = u1x - u0x, d0y = u1y - u0y, d1x = u2x - u0x, d1y = u2y - u0y;
        float det = fmaf(d0x, d1y, -(d0y * d1x));
        if (det != 0.f) {
            float inv_det = 1.f / det;
            // dp_du = fmsub(duv1.y, dp0, duv0.y * dp1) * inv_det
            dp_du = v3(fmaf(d1y, dp0.x, -(d0y * dp1.x)), fmaf(d1y, dp0.y, -(d0y * dp1.y)),
                       fmaf(d1y, dp0.z, -(d0y * dp1.z))) *
                    inv_det;
        }
    }
    if (!m.nrm.empty()) {
        V3 n0 = m.nrm[f0], n1 = m.nrm[f1], n2 = m.nrm[f2];
        V3 n = fma3(n2, b2, fma3(n1, b1, n0 * b0));
        float il = rsqrt_f(dot3(n, n));
        si.sh_n = n * il;
    } else {
        si.sh_n = si.n;
    }
    if (m.flip) {
        si.n = -si.n;
        si.sh_n = -si.sh_n;
    }
    if (g.animated) {
        // AnimatedTransform::eval (clamped), transform.h:451-456
        float tt = std::min(std::max((ray_time - g.t0) / (g.t1 - g.t0), 0.f), 1.f);
        M34 to_world = lerp_m34(g.m0, g.m1, tt);
        M34 to_object = inverse_m34(to_world);
        si.p = xf_point(to_world, si.p);
        si.n = normalize3(xf_normal(to_object, si.n));
        si.sh_n = normalize3(xf_normal(to_object, si.sh_n));
        dp_du = xf_vector(to_world, dp_du);
    }
    si.dp_du = dp_du;
    // initialize_sh_frame
    si.sh_s = normalize3(fma3(si.sh_n, -dot3(si.sh_n, dp_du), dp_du));
    if (dp_du.x == 0.f && dp_du.y == 0.f && dp_du.z == 0.f) {
        V3 tmp;
        coordinate_system(si.sh_n, si.sh_s, tmp);
    }
    si.sh_t = cross3(si.sh_n, si.sh_s);
    V3 md = -ray_d;
    si.wi = v3(dot3(md, si.sh_s), dot3(md, si.sh_t), dot3(md, si.sh_n));
    si.time = ray_time;
    si.valid = true;
    si.mesh = tr.mesh;
    return si;
}

constexpr float kRayEps = 1500.f * 5.9604644775390625e-08f; // math::RayEpsilon, include/mitsuba/core/math.h:17-22
constexpr float kShadowEps = kRayEps * 10.f;

// Interaction::offset_p, interaction.h:160-164
inline V3 offset_p(V3 p, V3 n, V3 d) {
    float mag = (1.f + std::max(std::max(fabsf(p.x), fabsf(p.y)), fabsf(p.z))) * kRayEps;
    mag = mulsign(mag, dot3(n, d));
    return fma3(n, mag, p);
}

inline float mis_weight(float a, float b) { // dopplertofpath.cpp:296-301
    a *= a;
    b *= b;
    float w = a / (a + b);
    return std::isfinite(w) ? w : 0.f;
}

struct DirSample {
    V3 p, n, d;
    float dist, pdf;
    bool delta;
    int emitter;
};

// Shape::sample_position for the emitter's shape: Rectangle::sample_position (rectangle.cpp:152-166)
// or Mesh::sample_position (mesh.cpp:518-570)
void sample_position(const OMesh &m, float sx, float sy, V3 &p, V3 &n, float &pdf) {
    if (m.kind == DTOF_SHAPE_RECTANGLE) {
        p = xf_point(m.rect_to_world, v3(sx * 2.f - 1.f, sy * 2.f - 1.f, 0.f));
        n = m.rect_n;
        pdf = m.rect_inv_area;
        return;
    }
    // DiscreteDistribution::sample_reuse on sample.y, include/mitsuba/core/distr_1d.h:117-176
    float value = sy * m.area_sum;
    uint32_t lo = m.valid_lo, hi = m.valid_hi;
    while (lo < hi) { // first index in [lo, hi] with !(cdf[i] < value)
        uint32_t mid = (lo + hi) / 2;
        if (m.area_cdf[mid] < value)
            lo = mid + 1;
        else
            hi = mid;
    }
    uint32_t face = lo;
    float pmf_n = m.area_pmf[face] * m.area_norm;
    float cdf_n = face > 0 ? m.area_cdf[face - 1] * m.area_norm : 0.f;
    sy = (sy - cdf_n) / pmf_n;
    uint32_t f0 = m.faces[3 * face], f1 = m.faces[3 * face + 1], f2 = m.faces[3 * face + 2];
    V3 p0 = m.pos[f0], p1 = m.pos[f1], p2 = m.pos[f2];
    V3 e0 = p1 - p0, e1 = p2 - p0;
    float bx, by;
    square_to_uniform_triangle(sx, sy, bx, by);
    p = fma3(e0, bx, fma3(e1, by, p0));
    pdf = m.area_norm;
    if (!m.nrm.empty()) {
        V3 n0 = m.nrm[f0], n1 = m.nrm[f1], n2 = m.nrm[f2];
        n = fma3(n0, 1.f - bx - by, fma3(n1, bx, n2 * by));
    } else {
        n = cross3(e0, e1);
    }
    n = normalize3(n);
    if (m.flip)
        n = -n;
}
inline float pdf_position(const OMesh &m) { return m.kind == DTOF_SHAPE_RECTANGLE ? m.rect_inv_area : m.area_norm; }

struct PathResult {
    V3 rgb;
    float path_length;
    uint32_t depth;
};

// DopplerToFPathIntegrator::sample, src/integrators/dopplertofpath.cpp:79-283 (JIT semantics: no early
// `break`, every draw of an active iteration is consumed -- SURVEY.md Appendix A.6)
PathResult trace_path(const dtof_oracle_scene &sc, const dtof_params &P, const Modulation &mod, LaneSampler &smp,
                      V3 ray_o, V3 ray_d, float ray_maxt, float ray_time_in, Counters &st) {
    PathResult out{ v3(0, 0, 0), 0.f, 0 };
    if (P.max_depth == 0)
        return out;
    uint32_t max_depth = (uint32_t) P.max_depth; // -1 -> 0xffffffff, integrator.cpp:573-577
    uint32_t rr_depth = (uint32_t) P.rr_depth;
    // PathIntegrator::sample (src/integrators/path.cpp:103-283) is the same loop without the time wrap, without the
    // modulation weight and with Sampler::next_1d / next_2d draws (LaneSampler::stock)
    const bool doppler = P.integrator != DTOF_INTEGRATOR_PATH;
    float ray_time = (!doppler || ray_time_in < P.time) ? ray_time_in : ray_time_in - P.time; // :93

    V3 throughput = v3(1, 1, 1), result = v3(0, 0, 0);
    float path_length = 0.f, eta = 1.f;
    uint32_t depth = 0;
    bool valid_ray = !P.hide_emitters && sc.env_emitter >= 0; // :102
    V3 prev_p = v3(0, 0, 0);
    float prev_bsdf_pdf = 1.f;
    bool prev_bsdf_delta = true;
    bool active = true;
    const uint32_t n_em = (uint32_t) sc.emitters.size();
    const float emitter_pmf = n_em ? 1.f / (float) n_em : 0.f; // Scene::m_emitter_pmf

    // loop.set_max_iterations(m_max_depth) is only a hint; the loop runs while(active)
    while (active) {
        bool correlate = doppler && (depth + 1) < P.path_correlation_depth; // :122

        Hit h;
        bool valid = intersect_closest(sc, ray_o, ray_d, ray_maxt, ray_time, h, st); // :136
        SI si{};
        si.valid = false;
        if (valid)
            si = compute_si(sc, h, ray_o, ray_d, ray_time);
        path_length += valid ? h.t * eta : 0.f; // :141

        // ---- direct emission (:150-168)
        if (valid && sc.meshes[si.mesh].emitter >= 0) {
            const OMesh &em_mesh = sc.meshes[si.mesh];
            const dtof_emitter &em = sc.emitters[em_mesh.emitter];
            // DirectionSample(scene, si, prev_si), records.h:173-180
            V3 rel = si.p - prev_p;
            float dist = sqrtf(dot3(rel, rel));
            V3 dsd = rel / dist;
            V3 dsn = si.sh_n; // PositionSample(si): n = si.sh_frame.n, records.h:76-78
            float em_pdf = 0.f;
            if (!prev_bsdf_delta) {
                // AreaLight::pdf_direction (area.cpp:148-166) -> Shape::pdf_direction (shape.cpp:389-400)
                float dp = dot3(dsd, dsn);
                float pdf = pdf_position(em_mesh), adp = fabsf(dp);
                pdf *= adp != 0.f ? (dist * dist) / adp : 0.f;
                em_pdf = (dp < 0.f ? pdf : 0.f) * emitter_pmf;
            }
            float mis_bsdf = mis_weight(prev_bsdf_pdf, em_pdf);
            float lw = doppler ? mod.eval(ray_time, path_length) : 1.f;
            // AreaLight::eval: radiance & (cos_theta(si.wi) > 0), area.cpp:82-89
            V3 Le = (si.wi.z > 0.f && prev_bsdf_pdf > 0.f) ? v3(em.value[0], em.value[1], em.value[2]) : v3(0, 0, 0);
            V3 c = Le * mis_bsdf * lw;
            result = v3(fmaf(throughput.x, c.x, result.x), fmaf(throughput.y, c.y, result.y),
                        fmaf(throughput.z, c.z, result.z));
        }

        // the ray left the scene: si.emitter(scene) is the environment emitter (scene.h:583-594). The path length is
        // not advanced (:141); DirectionSample(scene, si, prev_si).d = -si.wi = ray_d (records.h:173-180)
        if (!valid && sc.env_emitter >= 0) {
            const dtof_emitter &em = sc.emitters[sc.env_emitter];
            // Scene::pdf_emitter_direction (scene.cpp:293-299) x ConstantBackgroundEmitter::pdf_direction (:141-146)
            float em_pdf = prev_bsdf_delta ? 0.f : kInvFourPi * emitter_pmf;
            float mis_bsdf = mis_weight(prev_bsdf_pdf, em_pdf);
            float lw = doppler ? mod.eval(ray_time, path_length) : 1.f;
            V3 Le = prev_bsdf_pdf > 0.f ? v3(em.value[0], em.value[1], em.value[2]) : v3(0, 0, 0); // eval(si, active)
            V3 c = Le * mis_bsdf * lw;
            result = v3(fmaf(throughput.x, c.x, result.x), fmaf(throughput.y, c.y, result.y),
                        fmaf(throughput.z, c.z, result.z));
        }

        bool active_next = (depth + 1 < max_depth) && valid; // :171

        // ---- emitter sampling (:187-202); the 2D sample is ALWAYS drawn (Appendix A.6e)
        float e1 = smp.next_1d(correlate), e2 = smp.next_1d(correlate);
        const dtof_bsdf *bsdf = valid ? &sc.bsdfs[sc.meshes[si.mesh].bsdf] : nullptr;
        bool smooth = bsdf && (bsdf->kind == DTOF_BSDF_DIFFUSE || bsdf->kind == DTOF_BSDF_PLASTIC ||
                               bsdf->kind == DTOF_BSDF_ROUGHCONDUCTOR ||
                               bsdf->kind == DTOF_BSDF_ROUGHDIELECTRIC); // BSDFFlags::Smooth (diffuse or glossy lobe)
        bool active_em = active_next && smooth;
        DirSample ds{};
        V3 em_weight = v3(0, 0, 0), wo = v3(0, 0, 0);
        if (active_em && n_em > 0) {
            // Scene::sample_emitter_direction, scene.cpp:235-291
            uint32_t index = 0;
            float em_scale = 1.f;
            float sx = e1, sy = e2;
            if (n_em > 1) { // Scene::sample_emitter, scene.cpp:171-189
                float scaled = sx * (float) n_em;
                index = std::min((uint32_t) scaled, n_em - 1u);
                sx = scaled - (float) index;
                em_scale = (float) n_em;
            }
            const dtof_emitter &em = sc.emitters[index];
            ds.emitter = (int) index;
            bool em_active = true;
            V3 spec;
            if (em.kind == DTOF_EMITTER_POINT) { // PointLight::sample_direction, point.cpp:118-147
                ds.p = v3(em.position[0], em.position[1], em.position[2]);
                ds.n = v3(0, 0, 0);
                ds.pdf = 1.f;
                ds.delta = true;
                ds.d = ds.p - si.p;
                float dist2 = dot3(ds.d, ds.d), inv_dist = rsqrt_f(dist2);
                ds.dist = sqrtf(dist2);
                ds.d = ds.d * inv_dist;
                float f = inv_dist * inv_dist;
                spec = v3(em.value[0] * f, em.value[1] * f, em.value[2] * f);
            } else if (em.kind == DTOF_EMITTER_DIRECTIONAL) { // DirectionalEmitter::sample_direction, directional.cpp:149-176
                V3 d = v3(em.position[0], em.position[1], em.position[2]);
                V3 rel = si.p - sc.env_center;
                float radius = std::max(sc.env_radius, sqrtf(dot3(rel, rel)));
                ds.dist = 2.f * radius;
                ds.p = si.p - d * ds.dist;
                ds.n = d;
                ds.pdf = 1.f;
                ds.delta = true;
                ds.d = v3(-d.x, -d.y, -d.z);
                spec = v3(em.value[0], em.value[1], em.value[2]);
            } else if (em.kind == DTOF_EMITTER_SPOT) { // SpotLight::sample_direction (spot.cpp:180-215), falloff_curve (:146-154)
                ds.p = v3(em.position[0], em.position[1], em.position[2]);
                ds.n = v3(0, 0, 0);
                ds.pdf = 1.f;
                ds.delta = true;
                ds.d = ds.p - si.p;
                ds.dist = sqrtf(dot3(ds.d, ds.d));
                float inv_dist = 1.f / ds.dist;
                ds.d = ds.d * inv_dist;
                V3 nd = v3(-ds.d.x, -ds.d.y, -ds.d.z);
                const float *M = em.to_local; // Transform * Vector, column by column
                V3 local = v3(fmaf(M[2], nd.z, fmaf(M[1], nd.y, M[0] * nd.x)), fmaf(M[5], nd.z, fmaf(M[4], nd.y, M[3] * nd.x)),
                              fmaf(M[8], nd.z, fmaf(M[7], nd.y, M[6] * nd.x)));
                float cos_theta = normalize3(local).z;
                float cos_cutoff = cosf(em.cutoff_angle), cos_beam = cosf(em.beam_width);
                float inv_transition = 1.f / (em.cutoff_angle - em.beam_width);
                float beam_res = cos_theta >= cos_beam ? 1.f : (em.cutoff_angle - acosf(cos_theta)) * inv_transition;
                float falloff = cos_theta > cos_cutoff ? beam_res : 0.f;
                float f = falloff * (inv_dist * inv_dist);
                spec = falloff > 0.f ? v3(em.value[0] * f, em.value[1] * f, em.value[2] * f) : v3(0, 0, 0);
            } else if (em.kind == DTOF_EMITTER_CONSTANT) { // ConstantBackgroundEmitter::sample_direction, constant.cpp:112-139
                ds.d = square_to_uniform_sphere(sx, sy);
                V3 rel = si.p - sc.env_center;
                float radius = std::max(sc.env_radius, sqrtf(dot3(rel, rel))); // grows to contain the reference point
                ds.dist = 2.f * radius;
                ds.p = fma3(ds.d, ds.dist, si.p);
                ds.n = v3(-ds.d.x, -ds.d.y, -ds.d.z);
                ds.pdf = kInvFourPi;
                ds.delta = false;
                spec = v3(em.value[0] / ds.pdf, em.value[1] / ds.pdf, em.value[2] / ds.pdf);
            } else { // AreaLight::sample_direction (area.cpp:117-146) -> Shape::sample_direction (shape.cpp:370-387)
                const OMesh &em_mesh = sc.meshes[em.mesh];
                sample_position(em_mesh, sx, sy, ds.p, ds.n, ds.pdf);
                ds.delta = false;
                ds.d = ds.p - si.p;
                float dist2 = dot3(ds.d, ds.d);
                ds.dist = sqrtf(dist2);
                ds.d = ds.d / ds.dist;
                float dp = fabsf(dot3(ds.d, ds.n));
                float x = dist2 / dp;
                ds.pdf *= std::isfinite(x) ? x : 0.f;
                em_active = dot3(ds.d, ds.n) < 0.f && ds.pdf != 0.f;
                spec = em_active ? v3(em.value[0] / ds.pdf, em.value[1] / ds.pdf, em.value[2] / ds.pdf) : v3(0, 0, 0);
            }
            if (n_em > 1) {
                ds.pdf *= emitter_pmf;
                spec = spec * em_scale;
            }
            bool test = ds.pdf != 0.f;
            if (test) {
                // Interaction::spawn_ray_to, interaction.h:141-148
                V3 o = offset_p(si.p, si.n, ds.p - si.p);
                V3 d = ds.p - o;
                float dist = sqrtf(dot3(d, d));
                d = d / dist;
                if (intersect_any(sc, o, d, dist * (1.f - kShadowEps), ray_time, st)) {
                    spec = v3(0, 0, 0);
                    ds.pdf = 0.f;
                }
            }
            em_weight = spec;
            active_em = active_em && ds.pdf != 0.f; // :190
            wo = v3(dot3(ds.d, si.sh_s), dot3(ds.d, si.sh_t), dot3(ds.d, si.sh_n)); // si.to_local(ds.d)
        } else {
            active_em = false;
        }

        // ---- BSDF eval + sample (:206-210); draws ALWAYS consumed
        float s1 = smp.next_1d(correlate);
        float s2x = smp.next_1d(correlate), s2y = smp.next_1d(correlate);

        V3 bsdf_val = v3(0, 0, 0), bsdf_weight = v3(0, 0, 0), bs_wo = v3(0, 0, 0);
        float bsdf_pdf = 0.f, bs_pdf = 0.f, bs_eta = 0.f; // zero-initialised BSDFSample3f when nothing is sampled
        bool sampled_delta = false;
        if (valid && bsdf->kind == DTOF_BSDF_ROUGHDIELECTRIC) { // RoughDielectric::eval / pdf / sample, roughdielectric.cpp:240-490
            const Microfacet distr(bsdf->distribution == 1, bsdf->alpha[0], bsdf->alpha[1]);
            const float m_eta = bsdf->eta[0], m_inv_eta = 1.f / m_eta;
            const V3 wi = si.wi;
            const float cos_theta_i = wi.z;
            const V3 spec_r = v3(bsdf->reflectance[0], bsdf->reflectance[1], bsdf->reflectance[2]);
            const V3 spec_t = v3(bsdf->k[0], bsdf->k[1], bsdf->k[2]);
            const V3 wi_up = cos_theta_i >= 0.f ? wi : v3(-wi.x, -wi.y, -wi.z); // mulsign(wi, cos_theta_i)
            if (cos_theta_i != 0.f) {
                // ---- eval + pdf for the emitter sample `wo`
                const float cos_theta_o = wo.z;
                const bool reflect = cos_theta_i * cos_theta_o > 0.f;
                const float eta = cos_theta_i > 0.f ? m_eta : m_inv_eta, inv_eta = cos_theta_i > 0.f ? m_inv_eta : m_eta;
                V3 m = normalize3(wi + wo * (reflect ? 1.f : eta));
                if (std::signbit(m.z))
                    m = v3(-m.x, -m.y, -m.z); // mulsign(m, cos_theta(m))
                const float D = distr.eval(m);
                const float F = fresnel_r(dot3(wi, m), m_eta);
                const float G = distr.smith_g1(wi, m) * distr.smith_g1(wo, m);
                if (reflect) {
                    float value = F * D * G / (4.f * fabsf(cos_theta_i));
                    bsdf_val = spec_r * value;
                } else {
                    float scale = inv_eta * inv_eta;
                    float denom = dot3(wi, m) + eta * dot3(wo, m);
                    float value = fabsf((scale * (1.f - F) * D * G * eta * eta * dot3(wi, m) * dot3(wo, m)) / (cos_theta_i * (denom * denom)));
                    bsdf_val = spec_t * value;
                }
                if (dot3(wi, m) * wi.z > 0.f && dot3(wo, m) * wo.z > 0.f) { // pdf(), :432-490
                    float denom = dot3(wi, m) + eta * dot3(wo, m);
                    float dwh_dwo = reflect ? 1.f / (4.f * dot3(wo, m)) : (eta * eta * dot3(wo, m)) / (denom * denom);
                    float prob = distr.eval(m) * distr.smith_g1(wi_up, m) * fabsf(dot3(wi_up, m)) / wi_up.z; // distr.pdf(wi_up, m)
                    prob *= reflect ? F : 1.f - F;
                    bsdf_pdf = prob * fabsf(dwh_dwo);
                }
                // ---- sample
                float pdf_m;
                V3 ms = distr.sample(wi_up, s2x, s2y, pdf_m);
                bool ok = pdf_m != 0.f;
                float Fs, cos_theta_t, eta_it, eta_ti;
                fresnel_dielectric(dot3(wi, ms), m_eta, Fs, cos_theta_t, eta_it, eta_ti);
                const bool selected_r = s1 <= Fs;
                bs_pdf = pdf_m * (selected_r ? Fs : 1.f - Fs);
                float dwh;
                V3 weight;
                if (selected_r) {
                    float dwm = dot3(wi, ms);
                    bs_wo = v3(fmaf(2.f * dwm, ms.x, -wi.x), fmaf(2.f * dwm, ms.y, -wi.y), fmaf(2.f * dwm, ms.z, -wi.z)); // reflect(wi, m)
                    bs_eta = 1.f;
                    weight = spec_r;
                    dwh = 1.f / (4.f * dot3(bs_wo, ms));
                } else {
                    float c = fmaf(dot3(wi, ms), eta_ti, cos_theta_t); // refract(wi, m, cos_theta_t, eta_ti), fresnel.h:311-315
                    bs_wo = v3(fmaf(ms.x, c, -(wi.x * eta_ti)), fmaf(ms.y, c, -(wi.y * eta_ti)), fmaf(ms.z, c, -(wi.z * eta_ti)));
                    bs_eta = eta_it;
                    weight = spec_t * (eta_ti * eta_ti);
                    float denom = dot3(wi, ms) + bs_eta * dot3(bs_wo, ms);
                    dwh = (bs_eta * bs_eta * dot3(bs_wo, ms)) / (denom * denom);
                }
                weight = weight * distr.smith_g1(bs_wo, ms);
                bs_pdf *= fabsf(dwh);
                if (ok)
                    bsdf_weight = weight;
            }
        } else if (valid && bsdf->kind == DTOF_BSDF_ROUGHCONDUCTOR) { // RoughConductor::eval / pdf / sample, roughconductor.cpp:226-390
            const Microfacet distr(bsdf->distribution == 1, bsdf->alpha[0], bsdf->alpha[1]);
            V3 wi = si.wi, wo_l = wo;
            if (bsdf->twosided) {
                wo_l.z = mulsign(wo_l.z, wi.z);
                wi.z = fabsf(wi.z);
            }
            auto fresnel3 = [&](float c) {
                return v3(fresnel_conductor(c, bsdf->eta[0], bsdf->k[0]), fresnel_conductor(c, bsdf->eta[1], bsdf->k[1]),
                          fresnel_conductor(c, bsdf->eta[2], bsdf->k[2]));
            };
            const V3 spec = v3(bsdf->reflectance[0], bsdf->reflectance[1], bsdf->reflectance[2]);
            if (wi.z > 0.f && wo_l.z > 0.f) {
                V3 H = normalize3(wo_l + wi);
                float D = distr.eval(H);
                if (D != 0.f) {
                    float G = distr.smith_g1(wi, H) * distr.smith_g1(wo_l, H);
                    float result = D * G / (4.f * wi.z);
                    V3 F = fresnel3(dot3(wi, H));
                    bsdf_val = v3(F.x * (result * spec.x), F.y * (result * spec.y), F.z * (result * spec.z));
                }
                if (dot3(wi, H) > 0.f && dot3(wo_l, H) > 0.f)
                    bsdf_pdf = distr.eval(H) * distr.smith_g1(wi, H) / (4.f * wi.z);
            }
            if (wi.z > 0.f) {
                float pdf_m;
                V3 m = distr.sample(wi, s2x, s2y, pdf_m);
                float dwm = dot3(wi, m);
                bs_wo = v3(fmaf(2.f * dwm, m.x, -wi.x), fmaf(2.f * dwm, m.y, -wi.y), fmaf(2.f * dwm, m.z, -wi.z)); // reflect(wi, m)
                bs_eta = 1.f;
                bool ok = pdf_m != 0.f && bs_wo.z > 0.f;
                float weight = distr.smith_g1(bs_wo, m);
                bs_pdf = pdf_m / (4.f * dot3(bs_wo, m));
                V3 F = fresnel3(dwm);
                if (ok)
                    bsdf_weight = v3(F.x * (weight * spec.x), F.y * (weight * spec.y), F.z * (weight * spec.z));
                if (bsdf->twosided)
                    bs_wo.z = mulsign(bs_wo.z, si.wi.z);
            }
        } else if (valid && bsdf->kind == DTOF_BSDF_PLASTIC) { // SmoothPlastic::eval / pdf / sample, plastic.cpp:210-345
            const PlasticParams pp(*bsdf);
            float wi_z = si.wi.z, wo_z = wo.z;
            if (bsdf->twosided) {
                wo_z = mulsign(wo_z, wi_z);
                wi_z = fabsf(wi_z);
            }
            const float f_i = fresnel_r(wi_z, pp.eta);
            const float prob_s0 = f_i * pp.ssw, prob_d0 = (1.f - f_i) * (1.f - pp.ssw);
            if (wi_z > 0.f && wo_z > 0.f) {
                float f_o = fresnel_r(wo_z, pp.eta);
                float scale = kInvPi * wo_z * pp.inv_eta_2 * (1.f - f_i) * (1.f - f_o);
                bsdf_val = pp.diff(*bsdf) * scale;
                bsdf_pdf = kInvPi * wo_z * (prob_d0 / (prob_s0 + prob_d0));
            }
            if (wi_z > 0.f) {
                float prob_specular = prob_s0 / (prob_s0 + prob_d0), prob_diffuse = 1.f - prob_specular;
                bs_eta = 1.f;
                if (s1 < prob_specular) {
                    bs_wo = v3(-si.wi.x, -si.wi.y, wi_z); // reflect(wi) of the (flipped) incident direction
                    bs_pdf = prob_specular;
                    float value = f_i / bs_pdf;
                    bsdf_weight = v3(value * bsdf->k[0], value * bsdf->k[1], value * bsdf->k[2]);
                    sampled_delta = true;
                } else {
                    bs_wo = square_to_cosine_hemisphere(s2x, s2y);
                    bs_pdf = prob_diffuse * (kInvPi * bs_wo.z);
                    float f_o = fresnel_r(bs_wo.z, pp.eta);
                    float scale = pp.inv_eta_2 * (1.f - f_i) * (1.f - f_o) / prob_diffuse;
                    bsdf_weight = pp.diff(*bsdf) * scale;
                }
                if (bsdf->twosided)
                    bs_wo.z = mulsign(bs_wo.z, si.wi.z);
            }
        } else if (valid && smooth) {
            V3 refl = v3(bsdf->reflectance[0], bsdf->reflectance[1], bsdf->reflectance[2]);
            float wi_z = si.wi.z, wo_z = wo.z;
            if (bsdf->twosided) { // TwoSidedBRDF, twosided.cpp:111-125,219-235 (brdf[0]==brdf[1])
                wo_z = mulsign(wo_z, wi_z);
                wi_z = fabsf(wi_z);
            }
            // SmoothDiffuse::eval_pdf, diffuse.cpp:160-176
            if (wi_z > 0.f && wo_z > 0.f) {
                bsdf_val = refl * kInvPi * wo_z;
                bsdf_pdf = kInvPi * wo_z;
            }
            // SmoothDiffuse::sample, diffuse.cpp:101-125
            if (wi_z > 0.f) {
                bs_wo = square_to_cosine_hemisphere(s2x, s2y);
                bs_pdf = kInvPi * bs_wo.z;
                bs_eta = 1.f;
                if (bs_pdf > 0.f)
                    bsdf_weight = refl;
                if (bsdf->twosided)
                    bs_wo.z = mulsign(bs_wo.z, si.wi.z);
            }
        }

        if (valid && bsdf->kind == DTOF_BSDF_CONDUCTOR) { // SmoothConductor::sample, conductor.cpp:247-300
            // TwoSidedBRDF (twosided.cpp:124-127): |wi.z| goes in, the sign of wi.z is restored on wo.z
            float wi_z = bsdf->twosided ? fabsf(si.wi.z) : si.wi.z;
            if (wi_z > 0.f) {
                bs_wo = v3(-si.wi.x, -si.wi.y, si.wi.z); // reflect(wi)
                bs_pdf = 1.f;
                bs_eta = 1.f;
                bsdf_weight = v3(bsdf->reflectance[0] * fresnel_conductor(wi_z, bsdf->eta[0], bsdf->k[0]),
                                 bsdf->reflectance[1] * fresnel_conductor(wi_z, bsdf->eta[1], bsdf->k[1]),
                                 bsdf->reflectance[2] * fresnel_conductor(wi_z, bsdf->eta[2], bsdf->k[2]));
                sampled_delta = true; // bs.sampled_type = DeltaReflection
            }
        }

        bool sampled_null = false;
        if (valid && (bsdf->kind == DTOF_BSDF_DIELECTRIC || bsdf->kind == DTOF_BSDF_THINDIELECTRIC)) {
            // SmoothDielectric::sample (dielectric.cpp:250-366) / ThinDielectric::sample (thindielectric.cpp:140-189)
            const bool thin = bsdf->kind == DTOF_BSDF_THINDIELECTRIC;
            float r_i, cos_theta_t, eta_it, eta_ti;
            fresnel_dielectric(thin ? fabsf(si.wi.z) : si.wi.z, bsdf->eta[0], r_i, cos_theta_t, eta_it, eta_ti);
            if (thin)
                r_i *= 2.f / (1.f + r_i); // internal reflections: r' = r + trt + tr^3t + ..
            float t_i = 1.f - r_i;
            bool selected_r = s1 <= r_i;
            bs_pdf = selected_r ? r_i : t_i;
            if (selected_r) {
                bs_wo = v3(-si.wi.x, -si.wi.y, si.wi.z); // reflect(wi)
                bs_eta = 1.f;
                bsdf_weight = v3(bsdf->reflectance[0], bsdf->reflectance[1], bsdf->reflectance[2]);
            } else if (thin) {
                bs_wo = v3(-si.wi.x, -si.wi.y, -si.wi.z); // straight on: BSDFFlags::Null
                bs_eta = 1.f;
                bsdf_weight = v3(bsdf->k[0], bsdf->k[1], bsdf->k[2]);
                sampled_null = true;
            } else {
                bs_wo = v3(-eta_ti * si.wi.x, -eta_ti * si.wi.y, cos_theta_t); // refract(wi, cos_theta_t, eta_ti), fresnel.h
                bs_eta = eta_it;
                float f2 = eta_ti * eta_ti; // TransportMode::Radiance: solid-angle compression
                bsdf_weight = v3(bsdf->k[0] * f2, bsdf->k[1] * f2, bsdf->k[2] * f2);
            }
            sampled_delta = true;
        }

        // ---- emitter sampling contribution (:214-226)
        if (active_em) {
            float mis_em = ds.delta ? 1.f : mis_weight(ds.pdf, bsdf_pdf);
            float em_path_length = path_length + ds.dist;
            float lw = doppler ? mod.eval(ray_time, em_path_length) : 1.f;
            V3 c = bsdf_val * em_weight * mis_em * lw;
            result = v3(fmaf(throughput.x, c.x, result.x), fmaf(throughput.y, c.y, result.y),
                        fmaf(throughput.z, c.z, result.z));
        }

        // ---- BSDF sampling (:230-251)
        if (valid) {
            // si.to_world(wo) = fmadd(n, z, fmadd(t, y, s * x)), frame.h:39-41
            V3 wd = fma3(si.sh_n, bs_wo.z, fma3(si.sh_t, bs_wo.y, si.sh_s * bs_wo.x));
            ray_o = offset_p(si.p, si.n, wd); // spawn_ray, interaction.h:136-138
            ray_d = wd;
            ray_maxt = FLT_MAX;
            prev_p = si.p;
        }
        throughput = throughput * bsdf_weight;
        eta *= bs_eta;
        valid_ray = valid_ray || (valid && !sampled_null); // :253-254: a Null interaction does not make the ray valid
        prev_bsdf_pdf = bs_pdf;
        prev_bsdf_delta = sampled_delta; // has_flag(bsdf_sample.sampled_type, BSDFFlags::Delta), :250

        // ---- stopping criterion (:262-276)
        if (valid)
            depth += 1;
        float tmax = max3(throughput);
        float rr_prob = std::min(tmax * (eta * eta), 0.95f);
        bool rr_active = depth >= rr_depth;
        float q = smp.next_1d(correlate);
        bool rr_continue = q < rr_prob;
        if (rr_active)
            throughput = throughput * (1.f / rr_prob);
        active = active_next && (!rr_active || rr_continue) && tmax != 0.f;
    }
    out.rgb = valid_ray ? result : v3(0, 0, 0);
    out.path_length = path_length;
    out.depth = depth;
    return out;
}

// PerspectiveCamera::sample_ray_differential, src/sensors/perspective.cpp:238-279
inline void camera_ray(const dtof_camera &c, float u, float v, V3 &o, V3 &d, float &maxt) {
    const float *m = c.sample_to_camera;
    // Transform::operator*(Point) with projective divide, transform.h:110-118; input z = 0
    float r[4];
    for (int i = 0; i < 4; ++i) {
        float a = m[4 * i + 3];
        a = fmaf(m[4 * i + 0], u, a);
        a = fmaf(m[4 * i + 1], v, a);
        a = fmaf(m[4 * i + 2], 0.f, a);
        r[i] = a;
    }
    V3 near_p = v3(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
    V3 dl = normalize3(near_p);
    M34 tw = load_m34(c.to_world);
    o = v3(tw.m[3], tw.m[7], tw.m[11]);
    d = xf_vector(tw, dl);
    float inv_z = 1.f / dl.z;
    float near_t = c.near_clip * inv_z, far_t = c.far_clip * inv_z;
    o = o + d * near_t;
    maxt = far_t - near_t;
}

struct FilmSplat {
    float radius, inv_radius, g_alpha, g_bias;
    float ma3 = 0, ma2 = 0, ma0 = 0, mb3 = 0, mb2 = 0, mb1 = 0, mb0 = 0;   // Mitchell-Netravali polynomial coefficients
    uint32_t kind, w, h;
    int ox, oy;
    explicit FilmSplat(const dtof_film &f) {
        const float B = f.mitchell_b, Cc = f.mitchell_c;                  // mitchell.cpp:66-72
        ma3 = 12.f - 9.f * B - 6.f * Cc, ma2 = -18.f + 12.f * B + 6.f * Cc, ma0 = 6.f - 2.f * B;
        mb3 = -B - 6.f * Cc, mb2 = 6.f * B + 30.f * Cc, mb1 = -12.f * B - 48.f * Cc, mb0 = 8.f * B + 24.f * Cc;
        kind = f.rfilter;
        radius = f.rfilter_radius;
        inv_radius = 1.f / radius;
        w = f.width;
        h = f.height;
        ox = (int) f.crop_offset_x;
        oy = (int) f.crop_offset_y;
        g_alpha = -1.f / (2.f * f.gaussian_stddev * f.gaussian_stddev);
        g_bias = expf(g_alpha * radius * radius);
    }
    float eval(float x) const {
        if (kind == DTOF_RFILTER_TENT) // TentFilter::eval, src/rfilters/tent.cpp:53-55
            return std::max(0.f, 1.f - fabsf(x * inv_radius));
        if (kind == DTOF_RFILTER_MITCHELL) { // MitchellNetravaliFilter::eval, src/rfilters/mitchell.cpp:62-83
            x = fabsf(x);
            float x2 = x * x, x3 = x2 * x;
            float result = (1.f / 6.f) * (x < 1.f ? fmaf(ma3, x3, fmaf(ma2, x2, ma0)) : fmaf(mb3, x3, fmaf(mb2, x2, fmaf(mb1, x, mb0))));
            return x < 2.f ? result : 0.f;
        }
        if (kind == DTOF_RFILTER_CATMULLROM) { // CatmullRomFilter::eval, src/rfilters/catmullrom.cpp:40-55 (B = 0, C = 1/2)
            x = fabsf(x);
            float x2 = x * x, x3 = x2 * x;
            const float B = 0.f, Cc = .5f;
            float result = (1.f / 6.f) * (x < 1.f ? (12.f - 9.f * B - 6.f * Cc) * x3 + (-18.f + 12.f * B + 6.f * Cc) * x2 + (6.f - 2.f * B)
                                                  : (-B - 6.f * Cc) * x3 + (6.f * B + 30.f * Cc) * x2 + (-12.f * B - 48.f * Cc) * x +
                                                        (8.f * B + 24.f * Cc));
            return x < 2.f ? result : 0.f;
        }
        if (kind == DTOF_RFILTER_LANCZOS) { // LanczosSincFilter::eval, src/rfilters/lanczos.cpp:44-54 (radius = lobes)
            x = fabsf(x);
            float x1 = 3.14159265358979323846f * x, x2 = x1 / radius, s1, s2, c;
            dr_sincos(x1, s1, c);
            dr_sincos(x2, s2, c);
            float result = (s1 * s2) / (x1 * x2);
            return x < 0x1p-24f ? 1.f : (x > radius ? 0.f : result);
        }
        // GaussianFilter (src/rfilters/gaussian.cpp:94-103): exp(alpha x^2) - exp(alpha r^2), clamped at 0
        return std::max(0.f, expf(g_alpha * x * x) - g_bias);
    }
    // ImageBlock::put, src/render/imageblock.cpp:206-232 (box) and :418-477 (coalesced JIT path)
    template <typename Acc> void put(float px, float py, const float rgbw[4], Acc *accum) const {
        if (kind == DTOF_RFILTER_BOX) {
            int x = (int) floorf(px) - ox, y = (int) floorf(py) - oy;
            if ((uint32_t) x < w && (uint32_t) y < h)
                for (int k = 0; k < 4; ++k)
                    accum[((size_t) y * w + x) * 4 + k] += (Acc) rgbw[k];
            return;
        }
        int n = (int) ceilf(radius - .5f), count = 2 * n + 1;
        int pix = (int) floorf(px) - n, piy = (int) floorf(py) - n;
        float rx = (float) pix + .5f - px, ry = (float) piy + .5f - py;
        float wx[16], wy[16];
        for (int i = 0; i < count; ++i) {
            wx[i] = eval(rx);
            wy[i] = eval(ry);
            rx += 1.f;
            ry += 1.f;
        }
        int lx = pix - ox, ly = piy - oy;
        for (int ys = 0; ys < count; ++ys) {
            uint32_t y = (uint32_t) (ly + ys);
            if (y >= h)
                continue;
            for (int xs = 0; xs < count; ++xs) {
                uint32_t x = (uint32_t) (lx + xs);
                if (x >= w)
                    continue;
                float weight = wy[ys] * wx[xs];
                for (int k = 0; k < 4; ++k)
                    accum[((size_t) y * w + x) * 4 + k] += (Acc) (rgbw[k] * weight);
            }
        }
    }
};

int pass_info(const dtof_film &film, const dtof_params &p, dtof_pass_info *out) {
    // src/render/integrator.cpp:121-134,227-245
    uint32_t spp = p.sample_count;
    if (spp == 0)
        return 1;
    uint32_t spp_per_pass = spp, n_passes = 1;
    uint64_t wavefront = (uint64_t) film.width * film.height * spp_per_pass, limit = 0xffffffffull;
    if (wavefront > limit) {
        spp_per_pass /= (uint32_t) ((wavefront + limit - 1) / limit);
        if (spp_per_pass == 0)
            return 1;
        n_passes = spp / spp_per_pass;
        wavefront = (uint64_t) film.width * film.height * spp_per_pass;
    }
    if (spp % spp_per_pass != 0) // Sampler::set_samples_per_wavefront, sampler.cpp:81-82
        return 1;
    out->spp_per_pass = spp_per_pass;
    out->n_passes = n_passes;
    out->wavefront_size = wavefront;
    return 0;
}

// One lane, one pass: render_sample() Doppler branch, src/render/integrator.cpp:476-542
inline void lane_sample(const dtof_oracle_scene &sc, const dtof_params &P, const Modulation &mod, LaneSampler &smp,
                        uint32_t px, uint32_t py, dtof_sample_record &rec, Counters &st) {
    bool correlate_pixel = P.path_correlation_depth > 0;
    const dtof_film &f = sc.film;
    float scale_x = 1.f / (float) f.width, scale_y = 1.f / (float) f.height;
    float off_x = -(float) f.crop_offset_x * scale_x, off_y = -(float) f.crop_offset_y * scale_y;
    float posx = (float) (px + f.crop_offset_x), posy = (float) (py + f.crop_offset_y);
    const bool velocity = P.integrator == DTOF_INTEGRATOR_VELOCITY;
    const bool stock = P.integrator != DTOF_INTEGRATOR_DOPPLERTOFPATH;
    smp.stock = stock;
    // a non-Doppler integrator takes the stock branch of render_sample (src/render/integrator.cpp:409-472):
    // Sampler::next_2d / next_1d = the independent stream only (src/samplers/correlated.cpp:78-90)
    float jx, jy;
    if (stock) {
        jx = smp.rng.next_f32(), jy = smp.rng.next_f32();
        smp.draws += 2;
    } else {
        jx = smp.next_1d(correlate_pixel), jy = smp.next_1d(correlate_pixel);
    }
    float spx = posx + jx, spy = posy + jy;
    float ax = fmaf(spx, scale_x, off_x), ay = fmaf(spy, scale_y, off_y);
    float time = sc.cam.shutter_open;
    if (sc.cam.shutter_open_time > 0.f) {
        if (stock) {
            time += smp.rng.next_f32() * sc.cam.shutter_open_time;
            smp.draws++;
        } else {
            time += smp.next_time(P.time_sampling_method, P.antithetic_shift,
                                  P.use_stratified_sampling_for_each_interval != 0) *
                    sc.cam.shutter_open_time;
        }
    }
    V3 o, d;
    float maxt;
    camera_ray(sc.cam, ax, ay, o, d, maxt);
    PathResult r;
    if (velocity) {
        // VelocityIntegrator::sample, src/integrators/velocity.cpp:113-127: closest hits at ray.time = 0 and = m_time
        r.rgb = v3(0, 0, 0);
        r.path_length = 0.f;
        r.depth = 0;
        if (P.max_depth != 0) {
            Hit h1, h2;
            bool v1 = intersect_closest(sc, o, d, maxt, 0.f, h1, st);
            bool v2 = intersect_closest(sc, o, d, maxt, P.time, h2, st);
            if (v1 && v2) {
                float vel = ((v2 ? h2.t : 0.f) - (v1 ? h1.t : 0.f)) / P.time;
                r.rgb = v3(vel, vel, vel);
                r.depth = 1;
            }
        }
    } else {
        r = trace_path(sc, P, mod, smp, o, d, maxt, time, st);
    }
    bool box = f.rfilter == DTOF_RFILTER_BOX;
    rec.sample_pos[0] = box ? posx : spx;
    rec.sample_pos[1] = box ? posy : spy;
    rec.time = time;
    rec.ray_o[0] = o.x, rec.ray_o[1] = o.y, rec.ray_o[2] = o.z;
    rec.ray_d[0] = d.x, rec.ray_d[1] = d.y, rec.ray_d[2] = d.z;
    rec.ray_maxt = maxt;
    rec.rgb[0] = r.rgb.x, rec.rgb[1] = r.rgb.y, rec.rgb[2] = r.rgb.z;
    rec.path_length = r.path_length;
    rec.depth = r.depth;
    rec.rng_draws = smp.draws;
    st.samples++;
}

} // namespace

// ==========================================================================================
// C interface
// ==========================================================================================
extern "C" {

void dtof_oracle_tea32(uint32_t v0, uint32_t v1, int rounds, uint32_t out[2]) { tea32(v0, v1, rounds, out[0], out[1]); }

void dtof_oracle_pcg32(uint64_t initstate, uint64_t initseq, uint32_t n, uint32_t *out_u32, float *out_f32,
                       uint64_t *state_inc) {
    Pcg a;
    a.seed(initstate, initseq);
    if (state_inc) {
        state_inc[0] = a.state;
        state_inc[1] = a.inc;
    }
    Pcg b = a;
    for (uint32_t i = 0; i < n; ++i) {
        if (out_u32)
            out_u32[i] = a.next_u32();
        if (out_f32)
            out_f32[i] = b.next_f32();
    }
}

uint32_t dtof_oracle_permute_kensler(uint32_t index, uint32_t sample_count, uint32_t seed) {
    return permute_kensler(index, sample_count, seed);
}
void dtof_oracle_sincos(float x, float *s, float *c) { dr_sincos(x, *s, *c); }
float dtof_oracle_rfilter_eval(const dtof_film *film, float x) { return FilmSplat(*film).eval(x); }
float dtof_oracle_waveform_lowpass(float t, uint32_t type) { return waveform_lowpass(t, type); }
float dtof_oracle_waveform(float t, uint32_t type) { return waveform_full(t, type); }
float dtof_oracle_modulation_weight(const dtof_params *p, float ray_time, float path_length) {
    return Modulation(*p).eval(ray_time, path_length);
}
void dtof_oracle_square_to_cosine_hemisphere(float u, float v, float out[3]) {
    V3 r = square_to_cosine_hemisphere(u, v);
    out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void dtof_oracle_coordinate_system(const float n[3], float s[3], float t[3]) {
    V3 a, b;
    coordinate_system(v3(n[0], n[1], n[2]), a, b);
    s[0] = a.x, s[1] = a.y, s[2] = a.z, t[0] = b.x, t[1] = b.y, t[2] = b.z;
}
void dtof_oracle_seed_lane(const dtof_params *p, uint64_t idx, uint32_t spp_per_pass, uint64_t out[7]) {
    LaneSampler s;
    s.seed(*p, idx, spp_per_pass);
    out[0] = s.rng.state, out[1] = s.rng.inc, out[2] = s.rng_time.state, out[3] = s.rng_time.inc;
    out[4] = s.rng_path.state, out[5] = s.rng_path.inc, out[6] = s.perm_seed;
}
float dtof_oracle_time_sample(const dtof_params *p, uint64_t idx, uint32_t spp_per_pass, uint32_t pass) {
    LaneSampler s;
    s.seed(*p, idx, spp_per_pass);
    s.pass = pass;
    return s.next_time(p->time_sampling_method, p->antithetic_shift, p->use_stratified_sampling_for_each_interval != 0);
}
void dtof_oracle_camera_ray(const dtof_camera *cam, float u, float v, float o[3], float d[3], float *maxt) {
    V3 oo, dd;
    camera_ray(*cam, u, v, oo, dd, *maxt);
    o[0] = oo.x, o[1] = oo.y, o[2] = oo.z, d[0] = dd.x, d[1] = dd.y, d[2] = dd.z;
}
void dtof_oracle_film_put(const dtof_film *film, float px, float py, const float rgbw[4], double *accum) {
    FilmSplat(*film).put(px, py, rgbw, accum);
}
int dtof_oracle_pass_info(const dtof_scene_desc *scene, const dtof_params *p, dtof_pass_info *out) {
    return pass_info(scene->film, *p, out);
}

dtof_oracle_scene *dtof_oracle_scene_create(const dtof_scene_desc *d, int use_bvh) {
    auto *s = new dtof_oracle_scene();
    s->cam = d->camera;
    s->film = d->film;
    s->bsdfs.assign(d->bsdfs, d->bsdfs + d->n_bsdfs);
    s->emitters.assign(d->emitters, d->emitters + d->n_emitters);
    s->meshes.resize(d->n_meshes);
    for (uint32_t i = 0; i < d->n_meshes; ++i) {
        const dtof_mesh &m = d->meshes[i];
        OMesh &o = s->meshes[i];
        o.pos.resize(m.n_vertices);
        for (uint32_t v = 0; v < m.n_vertices; ++v)
            o.pos[v] = v3(m.positions[3 * v], m.positions[3 * v + 1], m.positions[3 * v + 2]);
        if (m.normals) {
            o.nrm.resize(m.n_vertices);
            for (uint32_t v = 0; v < m.n_vertices; ++v)
                o.nrm[v] = v3(m.normals[3 * v], m.normals[3 * v + 1], m.normals[3 * v + 2]);
        }
        if (m.texcoords)
            o.uv.assign(m.texcoords, m.texcoords + 2 * m.n_vertices);
        o.faces.assign(m.faces, m.faces + 3 * m.n_faces);
        o.bsdf = m.bsdf;
        o.emitter = m.emitter;
        o.flip = m.flip_normals;
        o.kind = m.kind;
        if (m.kind == DTOF_SHAPE_RECTANGLE) { // Rectangle::update, src/shapes/rectangle.cpp:101-113
            o.rect_to_world = load_m34(m.rect_to_world);
            o.rect_s = xf_vector(o.rect_to_world, v3(2, 0, 0));
            o.rect_t = xf_vector(o.rect_to_world, v3(0, 2, 0));
            o.rect_n = normalize3(xf_normal(inverse_m34(o.rect_to_world), v3(0, 0, 1)));
            V3 c = cross3(o.rect_s, o.rect_t);
            o.rect_inv_area = 1.f / sqrtf(dot3(c, c));
        }
        if (m.emitter >= 0 && m.kind == DTOF_SHAPE_MESH) { // Mesh::build_pmf + DiscreteDistribution::compute_cdf
            o.area_pmf.resize(m.n_faces);
            o.area_cdf.resize(m.n_faces);
            double sum = 0.0;
            bool any = false;
            for (uint32_t f = 0; f < m.n_faces; ++f) {
                V3 p0 = o.pos[o.faces[3 * f]], p1 = o.pos[o.faces[3 * f + 1]], p2 = o.pos[o.faces[3 * f + 2]];
                V3 c = cross3(p1 - p0, p2 - p0);
                float a = .5f * sqrtf(dot3(c, c));
                o.area_pmf[f] = a;
                sum += (double) a;
                o.area_cdf[f] = (float) sum;
                if (a > 0.f) {
                    if (!any)
                        o.valid_lo = f;
                    o.valid_hi = f;
                    any = true;
                }
            }
            o.area_sum = (float) sum;
            o.area_norm = (float) (1.0 / sum);
        }
    }
    s->groups.resize(d->n_instances);
    for (uint32_t i = 0; i < d->n_instances; ++i) {
        const dtof_instance &in = d->instances[i];
        OGroup &g = s->groups[i];
        g.first_tri = (uint32_t) s->tris.size();
        for (uint32_t mi = in.first_mesh; mi < in.first_mesh + in.n_meshes; ++mi) {
            const OMesh &m = s->meshes[mi];
            for (uint32_t f = 0; f < m.faces.size() / 3; ++f)
                s->tris.push_back(OTri{ m.pos[m.faces[3 * f]], m.pos[m.faces[3 * f + 1]], m.pos[m.faces[3 * f + 2]], mi, f,
                                        (uint32_t) s->tris.size() });
        }
        g.n_tris = (uint32_t) s->tris.size() - g.first_tri;
        g.animated = in.animated != 0;
        g.t0 = in.t0;
        g.t1 = in.t1;
        g.m0 = load_m34(in.m0);
        g.m1 = load_m34(in.m1);
        bool bvh = use_bvh > 0 || (use_bvh < 0 && g.n_tris > 64);
        if (bvh && g.n_tris > 0)
            build_bvh(s->tris, g);
    }
    bool need_bsphere = false;
    for (uint32_t i = 0; i < d->n_emitters; ++i) {
        if (d->emitters[i].kind == DTOF_EMITTER_CONSTANT)
            s->env_emitter = (int) i;
        need_bsphere = need_bsphere || d->emitters[i].kind == DTOF_EMITTER_CONSTANT || d->emitters[i].kind == DTOF_EMITTER_DIRECTIONAL;
    }
    if (need_bsphere) { // ConstantBackgroundEmitter / DirectionalEmitter::set_scene (constant.cpp:73-82, directional.cpp:98-108)
        // Scene::bbox (scene.cpp:36) = union of the shapes' boxes: static shapes by their vertices, an instance by the
        // 8 corners of its group's box under both keyframes (instance.cpp:101-114); then the bounding sphere with
        // radius * (1 + RayEpsilon), at least RayEpsilon (constant.cpp:73-82); an empty scene gives ((0,0,0), 1)
        V3 lo = v3(INFINITY, INFINITY, INFINITY), hi = v3(-INFINITY, -INFINITY, -INFINITY);
        auto grow = [&](V3 q) {
            lo = v3(std::min(lo.x, q.x), std::min(lo.y, q.y), std::min(lo.z, q.z));
            hi = v3(std::max(hi.x, q.x), std::max(hi.y, q.y), std::max(hi.z, q.z));
        };
        for (uint32_t i = 0; i < d->n_instances; ++i) {
            const dtof_instance &in = d->instances[i];
            V3 glo = v3(INFINITY, INFINITY, INFINITY), ghi = v3(-INFINITY, -INFINITY, -INFINITY);
            bool any = false;
            for (uint32_t mi = in.first_mesh; mi < in.first_mesh + in.n_meshes; ++mi)
                for (const V3 &q : s->meshes[mi].pos) {
                    glo = v3(std::min(glo.x, q.x), std::min(glo.y, q.y), std::min(glo.z, q.z));
                    ghi = v3(std::max(ghi.x, q.x), std::max(ghi.y, q.y), std::max(ghi.z, q.z));
                    any = true;
                }
            if (!any)
                continue;
            for (int c = 0; c < 8; ++c) {
                V3 q = v3((c & 1) ? ghi.x : glo.x, (c & 2) ? ghi.y : glo.y, (c & 4) ? ghi.z : glo.z);
                if (!in.animated) {
                    grow(q);
                } else {
                    grow(xf_point(s->groups[i].m0, q));
                    grow(xf_point(s->groups[i].m1, q));
                }
            }
        }
        if (lo.x <= hi.x) {
            s->env_center = (lo + hi) * .5f;
            V3 dv = s->env_center - hi;
            s->env_radius = std::max(kRayEps, sqrtf(dot3(dv, dv)) * (1.f + kRayEps));
        }
    }
    return s;
}

void dtof_oracle_scene_destroy(dtof_oracle_scene *s) { delete s; }

int dtof_oracle_trace_samples(const dtof_oracle_scene *s, const dtof_params *p, const uint64_t *lanes, uint32_t n,
                              dtof_sample_record *out) {
    return dtof_oracle_trace_samples_pass(s, p, lanes, n, 0, out);
}

// The lanes' samples of pass `pass`: the three streams of a lane persist from pass to pass (src/render/integrator.cpp:
// 299-308), Sampler::advance() bumps the sample index and resets the dimension (src/render/sampler.cpp:52-55,94-103),
// so the earlier passes of every lane are replayed and the last one is recorded.
int dtof_oracle_trace_samples_pass(const dtof_oracle_scene *s, const dtof_params *p, const uint64_t *lanes, uint32_t n,
                                   uint32_t pass, dtof_sample_record *out) {
    dtof_pass_info pi;
    if (pass_info(s->film, *p, &pi))
        return 1;
    if (pass >= pi.n_passes)
        return 3;
    Modulation mod(*p);
    Counters st;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t idx = lanes[i];
        if (idx >= pi.wavefront_size)
            return 2;
        LaneSampler smp;
        smp.seed(*p, idx, pi.spp_per_pass);
        uint32_t pixel = (uint32_t) (idx / pi.spp_per_pass);
        uint32_t py = pixel / s->film.width, px = pixel - py * s->film.width;
        for (uint32_t k = 0; k <= pass; ++k) {
            lane_sample(*s, *p, mod, smp, px, py, out[i], st);
            smp.advance();
        }
    }
    return 0;
}

int dtof_oracle_trace_rays(const dtof_oracle_scene *s, const dtof_ray *rays, uint32_t n, int any_hit, dtof_ray_hit *out) {
    for (uint32_t i = 0; i < n; ++i) {
        const dtof_ray &r = rays[i];
        Counters st;
        dtof_ray_hit o{};
        o.instance = -1;
        const V3 ro = v3(r.o[0], r.o[1], r.o[2]), rd = v3(r.d[0], r.d[1], r.d[2]);
        if (any_hit) {
            o.hit = intersect_any(*s, ro, rd, r.tmax, r.time, st) ? 1u : 0u;
        } else {
            Hit h;
            if (intersect_closest(*s, ro, rd, r.tmax, r.time, h, st)) {
                o.hit = 1;
                o.t = h.t, o.u = h.u, o.v = h.v;
                o.prim = s->tris[h.tri].gid;
                o.instance = s->groups[h.inst].animated ? (int32_t) h.inst : -1;
            }
        }
        o.nodes_visited = (uint32_t) st.nodes, o.tris_tested = (uint32_t) st.tris;
        out[i] = o;
    }
    return 0;
}

int dtof_oracle_render(const dtof_oracle_scene *s, const dtof_params *p, int n_threads, float *rgbw_out, float *image_out) {
    dtof_pass_info pi;
    if (pass_info(s->film, *p, &pi))
        return 1;
    if (n_threads <= 0)
        n_threads = (int) std::max(1u, std::thread::hardware_concurrency());
    const uint32_t W = s->film.width, H = s->film.height;
    const size_t npx = (size_t) W * H;
    uint64_t lane_begin = p->lane_begin, lane_end = p->lane_end ? p->lane_end : pi.wavefront_size;
    if (lane_end > pi.wavefront_size || lane_begin > lane_end)
        return 2;
    if (p->shard_block && (p->shard_count == 0 || p->shard_index >= p->shard_count))
        return 2;
    Modulation mod(*p);
    FilmSplat splat(s->film);
    std::vector<double> accum(npx * 4, 0.0);
    // work unit = one pixel's lanes (keeps each lane's pass sequence on one thread)
    uint64_t pix_begin = lane_begin / pi.spp_per_pass, pix_end = (lane_end + pi.spp_per_pass - 1) / pi.spp_per_pass;
    std::atomic<uint64_t> next{ pix_begin };
    std::vector<std::vector<double>> locals((size_t) n_threads);
    std::vector<Counters> cnt((size_t) n_threads);
    auto worker = [&](int tid) {
        std::vector<double> &acc = locals[(size_t) tid];
        acc.assign(npx * 4, 0.0);
        Counters &st = cnt[(size_t) tid];
        const uint64_t chunk = 16;
        for (;;) {
            uint64_t b = next.fetch_add(chunk);
            if (b >= pix_end)
                break;
            uint64_t e = std::min(b + chunk, pix_end);
            for (uint64_t pixel = b; pixel < e; ++pixel) {
                uint32_t py = (uint32_t) (pixel / W), px = (uint32_t) (pixel - (uint64_t) py * W);
                for (uint32_t slot = 0; slot < pi.spp_per_pass; ++slot) {
                    uint64_t idx = pixel * pi.spp_per_pass + slot;
                    if (idx < lane_begin || idx >= lane_end)
                        continue;
                    if (p->shard_block && ((idx - lane_begin) / p->shard_block) % p->shard_count != p->shard_index)
                        continue;
                    LaneSampler smp;
                    smp.seed(*p, idx, pi.spp_per_pass);
                    for (uint32_t pass = 0; pass < pi.n_passes; ++pass) {
                        dtof_sample_record rec;
                        lane_sample(*s, *p, mod, smp, px, py, rec, st);
                        float rgbw[4] = { rec.rgb[0], rec.rgb[1], rec.rgb[2], 1.f };
                        splat.put(rec.sample_pos[0], rec.sample_pos[1], rgbw, acc.data());
                        smp.advance();
                    }
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t)
        th.emplace_back(worker, t);
    worker(0);
    for (auto &t : th)
        t.join();
    Counters tot;
    for (int t = 0; t < n_threads; ++t) {
        for (size_t i = 0; i < npx * 4; ++i)
            accum[i] += locals[(size_t) t][i];
        tot.rays_closest += cnt[(size_t) t].rays_closest;
        tot.rays_shadow += cnt[(size_t) t].rays_shadow;
        tot.nodes += cnt[(size_t) t].nodes;
        tot.tris += cnt[(size_t) t].tris;
        tot.inst += cnt[(size_t) t].inst;
        tot.samples += cnt[(size_t) t].samples;
    }
    s->stats = tot;
    if (rgbw_out)
        for (size_t i = 0; i < npx * 4; ++i)
            rgbw_out[i] = (float) accum[i];
    if (image_out) // HDRFilm::develop, src/films/hdrfilm.cpp:393-394
        for (size_t i = 0; i < npx; ++i) {
            double w = accum[4 * i + 3];
            if (w == 0.0)
                w = 1.0;
            for (int k = 0; k < 3; ++k)
                image_out[3 * i + k] = (float) (accum[4 * i + k] / w);
        }
    return 0;
}

void dtof_oracle_get_stats(const dtof_oracle_scene *s, dtof_stats *out) {
    out->samples = s->stats.samples;
    out->rays_closest = s->stats.rays_closest;
    out->rays_shadow = s->stats.rays_shadow;
    out->nodes_visited = s->stats.nodes;
    out->tris_tested = s->stats.tris;
    out->inst_visits = s->stats.inst;
}

} // extern "C"
