// Device-side building blocks of the B200 Doppler-ToF path tracer (sm_100a).
//
// One thread owns one wavefront lane (pixel, sample slot) for its whole life: the three PCG32
// streams, the ray, the path state and the traversal stack live in registers / L1-resident local
// memory, so nothing of the per-lane state ever round-trips through HBM ("fused wavefront").
// Stages (all __device__ functions below, fused by the render kernel in dtof_api.cu):
//   sampler      LaneSampler            <- src/samplers/correlated.cpp:38-167, src/render/sampler.cpp:85-134
//   camera       camera_ray()           <- src/sensors/perspective.cpp:238-279
//   traversal    trace_bvh/trace_flat   <- Scene::ray_intersect / ray_test (src/render/scene.cpp:125-154),
//                                          Embree motion instances (ext/embree/kernels/common/scene_instance.h:133-206)
//   interaction  compute_si()           <- src/render/mesh.cpp:633-789, src/shapes/instance.cpp:155-250,
//                                          include/mitsuba/render/interaction.h:258-268,493-516
//   shading      trace_path() (dtof_path.cuh) <- src/integrators/dopplertofpath.cpp:79-283
//   modulation   Modulation::eval()     <- dopplertofpath.cpp:60-77, include/mitsuba/render/waveform_utils.h:24-62
//   film         splat_*()              <- src/render/imageblock.cpp:206-232,418-477
//
// Floating point: compiled with -fmad=false; every fused multiply-add is an explicit fmaf() placed where the
// reference fuses (dr::fmadd / Dr.Jit's dot, cross, transform helpers), IEEE division and square root.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dtof.h"
#include "dtof_layout.h"

namespace dtof {

#define DTOF_DEV __device__ __forceinline__

constexpr int kStackSize = 96;
constexpr int kSentinel = (int) 0x80000000;
// CTA size of the fused kernel (dtof_api.cu explains the choice); device code needs it for per-thread shared-memory layouts
#ifndef DTOF_BLOCK
#define DTOF_BLOCK 1024
#endif
constexpr unsigned kFullMask = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
DTOF_DEV V3 v3(float x, float y, float z) { return V3{ x, y, z }; }
DTOF_DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DTOF_DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DTOF_DEV V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
DTOF_DEV V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
DTOF_DEV V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
// ---- shading arithmetic precision ---------------------------------------------------------------------------------
// Sampler, camera ray and the accepted-hit arithmetic of the triangle test are IEEE (they decide WHICH pixel / time /
// triangle a sample sees and are compared bit for bit with the oracle). Everything after the hit -- surface frame,
// emitter / BSDF sampling, MIS, modulation argument reduction, Russian roulette -- uses the SFU approximations below
// (<= 2 ulp per operation; `.ftz` like the reference: the non-ftz forms wrap every MUFU in four instructions of
// denormal scaling), exactly the operations the reference's own CUDA variants emit (div/rcp/sqrt/rsqrt
// `.approx.ftz`, ext/drjit/ext/drjit-core/src/cuda_eval.cpp:433-655). They are ~1e-7 relative per operation against
// the 1e-4 per-sample tolerance, and shrink the kernel by a fifth (an IEEE division is ~10 SASS instructions plus a
// slow-path call, an approximate one 1-4), which matters because the kernel is instruction-cache / issue bound
// (profiles/r01_tuning.md: +37 %). -DDTOF_EXACT_MATH restores IEEE everywhere.
#ifdef DTOF_EXACT_MATH
DTOF_DEV float fdiv(float a, float b) { return a / b; }
DTOF_DEV float frcp(float a) { return 1.f / a; }
DTOF_DEV float fsqrt(float a) { return sqrtf(a); }
DTOF_DEV float frsqrt(float a) { return 1.f / sqrtf(a); }
DTOF_DEV V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
#else
DTOF_DEV float fdiv(float a, float b) {
    float r;
    asm("div.full.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
DTOF_DEV float frcp(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
DTOF_DEV float fsqrt(float a) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
DTOF_DEV float frsqrt(float a) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
DTOF_DEV V3 operator/(V3 a, float s) {
    float r = frcp(s);
    return v3(a.x * r, a.y * r, a.z * r);
}
#endif
// rcp.approx that stays where it is written: leaving an instance recomputes the world-space reciprocal direction there
// (rare) instead of keeping nine loop-invariant values live through the whole walk
DTOF_DEV float frcp_here(float a) {
    float r;
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
DTOF_DEV V3 fma3(V3 a, float s, V3 b) { return v3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }
DTOF_DEV float dot3(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
DTOF_DEV V3 cross3(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
DTOF_DEV float rsqrt_ieee(float x) { return frsqrt(x); }   // IEEE only with DTOF_EXACT_MATH
DTOF_DEV V3 normalize3(V3 a) { return a * rsqrt_ieee(dot3(a, a)); }
DTOF_DEV float mulsign(float v, float s) {
    return __uint_as_float(__float_as_uint(v) ^ (__float_as_uint(s) & 0x80000000u));
}
DTOF_DEV float max3(V3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }

struct M34 {
    float m[12];
};
DTOF_DEV V3 xf_point(const M34 &M, V3 p) {
    return v3(fmaf(M.m[2], p.z, fmaf(M.m[1], p.y, fmaf(M.m[0], p.x, M.m[3]))),
              fmaf(M.m[6], p.z, fmaf(M.m[5], p.y, fmaf(M.m[4], p.x, M.m[7]))),
              fmaf(M.m[10], p.z, fmaf(M.m[9], p.y, fmaf(M.m[8], p.x, M.m[11]))));
}
DTOF_DEV V3 xf_vector(const M34 &M, V3 v) {
    return v3(fmaf(M.m[2], v.z, fmaf(M.m[1], v.y, M.m[0] * v.x)), fmaf(M.m[6], v.z, fmaf(M.m[5], v.y, M.m[4] * v.x)),
              fmaf(M.m[10], v.z, fmaf(M.m[9], v.y, M.m[8] * v.x)));
}
// (M^-1)^T n given Minv
DTOF_DEV V3 xf_normal(const M34 &Minv, V3 n) {
    return v3(fmaf(Minv.m[8], n.z, fmaf(Minv.m[4], n.y, Minv.m[0] * n.x)),
              fmaf(Minv.m[9], n.z, fmaf(Minv.m[5], n.y, Minv.m[1] * n.x)),
              fmaf(Minv.m[10], n.z, fmaf(Minv.m[6], n.y, Minv.m[2] * n.x)));
}
DTOF_DEV M34 inverse_m34(const M34 &M) {
    const float *a = M.m;
    float c00 = fmaf(a[5], a[10], -(a[6] * a[9])), c01 = fmaf(a[6], a[8], -(a[4] * a[10])),
          c02 = fmaf(a[4], a[9], -(a[5] * a[8]));
    float det = fmaf(a[0], c00, fmaf(a[1], c01, a[2] * c02));
    float id = frcp(det);
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
#pragma unroll
    for (int i = 0; i < 3; ++i)
        r.m[4 * i + 3] = -fmaf(r.m[4 * i + 2], a[11], fmaf(r.m[4 * i + 1], a[7], r.m[4 * i + 0] * a[3]));
    return r;
}

// ------------------------------------------------------------------------------------------------
// RNG
DTOF_DEV void tea32(uint32_t v0, uint32_t v1, uint32_t &o0, uint32_t &o1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    o0 = v0;
    o1 = v1;
}

constexpr uint64_t kPcgMult = 0x5851f42d4c957f2dull;

DTOF_DEV uint32_t pcg_output(uint64_t old) {
    uint32_t xorshifted = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return __funnelshift_r(xorshifted, xorshifted, rot);   // rotate right
}
DTOF_DEV float u32_to_float(uint32_t u) { return __uint_as_float((u >> 9) | 0x3f800000u) - 1.f; }

struct Pcg {
    uint64_t state, inc;
    DTOF_DEV uint64_t step() {
        uint64_t old = state;
        state = old * kPcgMult + inc;
        return old;
    }
    DTOF_DEV void seed(uint32_t initstate, uint32_t initseq) {
        state = 0;
        inc = ((uint64_t) initseq << 1) | 1u;
        step();
        state += initstate;
        step();
    }
    DTOF_DEV float next_f32() { return u32_to_float(pcg_output(step())); }
};

DTOF_DEV uint32_t permute_kensler(uint32_t index, uint32_t sample_count, uint32_t seed) {
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

// CorrelatedSampler (JIT branch). Only the two streams the bounce loop draws from live in registers; the time
// stream (one draw per pass) and the per-pixel permutation seed are re-derived from (idx, pass) when a pass starts.
struct LaneSampler {
    Pcg rng, rng_path;
    uint32_t draws;

    DTOF_DEV void seed(const dtof_params &p, uint32_t idx) {
        uint32_t S = p.base_seed + p.seed, a, b;
        tea32(S, idx, a, b);
        rng.seed(a, b);
        tea32(S + 2, idx / p.path_correlate_number, a, b);
        rng_path.seed(a, b);
        draws = 0;
    }
    // CORRELATED: next_1d_correlate -- both streams step, one output is formed (correlated.cpp:156-161);
    // otherwise Sampler::next_1d -- the independent stream only (correlated.cpp:78-83)
    template <bool CORRELATED = true> DTOF_DEV float next_1d(bool correlate) {
        uint64_t a = CORRELATED ? rng_path.step() : 0ull, b = rng.step();
        draws++;
        return u32_to_float(pcg_output((CORRELATED && correlate) ? a : b));
    }
    // a draw whose value is never used: only the states move
    template <bool CORRELATED = true> DTOF_DEV void skip_1d() {
        if (CORRELATED)
            rng_path.step();
        rng.step();
        draws++;
    }
    // next_1d_time of pass `pass` (src/samplers/correlated.cpp:92-153); dimension_index restarts at 0 every pass
    DTOF_DEV float next_time(const dtof_params &p, uint32_t idx, uint32_t spp_pp, uint32_t pass) {
        uint32_t strategy = p.time_sampling_method, tcn = p.time_correlate_number;
        if (strategy == DTOF_TIME_UNIFORM) {
            draws++;
            return rng.next_f32();
        }
        uint32_t si = pass * spp_pp + (spp_pp > 1 ? idx % spp_pp : 0);
        float r;
        if (strategy == DTOF_TIME_STRATIFIED) {
            r = rng.next_f32();
            draws++;
        } else {   // the time stream advances exactly once per pass in the antithetic modes
            uint32_t a, b;
            tea32(p.base_seed + p.seed + 1, idx / tcn, a, b);
            Pcg rng_time;
            rng_time.seed(a, b);
#pragma unroll 1
            for (uint32_t k = 0; k < pass; ++k)
                rng_time.step();
            r = rng_time.next_f32();
        }
        if (p.use_stratified_sampling_for_each_interval) {
            uint32_t n_stratum = p.sample_count / tcn;
            uint32_t pp;
            if (strategy == DTOF_TIME_STRATIFIED) {
                uint32_t perm_seed, unused;
                tea32(p.base_seed, spp_pp * (idx / spp_pp) + p.seed, perm_seed, unused);
                uint32_t p1 = permute_kensler(si / tcn, n_stratum, perm_seed);
                uint32_t p2 = permute_kensler(si / tcn, n_stratum, perm_seed + 1);
                pp = (si % tcn != 0) ? p1 : p2;
            } else {
                pp = si / tcn;
            }
            r = ((float) pp + r) / (float) (int) n_stratum;
        }
        uint32_t rem = si % tcn;
        if (strategy == DTOF_TIME_STRATIFIED)
            return ((float) rem + r) * (1.f / (float) tcn);
        if (strategy == DTOF_TIME_ANTITHETIC) {
            if (tcn == 2)
                return rem != 1 ? r : r + p.antithetic_shift;
            return r + (float) rem / (float) tcn;
        }
        float r2 = 1.0f - r + p.antithetic_shift;   // ANTITHETIC_MIRROR
        return rem != 1 ? r : r2;
    }
};

// ------------------------------------------------------------------------------------------------
// Math: Dr.Jit's Cephes sincos (ext/drjit/include/drjit/math.h:76-176) and fmod (array_router.h:484-486)
DTOF_DEV void dr_sincos(float x, float &s_out, float &c_out) {
    float xa = fabsf(x);
    int32_t j = (int32_t) (xa * 1.2732395447351626862f);
    j = (j + 1) & ~1;
    float y = (float) j;
    uint32_t sign_sin = ((uint32_t) j << 29) ^ __float_as_uint(x);
    uint32_t sign_cos = (uint32_t) (~(j - 2)) << 29;
    y = xa - y * 0.78515625f - y * 2.4187564849853515625e-4f - y * 3.77489497744594108e-8f;
    float z = y * y;
    if (xa == __int_as_float(0x7f800000))
        z = __int_as_float(0x7fc00000);
    float z2 = z * z;
    float s = fmaf(z2, -1.9515295891e-4f, fmaf(z, 8.3321608736e-3f, -1.6666654611e-1f)) * z;
    float c = fmaf(z2, 2.443315711809948e-5f, fmaf(z, -1.388731625493765e-3f, 4.166664568298827e-2f)) * z;
    s = fmaf(s, y, y);
    c = fmaf(c, z, fmaf(z, -0.5f, 1.f));
    bool polymask = (j & 2) == 0;
    float rs = polymask ? s : c, rc = polymask ? c : s;
    s_out = __uint_as_float(__float_as_uint(rs) ^ (sign_sin & 0x80000000u));
    c_out = __uint_as_float(__float_as_uint(rc) ^ (sign_cos & 0x80000000u));
}
DTOF_DEV float dr_cos(float x) {
    float s, c;
    dr_sincos(x, s, c);
    return c;
}
DTOF_DEV float dr_fmod(float x, float y) { return fmaf(-truncf(fdiv(x, y)), y, x); }

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvPi = 0.31830988618379067154f;

DTOF_DEV float waveform_lowpass(float t_, uint32_t type) {
    float t = dr_fmod(t_, kTwoPi);
    if (type == DTOF_WAVE_SINUSOIDAL)
        return dr_cos(t);
    float a = t * kInvPi, b = 2.f - a, c = a < b ? a : b;
    if (type == DTOF_WAVE_RECTANGULAR)
        return 2.f - 4.f * c;
    if (type == DTOF_WAVE_TRIANGULAR)
        return fdiv((4.f * c * c * c - 6.f * c * c + 1.f) * 2.f, 3.f);
    float r = 2.f - 4.f * c;
    return fminf(fmaxf(2.f * r, -2.f), 2.f);
}
DTOF_DEV float waveform_full(float t_, uint32_t type) {
    float t = dr_fmod(t_, kTwoPi);
    if (type == DTOF_WAVE_RECTANGULAR)
        return fabsf(t - kPi) > 0.5f * kPi ? 1.f : -1.f;
    if (type == DTOF_WAVE_TRIANGULAR)
        return t < kPi ? 1.f - fdiv(2.f * t, kPi) : -3.f + fdiv(2.f * t, kPi);
    return dr_cos(t);   // sinusoidal, and trapezoidal falls through (waveform_utils.h:27-32)
}

// constants derived on the host in double, then narrowed (dopplertofpath.cpp:60-65)
struct Modulation {
    float w_g, w_d, k_phi, phase, half_g1, g1, g0, w_gd;
    uint32_t type, lowpass;
    __device__ __noinline__ float eval(float ray_time, float path_length) const {
        float phi = k_phi * path_length;
        if (lowpass) {
            float t = w_d * ray_time + phase + phi;
            return half_g1 * waveform_lowpass(t, type);
        }
        float t1 = fmaf(w_g, ray_time, -phi);
        float t2 = fmaf(w_gd, ray_time, phase);
        float g_t = g1 * waveform_full(t1, type) + g0;
        float s_t = waveform_full(t2, type);
        return s_t * g_t;
    }
};

DTOF_DEV V3 square_to_cosine_hemisphere(float sx, float sy) {
    float x = fmaf(2.f, sx, -1.f), y = fmaf(2.f, sy, -1.f);
    bool is_zero = x == 0.f && y == 0.f, q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = fdiv(0.25f * kPi * rp, r);
    if (q13)
        phi = 0.5f * kPi - phi;
    if (is_zero)
        phi = 0.f;
    float s, c;
    dr_sincos(phi, s, c);
    float px = r * c, py = r * s;
    float z = fsqrt(fmaxf(1.f - fmaf(py, py, px * px), 0.f));
    return v3(px, py, z);
}
// fresnel(cos_theta_i, eta) of a dielectric interface (include/mitsuba/render/fresnel.h:35-71)
DTOF_DEV void fresnel_dielectric(float cos_theta_i, float eta, float &r, float &cos_theta_t, float &eta_it, float &eta_ti) {
    const bool outside = cos_theta_i >= 0.f;
    const float rcp_eta = frcp(eta);
    eta_it = outside ? eta : rcp_eta;
    eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t_sqr = fmaf(-fmaf(-cos_theta_i, cos_theta_i, 1.f), eta_ti * eta_ti, 1.f);
    float ci = fabsf(cos_theta_i), ct = fsqrt(fmaxf(cos_theta_t_sqr, 0.f));
    float a_s = fdiv(fmaf(-eta_it, ct, ci), fmaf(eta_it, ct, ci));
    float a_p = fdiv(fmaf(-eta_it, ci, ct), fmaf(eta_it, ci, ct));
    r = 0.5f * (a_s * a_s + a_p * a_p);
    if (eta == 1.f)
        r = 0.f;
    else if (ci == 0.f)
        r = 1.f;
    cos_theta_t = outside ? -ct : ct;   // mulsign_neg(cos_theta_t_abs, cos_theta_i)
}

// ---- MicrofacetDistribution with visible-normal sampling (include/mitsuba/render/microfacet.h:185-428) -------------
// dr::erfinv (ext/drjit/include/drjit/math.h:1527-1550, after M. Giles)
DTOF_DEV float dr_erfinv(float x) {
    float w = -logf((1.f - x) * (1.f + x));
    float w1 = w - 2.5f, w2 = fsqrt(w) - 3.f;
    float p1 = 2.81022636e-08f, p2 = -0.000200214257f;
    p1 = fmaf(p1, w1, 3.43273939e-07f), p2 = fmaf(p2, w2, 0.000100950558f);
    p1 = fmaf(p1, w1, -3.5233877e-06f), p2 = fmaf(p2, w2, 0.00134934322f);
    p1 = fmaf(p1, w1, -4.39150654e-06f), p2 = fmaf(p2, w2, -0.00367342844f);
    p1 = fmaf(p1, w1, 0.00021858087f), p2 = fmaf(p2, w2, 0.00573950773f);
    p1 = fmaf(p1, w1, -0.00125372503f), p2 = fmaf(p2, w2, -0.0076224613f);
    p1 = fmaf(p1, w1, -0.00417768164f), p2 = fmaf(p2, w2, 0.00943887047f);
    p1 = fmaf(p1, w1, 0.246640727f), p2 = fmaf(p2, w2, 1.00167406f);
    p1 = fmaf(p1, w1, 1.50140941f), p2 = fmaf(p2, w2, 2.83297682f);
    return (w < 5.f ? p1 : p2) * x;
}
struct Microfacet {
    bool ggx;
    float au, av;
    DTOF_DEV float eval(V3 m) const {   // :185-210
        float alpha_uv = au * av, cos_theta = m.z, cos_theta_2 = cos_theta * cos_theta, result;
        float ex = fdiv(m.x, au), ey = fdiv(m.y, av);
        if (!ggx) {
            result = fdiv(expf(-fdiv(ex * ex + ey * ey, cos_theta_2)), kPi * alpha_uv * (cos_theta_2 * cos_theta_2));
        } else {
            float q = ex * ex + ey * ey + m.z * m.z;
            result = frcp(kPi * alpha_uv * (q * q));
        }
        return result * cos_theta > 1e-20f ? result : 0.f;
    }
    DTOF_DEV float smith_g1(V3 v, V3 m) const {   // :341-365
        float ax = au * v.x, ay = av * v.y;
        float xy_alpha_2 = ax * ax + ay * ay, tan_theta_alpha_2 = fdiv(xy_alpha_2, v.z * v.z), result;
        if (!ggx) {
            float a = frsqrt(tan_theta_alpha_2), a_sqr = a * a;
            result = a >= 1.6f ? 1.f : fdiv(3.535f * a + 2.181f * a_sqr, 1.f + 2.276f * a + 2.577f * a_sqr);
        } else {
            result = fdiv(2.f, 1.f + fsqrt(1.f + tan_theta_alpha_2));
        }
        if (xy_alpha_2 == 0.f)
            result = 1.f;
        if (dot3(v, m) * v.z <= 0.f)
            result = 0.f;
        return result;
    }
    DTOF_DEV void sample_visible_11(float cos_theta_i, float sx, float sy, float &slope_x, float &slope_y) const {   // :368-418
        if (!ggx) {
            const float inv_sqrt_pi = 0.56418958354775628695f;
            float tan_theta_i = fdiv(fsqrt(fmaxf(fmaf(-cos_theta_i, cos_theta_i, 1.f), 0.f)), cos_theta_i);
            float cot_theta_i = frcp(tan_theta_i);
            float maxval = erff(cot_theta_i);
            sx = fmaxf(fminf(sx, 1.f - 1e-6f), 1e-6f);
            sy = fmaxf(fminf(sy, 1.f - 1e-6f), 1e-6f);
            float x = maxval - (maxval + 1.f) * erff(fsqrt(-logf(sx)));
            sx *= 1.f + maxval + inv_sqrt_pi * tan_theta_i * expf(-(cot_theta_i * cot_theta_i));
#pragma unroll 1
            for (int i = 0; i < 3; ++i) {   // three Newton iterations
                float slope = dr_erfinv(x);
                float value = 1.f + x + inv_sqrt_pi * tan_theta_i * expf(-(slope * slope)) - sx;
                float derivative = 1.f - slope * tan_theta_i;
                x -= fdiv(value, derivative);
            }
            slope_x = dr_erfinv(x);
            slope_y = dr_erfinv(fmaf(2.f, sy, -1.f));
        } else {
            // warp::square_to_uniform_disk_concentric (warp.h:54-89), as in square_to_cosine_hemisphere below
            float x0 = fmaf(2.f, sx, -1.f), y0 = fmaf(2.f, sy, -1.f);
            bool is_zero = x0 == 0.f && y0 == 0.f, q13 = fabsf(x0) < fabsf(y0);
            float r = q13 ? y0 : x0, rp = q13 ? x0 : y0;
            float phi = fdiv(0.25f * kPi * rp, r);
            if (q13)
                phi = 0.5f * kPi - phi;
            if (is_zero)
                phi = 0.f;
            float s, c;
            dr_sincos(phi, s, c);
            float px = r * c, py = r * s;
            float t = 0.5f * (1.f + cos_theta_i);
            float a = fsqrt(fmaxf(1.f - px * px, 0.f));
            py = fmaf(py, t, fmaf(-a, t, a));   // dr::lerp(a, py, t)
            float z = fsqrt(fmaxf(1.f - (px * px + py * py), 0.f));
            float sin_theta_i = fsqrt(fmaxf(1.f - cos_theta_i * cos_theta_i, 0.f));
            float norm = frcp(fmaf(sin_theta_i, py, cos_theta_i * z));
            slope_x = fmaf(cos_theta_i, py, -(sin_theta_i * z)) * norm;
            slope_y = px * norm;
        }
    }
    DTOF_DEV V3 sample(V3 wi, float sx, float sy, float &pdf) const {   // :296-326
        V3 wi_p = normalize3(v3(au * wi.x, av * wi.y, wi.z));
        float sin_theta_2 = fmaf(wi_p.x, wi_p.x, wi_p.y * wi_p.y), inv = frsqrt(sin_theta_2);
        float sin_phi = fminf(fmaxf(wi_p.y * inv, -1.f), 1.f), cos_phi = fminf(fmaxf(wi_p.x * inv, -1.f), 1.f);
        if (fabsf(sin_theta_2) <= 4.f * 5.9604644775390625e-08f)
            sin_phi = 0.f, cos_phi = 1.f;
        float slx, sly;
        sample_visible_11(wi_p.z, sx, sy, slx, sly);
        float rx = fmaf(cos_phi, slx, -(sin_phi * sly)) * au, ry = fmaf(sin_phi, slx, cos_phi * sly) * av;
        V3 m = normalize3(v3(-rx, -ry, 1.f));
        pdf = fdiv(eval(m) * smith_g1(wi, m) * fabsf(dot3(wi, m)), wi.z);
        return m;
    }
};

DTOF_DEV float fresnel_r(float cos_theta_i, float eta) {
    float r, ct, eit, eti;
    fresnel_dielectric(cos_theta_i, eta, r, ct, eit, eti);
    return r;
}

// fresnel_conductor (include/mitsuba/render/fresnel.h:93-117), one colour channel
DTOF_DEV float fresnel_conductor(float cos_theta_i, float eta_r, float eta_i) {
    float cos2 = cos_theta_i * cos_theta_i, sin2 = 1.f - cos2, sin4 = sin2 * sin2;
    float temp_1 = eta_r * eta_r - eta_i * eta_i - sin2;
    float a_2_pb_2 = fsqrt(fmaxf(temp_1 * temp_1 + 4.f * eta_i * eta_i * eta_r * eta_r, 0.f));
    float a = fsqrt(fmaxf(.5f * (a_2_pb_2 + temp_1), 0.f));
    float term_1 = a_2_pb_2 + cos2, term_2 = 2.f * cos_theta_i * a;
    float r_s = fdiv(term_1 - term_2, term_1 + term_2);
    float term_3 = a_2_pb_2 * cos2 + sin4, term_4 = term_2 * sin2;
    float r_p = r_s * fdiv(term_3 - term_4, term_3 + term_4);
    return .5f * (r_s + r_p);
}

// warp::square_to_uniform_sphere (include/mitsuba/core/warp.h:250-255)
DTOF_DEV V3 square_to_uniform_sphere(float sx, float sy) {
    float z = fmaf(-2.f, sy, 1.f);
    float r = fsqrt(fmaxf(fmaf(-z, z, 1.f), 0.f));   // circ(z) = safe_sqrt(fnmadd(z, z, 1))
    float s, c;
    dr_sincos(2.f * kPi * sx, s, c);
    return v3(r * c, r * s, z);
}
constexpr float kInvFourPi = 0.07957747154594766788f;
DTOF_DEV void coordinate_system(V3 n, V3 &s, V3 &t) {
    float sign = copysignf(1.f, n.z);
    float a = -frcp(sign + n.z), b = n.x * n.y * a;
    s = v3(mulsign(n.x * n.x * a, n.z) + 1.f, mulsign(b, n.z), mulsign(-n.x, n.z));
    t = v3(b, fmaf(n.y, n.y * a, sign), -n.y);
}
DTOF_DEV void square_to_uniform_triangle(float sx, float sy, float &bx, float &by) {
    float t = fsqrt(fmaxf(1.f - sx, 0.f));
    bx = 1.f - t;
    by = t * sy;
}

// ------------------------------------------------------------------------------------------------
// Scene access
struct DeviceScene {
    const float4 *nodes;    // BvhNode  = 4 x float4
    const float4 *tris;     // TriIsect = 3 x float4
    const float4 *shade;    // TriShade = 7 x float4
    const float4 *insts;    // InstRec  = 7 x float4
    const MeshRec *meshes;
    const BsdfRec *bsdfs;
    const EmitterRec *emitters;
    const SpotRec *spots;    // parallel to `emitters` when the scene has a spot light, else NULL
    const float *area_cdf, *area_pmf;
    int32_t root;
    uint32_t n_emitters, n_insts, n_nodes, n_tris, has_geometry;
    // constant environment emitter (src/emitters/constant.cpp): index into `emitters` or -1; radiance; bounding sphere
    uint32_t extended;     // the scene uses an environment emitter or a conductor (kernel template parameter ENV)
    int32_t env_emitter;
    float env_r, env_g, env_b;
    float env_cx, env_cy, env_cz, env_radius;
};

struct Counters {
    unsigned long long rays_closest, rays_shadow, nodes, tris, inst, samples;
};

struct Hit {
    float t, u, v;
    uint32_t gid;
    int32_t inst;
};

// ---- box test of one BVH node (both children) ----------------------------------------------------------------------
// A child box is stored as centre m and half extent h per axis (dtof_layout.h). Per axis the ray's parameter interval is
//     c = (m - o) * idir,   [c - h |idir|, c + h |idir|]      (sub, mul, two fmas)
// -- no min / max per axis: the lower end is the lower end whatever the sign of the direction. The fused kernel's walk
// is bound by the half-rate ALU pipe that executes FMNMX / FSETP / SEL (ncu, round 1: ALU 54 %, FMA 29 % of peak over the
// whole kernel): the lo / hi form costs 20 min / max per node, this form 8.
// Conservative for any origin: every rounding error is RELATIVE to the value it affects ((m - o) is one rounding, the
// product another, each fma one more; idir is 1 ulp off per axis), covered by widening the far side by 3e-6 and by the
// builder, which rounds the half extents up and pads the boxes by 1e-5. (A form with one fma per plane, m * idir - o *
// idir, carries an ABSOLUTE error of ulp(o * idir); bounding it per ray loosens every axis by the error of the axis the
// ray is most parallel to -- a ray with a zero direction component then accepts every box of a 1.3 M-node BVH.)
// An infinite idir (zero direction component) gives inf / NaN on that axis; fminf / fmaxf drop NaN operands, so the axis is
// at worst not used for culling -- unless the reciprocal direction is clamped (RaySlab::set<ROBUST> below).
constexpr float kSlabWiden = 1.000003f;
// A direction component that is exactly zero (or denormal: rcp.approx.ftz) has 1 / d = inf, and (m - o) * inf -/+ h * inf
// is NaN on one side: the axis would drop out of the test and an axis-parallel ray would walk every node whose OTHER two
// slabs it crosses. Clamping |1 / d| to 1e18 keeps the arithmetic finite: inside the slab (|m - o| <= h) the interval is
// [-huge, +huge], outside it lies entirely beyond any finite t, which is exactly what the limit d -> 0 means. For a
// component with |d| < 1e-18 the clamp rescales that axis' interval, whose ends are beyond 1e18 * (distance to the slab
// planes) either way -- farther than anything the other axes or `best` can produce unless the ray starts within 1e-18 of
// the plane, far below the boxes' 1e-5 padding.
// The clamp is compiled in where rays come from a CALLER (`ROBUST`: dtof_trace_rays, where axis-parallel probe rays are the
// common case); the render kernels' rays have continuous random directions, an exactly zero component has probability
// ~2^-23 per ray and only costs that ray a longer walk, so they skip the three extra instructions per ray set-up (measured:
// 1 % of the fused kernel).
constexpr float kSlabMaxInvDir = 1e18f;
struct RaySlab {
    V3 o, id, aid;
    template <bool ROBUST = false> DTOF_DEV void set(V3 ro, V3 rid) {
        o = ro;
        if (ROBUST) {
            aid = v3(fminf(fabsf(rid.x), kSlabMaxInvDir), fminf(fabsf(rid.y), kSlabMaxInvDir), fminf(fabsf(rid.z), kSlabMaxInvDir));
            id = v3(copysignf(aid.x, rid.x), copysignf(aid.y, rid.y), copysignf(aid.z, rid.z));
        } else {
            id = rid;
            aid = v3(fabsf(rid.x), fabsf(rid.y), fabsf(rid.z));
        }
    }
};
// n0 = child 0 {mx, hx, my, hy}, n1 = child 1 {mx, hx, my, hy}, n2 = {c0 mz, c0 hz, c1 mz, c1 hz}
DTOF_DEV void node_test(const float4 n0, const float4 n1, const float4 n2, const RaySlab &R, float best, bool &h0, bool &h1,
                        float &t0n, float &t1n) {
    const float c0x = (n0.x - R.o.x) * R.id.x, c0y = (n0.z - R.o.y) * R.id.y, c0z = (n2.x - R.o.z) * R.id.z;
    const float c1x = (n1.x - R.o.x) * R.id.x, c1y = (n1.z - R.o.y) * R.id.y, c1z = (n2.z - R.o.z) * R.id.z;
    t0n = fmaxf(fmaxf(fmaf(-n0.y, R.aid.x, c0x), fmaf(-n0.w, R.aid.y, c0y)), fmaxf(fmaf(-n2.y, R.aid.z, c0z), 0.f));
    t1n = fmaxf(fmaxf(fmaf(-n1.y, R.aid.x, c1x), fmaf(-n1.w, R.aid.y, c1y)), fmaxf(fmaf(-n2.w, R.aid.z, c1z), 0.f));
    const float t0f = fminf(fminf(fmaf(n0.y, R.aid.x, c0x), fmaf(n0.w, R.aid.y, c0y)), fmaf(n2.y, R.aid.z, c0z));
    const float t1f = fminf(fminf(fmaf(n1.y, R.aid.x, c1x), fmaf(n1.w, R.aid.y, c1y)), fmaf(n2.w, R.aid.z, c1z));
    // the widening applies to `best` as well: for a far origin one ulp of t exceeds the boxes' padding, and a box whose
    // entry distance ties with the best hit so far must still be visited (the triangle test itself compares exactly)
    h0 = t0n <= fminf(t0f, best) * kSlabWiden;
    h1 = t1n <= fminf(t1f, best) * kSlabWiden;
}

// Moeller-Trumbore (include/mitsuba/render/mesh.h:342-365) against one stored triangle (p0 | e1 | e2), split into a
// conservative pre-test and the exact reference arithmetic. The pre-test works on the un-normalised numerators
// (U = tvec.pvec, V = d.qvec, D = det), never rejects a triangle the exact test accepts (margins >> the 2 ulp that
// u = U * (1/D) can move), and lets most tests finish without the IEEE reciprocal. Accepted hits are bit-identical to
// the plain formula.
DTOF_DEV bool tri_test(const float4 a, const float4 b, const float4 c, V3 ro, V3 rd, float best, float &t_out,
                       float &u_out, float &v_out) {
    V3 e1 = v3(b.x, b.y, b.z), e2 = v3(c.x, c.y, c.z);
    V3 pvec = cross3(rd, e2);
    float det = dot3(e1, pvec);
    V3 tvec = ro - v3(a.x, a.y, a.z);
    float U = dot3(tvec, pvec);
    float Da = fabsf(det);
    float Us = mulsign(U, det);
    if (!(Us >= -1e-20f * Da && Us <= Da * 1.000001f))
        return false;
    V3 qvec = cross3(tvec, e1);
    float V = dot3(rd, qvec);
    float Vs = mulsign(V, det);
    if (!(Vs >= -1e-20f * Da && Us + Vs <= Da * 1.000002f))
        return false;
    float inv_det = 1.f / det;
    float u = U * inv_det;
    float v = V * inv_det;
    float t = dot3(e2, qvec) * inv_det;
    t_out = t;
    u_out = u;
    v_out = v;
    return u >= 0.f && u <= 1.f && v >= 0.f && u + v <= 1.f && t >= 0.f && t <= best;
}

DTOF_DEV void load_inst_matrix(const float4 *ip, float time, bool clamp, M34 &M) {
    // AnimatedTransform::eval (clamped, transform.h:451-456) or Embree's unclamped fraction (default.h:225-231)
    float4 q6 = ip[6];
    float f = fdiv(time - q6.x, q6.y - q6.x);
    if (clamp)
        f = fminf(fmaxf(f, 0.f), 1.f);
    float s = 1.f - f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float4 a = ip[k], b = ip[3 + k];
        M.m[4 * k + 0] = fmaf(b.x, f, a.x * s);
        M.m[4 * k + 1] = fmaf(b.y, f, a.y * s);
        M.m[4 * k + 2] = fmaf(b.z, f, a.z * s);
        M.m[4 * k + 3] = fmaf(b.w, f, a.w * s);
    }
}

// Enter an animated instance (Embree: world->local = rcp(lerp(M0, M1, f)) with the UNCLAMPED fraction f,
// ext/embree/kernels/common/scene_instance.h:133-138,186-206, default.h:225-231)
DTOF_DEV void enter_instance(const float4 *ip, V3 o, V3 d, float time, V3 &ro, V3 &rd) {
    M34 M;
    load_inst_matrix(ip, time, false, M);
    M34 inv = inverse_m34(M);
    ro = xf_point(inv, o);
    rd = xf_vector(inv, d);
}

constexpr int kDone = 0x7fffffff;

// Closest-hit (ANY=false) or any-hit (ANY=true) traversal of the BVH with per-ray time.
// Static geometry is single-level (the static group's BLAS is spliced into the TLAS); an animated instance is an
// "instance leaf": the ray is moved into the instance's space, a sentinel marks the way back.
// Loop shape after Aila & Laine: a lane stays in the inner-node loop (popping included) until it holds a leaf.
// `N`, `T`, `I` are the node / triangle / instance arrays (global memory, or their shared-memory copies).
// The box test is node_test() above.
template <bool STATS, bool ROBUST = false>
DTOF_DEV bool trace_bvh(const float4 *__restrict__ N, const float4 *__restrict__ T, const float4 *__restrict__ I,
                        int32_t root, const bool ANY, V3 o, V3 d, float tmax, float time, Hit &hit, Counters &st) {
    int stack[kStackSize];
    int sp = 0;
    int node = root;
    int cur_inst = -1;
    V3 ro = o, rd = d;                                       // ray in the current (world / instance) space
    RaySlab R;
    R.template set<ROBUST>(o, v3(frcp(d.x), frcp(d.y), frcp(d.z)));
    float best = tmax;
    bool found = false;
    if (STATS) {
        if (ANY) st.rays_shadow++; else st.rays_closest++;
    }
#define DTOF_POP()                                                                                         \
    do {                                                                                                   \
        if (sp == 0) {                                                                                     \
            node = kDone;                                                                                  \
        } else {                                                                                           \
            node = stack[--sp];                                                                            \
            if (node == kSentinel) { /* leave the instance: back to the world-space ray */                 \
                ro = o, rd = d, cur_inst = -1;                                                             \
                R.template set<ROBUST>(o, v3(frcp_here(d.x), frcp_here(d.y), frcp_here(d.z)));                              \
                node = sp ? stack[--sp] : kDone;                                                           \
            }                                                                                              \
        }                                                                                                  \
    } while (0)

    while (node != kDone) {
        // ---- inner nodes (node in [0, kDone))
        while ((unsigned) node < (unsigned) kDone) {
            const float4 *np = N + 4 * (size_t) node;
            const float4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
            if (STATS) st.nodes++;
            bool h0, h1;
            float t0n, t1n;
            node_test(n0, n1, n2, R, best, h0, h1, t0n, t1n);
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swap = t1n < t0n;
                stack[sp++] = swap ? c0 : c1;
                node = swap ? c1 : c0;
            } else if (h0 || h1) {
                node = h0 ? c0 : c1;
            } else {
                DTOF_POP();
            }
        }
        if (node == kDone)
            break;
        // ---- leaf
        uint32_t code = (uint32_t) ~node;
        uint32_t count = code & 15u;
        if (count == 0) {
            cur_inst = (int) (code >> 4);
            const float4 *ip = I + 8 * (size_t) cur_inst;
            if (STATS) st.inst++;
            enter_instance(ip, o, d, time, ro, rd);
            R.template set<ROBUST>(ro, v3(frcp(rd.x), frcp(rd.y), frcp(rd.z)));
            stack[sp++] = kSentinel;
            node = __float_as_int(ip[6].z);
            continue;
        }
        uint32_t first = code >> 4;
        for (uint32_t i = 0; i < count; ++i) {
            const float4 *tp = T + 3 * (size_t) (first + i);
            float4 a = tp[0], b = tp[1], c = tp[2];
            if (STATS) st.tris++;
            float t, u, v;
            if (tri_test(a, b, c, ro, rd, best, t, u, v)) {
                if (ANY)
                    return true;
                uint32_t gid = __float_as_uint(a.w);
                if (t < best || !found || gid < hit.gid) {
                    best = t;
                    hit.t = t;
                    hit.u = u;
                    hit.v = v;
                    hit.gid = gid;
                    hit.inst = cur_inst;
                    found = true;
                }
            }
        }
        DTOF_POP();
    }
#undef DTOF_POP
    return found;
}

// ------------------------------------------------------------------------------------------------
// Traversal of a scene that is staged in SHARED memory (BVH_SMEM mode). Same walk as trace_bvh (same child order,
// triangle test and tie-break, so hits are identical), rebuilt around what the issue-bound fused kernel pays for:
//  * 32-bit shared-window addresses + ld.shared: no generic-to-shared base is re-derived inside the loops;
//  * the traversal stack lives in shared memory, one 128-byte row per level and warp (slot = lane): a push / pop is
//    conflict-free whatever the lanes' depths are (an L1-resident local-memory stack serialises over the distinct
//    levels of a warp) and no stack traffic reaches L2 / DRAM;
//  * the box test is node_test() above (centre / half-extent boxes, no per-axis min / max);
//  * the world-space reciprocal direction is recomputed when a lane leaves an instance instead of being kept live.
struct SmemScene {
    uint32_t N, T, I;   // byte addresses (shared window) of nodes, leaf-order triangles, instance records
    uint32_t stack;     // this lane's slot of stack level 0; level l is at stack + 128 * l
};
DTOF_DEV float4 lds128(uint32_t a) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
DTOF_DEV int lds_stack(uint32_t a) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
DTOF_DEV void sts_stack(uint32_t a, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// a value the compiler must keep in a register as it is (it would otherwise re-derive the shared-window base from
// %cluster_ctaid inside the traversal loop)
DTOF_DEV uint32_t opaque_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

template <bool STATS, bool ROBUST = false>
DTOF_DEV bool trace_bvh_smem(const SmemScene M, int32_t root, const bool ANY, V3 o, V3 d, float tmax, float time, Hit &hit,
                             Counters &st) {
#ifdef DTOF_LOCAL_STACK   // A/B builds only: the stack in (L1-resident) local memory, the shared-memory carve-out stays small
    int lstack[40];
    uint32_t sp = 0;
#define DTOF_STK_EMPTY() (sp == 0u)
#define DTOF_STK_PUSH(v) (lstack[sp++] = (v))
#define DTOF_STK_POP() (lstack[--sp])
#else
    uint32_t sp = M.stack;
#define DTOF_STK_EMPTY() (sp == M.stack)
#define DTOF_STK_PUSH(v) (sts_stack(sp, (v)), sp += 128u)
#define DTOF_STK_POP() (sp -= 128u, lds_stack(sp))
#endif
    int node = root;
    int cur_inst = -1;
    V3 ro = o, rd = d;
    RaySlab R;
    R.template set<ROBUST>(o, v3(frcp(d.x), frcp(d.y), frcp(d.z)));
    float best = tmax;
    bool found = false;
    if (STATS) {
        if (ANY) st.rays_shadow++; else st.rays_closest++;
    }
#define DTOF_SPOP()                                                                                        \
    do {                                                                                                   \
        if (DTOF_STK_EMPTY()) {                                                                            \
            node = kDone;                                                                                  \
        } else {                                                                                           \
            node = DTOF_STK_POP();                                                                         \
            if (node == kSentinel) { /* leave the instance: back to the world-space ray */                 \
                ro = o, rd = d, cur_inst = -1;                                                             \
                R.template set<ROBUST>(o, v3(frcp_here(d.x), frcp_here(d.y), frcp_here(d.z)));                              \
                node = DTOF_STK_EMPTY() ? kDone : DTOF_STK_POP();                                          \
            }                                                                                              \
        }                                                                                                  \
    } while (0)

    while (node != kDone) {
        // ---- inner nodes
        while ((unsigned) node < (unsigned) kDone) {
            const uint32_t na = M.N + ((uint32_t) node << 6);
            const float4 n0 = lds128(na), n1 = lds128(na + 16), n2 = lds128(na + 32), n3 = lds128(na + 48);
            if (STATS) st.nodes++;
            bool h0, h1;
            float t0n, t1n;
            node_test(n0, n1, n2, R, best, h0, h1, t0n, t1n);
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swap = t1n < t0n;
                DTOF_STK_PUSH(swap ? c0 : c1);
                node = swap ? c1 : c0;
            } else if (h0 || h1) {
                node = h0 ? c0 : c1;
            } else {
                DTOF_SPOP();
            }
        }
        if (node == kDone)
            break;
        // ---- leaf
        const uint32_t code = (uint32_t) ~node;
        const uint32_t count = code & 15u;
        if (count == 0) {   // animated instance (Embree semantics, see enter_instance)
            cur_inst = (int) (code >> 4);
            const uint32_t ia = M.I + ((uint32_t) cur_inst << 7);
            if (STATS) st.inst++;
            const float4 q6 = lds128(ia + 96);
            const float f = fdiv(time - q6.x, q6.y - q6.x), s = 1.f - f;
            M34 Mx;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 a = lds128(ia + 16 * k), b = lds128(ia + 48 + 16 * k);
                Mx.m[4 * k + 0] = fmaf(b.x, f, a.x * s);
                Mx.m[4 * k + 1] = fmaf(b.y, f, a.y * s);
                Mx.m[4 * k + 2] = fmaf(b.z, f, a.z * s);
                Mx.m[4 * k + 3] = fmaf(b.w, f, a.w * s);
            }
            const M34 inv = inverse_m34(Mx);
            ro = xf_point(inv, o);
            rd = xf_vector(inv, d);
            R.template set<ROBUST>(ro, v3(frcp(rd.x), frcp(rd.y), frcp(rd.z)));
            DTOF_STK_PUSH(kSentinel);
            node = __float_as_int(q6.z);
            continue;
        }
        uint32_t ta = M.T + 48u * (code >> 4);
        for (uint32_t i = 0; i < count; ++i, ta += 48u) {
            const float4 a = lds128(ta), b = lds128(ta + 16), c = lds128(ta + 32);
            if (STATS) st.tris++;
            float t, u, v;
            if (tri_test(a, b, c, ro, rd, best, t, u, v)) {
                if (ANY)
                    return true;
                const uint32_t gid = __float_as_uint(a.w);
                if (t < best || !found || gid < hit.gid) {
                    best = t;
                    hit.t = t;
                    hit.u = u;
                    hit.v = v;
                    hit.gid = gid;
                    hit.inst = cur_inst;
                    found = true;
                }
            }
        }
        DTOF_SPOP();
    }
#undef DTOF_SPOP
#undef DTOF_STK_EMPTY
#undef DTOF_STK_PUSH
#undef DTOF_STK_POP
    return found;
}

// Flat, warp-coherent traversal: every lane walks ALL triangles of every instance in scene order (uniform control
// flow, shared-memory reads are broadcasts, no stack). An animated instance is skipped when no lane of the warp
// touches its padded world bounds. Kept as a measured alternative (DTOF_MODE=2): since the kernel runs at 4 CTAs/SM
// with leaves of <= 2 triangles the per-lane BVH walk is faster even for 22-36 triangle scenes
// (profiles/r01_tuning.md), so it is never selected automatically.
// `TF` = triangles in scene (gid) order, `B` = instance world boxes. Must be called by all 32 lanes of the warp;
// `lane_active` masks lanes without a ray.
template <bool STATS>
DTOF_DEV bool trace_flat(const float4 *__restrict__ TF, const float4 *__restrict__ I, const float4 *__restrict__ B,
                         uint32_t n_insts, const bool ANY, V3 o, V3 d, float tmax, float time, bool lane_active, Hit &hit,
                         Counters &st) {
    float best = tmax;
    bool found = false;
    const V3 wid = v3(frcp(d.x), frcp(d.y), frcp(d.z));   // boxes are padded by 1e-5 relative: 1 ulp is immaterial
    if (STATS && lane_active) {
        if (ANY) st.rays_shadow++; else st.rays_closest++;
    }
    for (uint32_t g = 0; g < n_insts; ++g) {
        const float4 *ip = I + 8 * (size_t) g;
        const float4 q6 = ip[6], q7 = ip[7];
        const uint32_t first = __float_as_uint(q7.x), n = __float_as_uint(q7.y);
        const bool animated = __float_as_uint(q6.w) != 0u;
        V3 ro = o, rd = d;
        bool want = lane_active && !(ANY && found);
        if (animated) {
            // world-bounds cull (slab test as in the BVH walk), decided per warp
            float4 lo = B[2 * g], hi = B[2 * g + 1];
            float lx = (lo.x - o.x) * wid.x, hx = (hi.x - o.x) * wid.x;
            float ly = (lo.y - o.y) * wid.y, hy = (hi.y - o.y) * wid.y;
            float lz = (lo.z - o.z) * wid.z, hz = (hi.z - o.z) * wid.z;
            float tn = fmaxf(fmaxf(fminf(lx, hx), fminf(ly, hy)), fmaxf(fminf(lz, hz), 0.f));
            float tf = fminf(fminf(fmaxf(lx, hx), fmaxf(ly, hy)), fmaxf(lz, hz)) * 1.0000005f;
            want = want && tn <= fminf(tf, best);
            if (!__any_sync(kFullMask, want))
                continue;
            if (STATS && want) st.inst++;
            enter_instance(ip, o, d, time, ro, rd);
        } else if (!__any_sync(kFullMask, want)) {
            continue;
        }
        const float4 *tp = TF + 3 * (size_t) first;
#pragma unroll 2
        for (uint32_t i = 0; i < n; ++i, tp += 3) {
            float4 a = tp[0], b = tp[1], c = tp[2];
            if (STATS && want) st.tris++;
            float t, u, v;
            if (want && tri_test(a, b, c, ro, rd, best, t, u, v)) {
                if (ANY) {
                    found = true;
                    want = false;
                } else if (t < best || !found) {   // scene order == ascending gid: ties keep the first
                    best = t;
                    hit.t = t;
                    hit.u = u;
                    hit.v = v;
                    hit.gid = __float_as_uint(a.w);
                    hit.inst = animated ? (int) g : -1;
                    found = true;
                }
            }
        }
        if (ANY && __all_sync(kFullMask, found || !lane_active))
            break;
    }
    return found;
}

// ------------------------------------------------------------------------------------------------
struct SI {
    V3 p, n, sh_n, sh_s, sh_t, wi;
    uint32_t mesh;
};

DTOF_DEV void compute_si(const DeviceScene &S, const float4 *__restrict__ I, const Hit &h, V3 ray_d, float ray_time,
                         SI &si) {
    const float4 *sp = S.shade + 7 * (size_t) h.gid;
    float4 s0 = __ldg(sp + 0), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2);
    V3 p0 = v3(s0.x, s0.y, s0.z), p1 = v3(s1.x, s1.y, s1.z), p2 = v3(s2.x, s2.y, s2.z);
    uint32_t flags = __float_as_uint(s1.w);
    si.mesh = __float_as_uint(s0.w);
    float b1 = h.u, b2 = h.v, b0 = 1.f - b1 - b2;
    V3 dp0 = p1 - p0, dp1 = p2 - p0;
    si.p = fma3(p0, b0, fma3(p1, b1, p2 * b2));
    si.n = normalize3(cross3(dp0, dp1));
    V3 dp_du, dp_dv;
    coordinate_system(si.n, dp_du, dp_dv);
    float4 s3, s4, s5, s6;
    if (flags & (TRI_HAS_NORMALS | TRI_HAS_UV)) {
        s3 = __ldg(sp + 3), s4 = __ldg(sp + 4), s5 = __ldg(sp + 5), s6 = __ldg(sp + 6);
    }
    if (flags & TRI_HAS_UV) {
        float u0x = s5.y, u0y = s5.z, u1x = s5.w, u1y = s6.x, u2x = s6.y, u2y = s6.z;
        float d0x = u1x - u0x, d0y = u1y - u0y, d1x = u2x - u0x, d1y = u2y - u0y;
        float det = fmaf(d0x, d1y, -(d0y * d1x));
        if (det != 0.f) {
            float inv_det = frcp(det);
            dp_du = v3(fmaf(d1y, dp0.x, -(d0y * dp1.x)), fmaf(d1y, dp0.y, -(d0y * dp1.y)),
                       fmaf(d1y, dp0.z, -(d0y * dp1.z))) *
                    inv_det;
        }
    }
    if (flags & TRI_HAS_NORMALS) {
        V3 n0 = v3(s3.x, s3.y, s3.z), n1 = v3(s3.w, s4.x, s4.y), n2 = v3(s4.z, s4.w, s5.x);
        V3 n = fma3(n2, b2, fma3(n1, b1, n0 * b0));
        si.sh_n = n * rsqrt_ieee(dot3(n, n));
    } else {
        si.sh_n = si.n;
    }
    if (flags & TRI_FLIP) {
        si.n = -si.n;
        si.sh_n = -si.sh_n;
    }
    if (h.inst >= 0) {   // only animated instances are recorded on a hit
        const float4 *ip = I + 8 * (size_t) h.inst;
        M34 to_world;
        load_inst_matrix(ip, ray_time, true, to_world);
        M34 to_object = inverse_m34(to_world);
        si.p = xf_point(to_world, si.p);
        si.n = normalize3(xf_normal(to_object, si.n));
        si.sh_n = normalize3(xf_normal(to_object, si.sh_n));
        dp_du = xf_vector(to_world, dp_du);
    }
    si.sh_s = normalize3(fma3(si.sh_n, -dot3(si.sh_n, dp_du), dp_du));
    if (dp_du.x == 0.f && dp_du.y == 0.f && dp_du.z == 0.f) {
        V3 tmp;
        coordinate_system(si.sh_n, si.sh_s, tmp);
    }
    si.sh_t = cross3(si.sh_n, si.sh_s);
    V3 md = -ray_d;
    si.wi = v3(dot3(md, si.sh_s), dot3(md, si.sh_t), dot3(md, si.sh_n));
}

constexpr float kRayEps = 1500.f * 5.9604644775390625e-08f;
constexpr float kShadowEps = kRayEps * 10.f;

DTOF_DEV V3 offset_p(V3 p, V3 n, V3 d) {
    float mag = (1.f + fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z))) * kRayEps;
    mag = mulsign(mag, dot3(n, d));
    return fma3(n, mag, p);
}
DTOF_DEV float mis_weight(float a, float b) {
    a *= a;
    b *= b;
    float w = fdiv(a, a + b);
    return isfinite(w) ? w : 0.f;
}

DTOF_DEV void sample_position(const DeviceScene &S, const MeshRec &m, float sx, float sy, V3 &p, V3 &n, float &pdf) {
    if (m.kind == DTOF_SHAPE_RECTANGLE) {
        M34 tw;
#pragma unroll
        for (int i = 0; i < 12; ++i)
            tw.m[i] = m.rect_to_world[i];
        p = xf_point(tw, v3(sx * 2.f - 1.f, sy * 2.f - 1.f, 0.f));
        n = v3(m.rect_n[0], m.rect_n[1], m.rect_n[2]);
        pdf = m.inv_area;
        return;
    }
    const float *cdf = S.area_cdf + m.cdf_offset, *pmf = S.area_pmf + m.cdf_offset;
    float value = sy * m.area_sum;
    uint32_t lo = m.valid_lo, hi = m.valid_hi;
    while (lo < hi) {
        uint32_t mid = (lo + hi) / 2;
        if (cdf[mid] < value)
            lo = mid + 1;
        else
            hi = mid;
    }
    uint32_t face = lo;
    float pmf_n = pmf[face] * m.inv_area;
    float cdf_n = face > 0 ? cdf[face - 1] * m.inv_area : 0.f;
    sy = fdiv(sy - cdf_n, pmf_n);
    const float4 *sp = S.shade + 7 * (size_t) (m.first_gid + face);
    float4 s0 = __ldg(sp + 0), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2);
    V3 p0 = v3(s0.x, s0.y, s0.z), p1 = v3(s1.x, s1.y, s1.z), p2 = v3(s2.x, s2.y, s2.z);
    uint32_t flags = __float_as_uint(s1.w);
    V3 e0 = p1 - p0, e1 = p2 - p0;
    float bx, by;
    square_to_uniform_triangle(sx, sy, bx, by);
    p = fma3(e0, bx, fma3(e1, by, p0));
    pdf = m.inv_area;
    if (flags & TRI_HAS_NORMALS) {
        float4 s3 = __ldg(sp + 3), s4 = __ldg(sp + 4), s5 = __ldg(sp + 5);
        V3 n0 = v3(s3.x, s3.y, s3.z), n1 = v3(s3.w, s4.x, s4.y), n2 = v3(s4.z, s4.w, s5.x);
        n = fma3(n0, 1.f - bx - by, fma3(n1, bx, n2 * by));
    } else {
        n = cross3(e0, e1);
    }
    n = normalize3(n);
    if (flags & TRI_FLIP)
        n = -n;
}

struct PathOut {
    V3 rgb;
    float path_length;
    uint32_t depth;
};

DTOF_DEV void camera_ray(const dtof_camera &c, float u, float v, V3 &o, V3 &d, float &maxt) {
    const float *m = c.sample_to_camera;
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        r[i] = fmaf(m[4 * i + 2], 0.f, fmaf(m[4 * i + 1], v, fmaf(m[4 * i + 0], u, m[4 * i + 3])));
    V3 near_p = v3(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
    V3 dl = near_p * (1.f / sqrtf(dot3(near_p, near_p)));   // IEEE: the camera ray is compared bit for bit
    M34 tw;
#pragma unroll
    for (int i = 0; i < 12; ++i)
        tw.m[i] = c.to_world[i];
    o = v3(tw.m[3], tw.m[7], tw.m[11]);
    d = xf_vector(tw, dl);
    float inv_z = 1.f / dl.z;
    float near_t = c.near_clip * inv_z, far_t = c.far_clip * inv_z;
    o = o + d * near_t;
    maxt = far_t - near_t;
}

// ------------------------------------------------------------------------------------------------
// Film
struct FilmParams {
    float *rgbw;               // (H, W, 4) accumulation tensor
    uint32_t width, height, crop_x, crop_y, rfilter;
    float radius, inv_radius, g_alpha, g_bias;
    int n;                     // ceil(radius - .5)
    float mitchell_b, mitchell_c;
};

// The filters with negative lobes (src/rfilters/{mitchell,catmullrom,lanczos}.cpp). Out of line, and reachable only from
// the LOBED = true instantiations (a film with such a filter makes the scene "extended", the kernels' ENV parameter): the
// common kernels' code must stay as it is.
__device__ __noinline__ float rfilter_eval_lobed(uint32_t kind, float radius, float B, float Cc, float x) {
    x = fabsf(x);
    if (kind == DTOF_RFILTER_LANCZOS) {            // LanczosSincFilter::eval, lanczos.cpp:44-54 (radius = lobes)
        float x1 = 3.14159265358979323846f * x, x2 = x1 / radius, s1, s2, c;
        dr_sincos(x1, s1, c);
        dr_sincos(x2, s2, c);
        float result = (s1 * s2) / (x1 * x2);
        return x < 0x1p-24f ? 1.f : (x > radius ? 0.f : result);
    }
    float x2 = x * x, x3 = x2 * x, result;
    if (kind == DTOF_RFILTER_MITCHELL) {           // MitchellNetravaliFilter::eval, mitchell.cpp:62-83
        const float a3 = 12.f - 9.f * B - 6.f * Cc, a2 = -18.f + 12.f * B + 6.f * Cc, a0 = 6.f - 2.f * B;
        const float b3 = -B - 6.f * Cc, b2 = 6.f * B + 30.f * Cc, b1 = -12.f * B - 48.f * Cc, b0 = 8.f * B + 24.f * Cc;
        result = (1.f / 6.f) * (x < 1.f ? fmaf(a3, x3, fmaf(a2, x2, a0)) : fmaf(b3, x3, fmaf(b2, x2, fmaf(b1, x, b0))));
    } else {                                       // CatmullRomFilter::eval, catmullrom.cpp:40-55 (B = 0, C = 1/2, no fused ops)
        B = 0.f, Cc = .5f;
        result = (1.f / 6.f) * (x < 1.f ? (12.f - 9.f * B - 6.f * Cc) * x3 + (-18.f + 12.f * B + 6.f * Cc) * x2 + (6.f - 2.f * B)
                                        : (-B - 6.f * Cc) * x3 + (6.f * B + 30.f * Cc) * x2 + (-12.f * B - 48.f * Cc) * x +
                                              (8.f * B + 24.f * Cc));
    }
    return x < 2.f ? result : 0.f;
}

template <bool LOBED = false> DTOF_DEV float rfilter_eval(const FilmParams &F, float x) {
    if (F.rfilter == DTOF_RFILTER_TENT)
        return fmaxf(0.f, 1.f - fabsf(x * F.inv_radius));
    if (LOBED && F.rfilter >= DTOF_RFILTER_MITCHELL)
        return rfilter_eval_lobed(F.rfilter, F.radius, F.mitchell_b, F.mitchell_c, x);
    return fmaxf(0.f, expf(F.g_alpha * x * x) - F.g_bias);
}

// generic path: per-lane atomics (any filter, lanes of a warp on different pixels)
template <bool LOBED> DTOF_DEV void splat_generic(const FilmParams &F, float px, float py, V3 rgb) {
    if (F.rfilter == DTOF_RFILTER_BOX) {
        int x = (int) floorf(px) - (int) F.crop_x, y = (int) floorf(py) - (int) F.crop_y;
        if ((uint32_t) x < F.width && (uint32_t) y < F.height) {
            float *dst = F.rgbw + ((size_t) y * F.width + x) * 4;
            atomicAdd(dst + 0, rgb.x);
            atomicAdd(dst + 1, rgb.y);
            atomicAdd(dst + 2, rgb.z);
            atomicAdd(dst + 3, 1.f);
        }
        return;
    }
    int n = F.n, count = 2 * n + 1;
    int pix = (int) floorf(px) - n, piy = (int) floorf(py) - n;
    float rx = (float) pix + .5f - px, ry0 = (float) piy + .5f - py;
    int lx = pix - (int) F.crop_x, ly = piy - (int) F.crop_y;
    for (int xs = 0; xs < count; ++xs) {
        float wx = rfilter_eval<LOBED>(F, rx);
        rx += 1.f;
        uint32_t x = (uint32_t) (lx + xs);
        if (x >= F.width)
            continue;
        float ry = ry0;
        for (int ys = 0; ys < count; ++ys) {
            float wy = rfilter_eval<LOBED>(F, ry);
            ry += 1.f;
            uint32_t y = (uint32_t) (ly + ys);
            if (y >= F.height)
                continue;
            float w = wy * wx;
            float *dst = F.rgbw + ((size_t) y * F.width + x) * 4;
            atomicAdd(dst + 0, rgb.x * w);
            atomicAdd(dst + 1, rgb.y * w);
            atomicAdd(dst + 2, rgb.z * w);
            atomicAdd(dst + 3, w);
        }
    }
}

// Transposing butterfly: every lane contributes v[0..31]; afterwards lane l holds sum over lanes of v[l].
// 31 shuffles instead of 32 x 5.
DTOF_DEV float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            float keep = up ? v[i + off] : v[i];
            float send = up ? v[i] : v[i + off];
            v[i] = keep + __shfl_xor_sync(kFullMask, send, off);
        }
    }
    return v[0];
}
// 16-value variant: lane l (and l ^ 16) ends with the sum over all 32 lanes of v[l & 15]
DTOF_DEV float warp_transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            float keep = up ? v[i + off] : v[i];
            float send = up ? v[i] : v[i + off];
            v[i] = keep + __shfl_xor_sync(kFullMask, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(kFullMask, v[0], 16);
}

// Warp-aggregated 3x3 tent splat: all lanes of the warp belong to pixel (ix, iy) (pixel-major lanes), so the
// 9 taps x 4 channels are reduced across the warp first and only 36 atomics per warp reach L2.
DTOF_DEV void splat_tent3_warp(const FilmParams &F, int ix, int iy, float px, float py, V3 rgb, bool lane_on, int lane) {
    // ix, iy = floor(sample_pos) (identical on all lanes); taps ix-1..ix+1
    float rx = (float) (ix - 1) + .5f - px, ry = (float) (iy - 1) + .5f - py;
    float wx[3], wy[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        wx[i] = lane_on ? rfilter_eval(F, rx) : 0.f;
        wy[i] = lane_on ? rfilter_eval(F, ry) : 0.f;
        rx += 1.f;
        ry += 1.f;
    }
    float vc[32], vw[16];
#pragma unroll
    for (int ys = 0; ys < 3; ++ys)
#pragma unroll
        for (int xs = 0; xs < 3; ++xs) {
            float w = wy[ys] * wx[xs];
            int tap = ys * 3 + xs;
            vc[tap * 3 + 0] = rgb.x * w;
            vc[tap * 3 + 1] = rgb.y * w;
            vc[tap * 3 + 2] = rgb.z * w;
            vw[tap] = w;
        }
#pragma unroll
    for (int i = 27; i < 32; ++i)
        vc[i] = 0.f;
#pragma unroll
    for (int i = 9; i < 16; ++i)
        vw[i] = 0.f;
    float c = warp_transpose_reduce32(vc, lane);
    float w = warp_transpose_reduce16(vw, lane);
    int lx = ix - 1 - (int) F.crop_x, ly = iy - 1 - (int) F.crop_y;
    if (lane < 27) {
        int tap = lane / 3, ch = lane - tap * 3;
        uint32_t x = (uint32_t) (lx + tap % 3), y = (uint32_t) (ly + tap / 3);
        if (x < F.width && y < F.height)
            atomicAdd(F.rgbw + ((size_t) y * F.width + x) * 4 + ch, c);
    }
    if (lane < 9) {
        uint32_t x = (uint32_t) (lx + lane % 3), y = (uint32_t) (ly + lane / 3);
        if (x < F.width && y < F.height)
            atomicAdd(F.rgbw + ((size_t) y * F.width + x) * 4 + 3, w);
    }
}

} // namespace dtof
