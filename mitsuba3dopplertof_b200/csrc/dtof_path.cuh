// The bounce loop of the Doppler-ToF path tracer: DopplerToFPathIntegrator::sample
// (src/integrators/dopplertofpath.cpp:79-283) with the semantics of the reference's JIT variants
// (SURVEY.md Appendix A.6): no early `break`, every draw of an active iteration is consumed, state of a
// finished lane is frozen.
//
// Control flow is WARP-UNIFORM: the loop runs while any lane of the warp is alive and both traversals of a bounce
// (closest hit, shadow ray) are entered by all 32 lanes with a per-lane mask. That is what lets the
// flat traversal of tiny scenes share its triangle fetches across the warp (dtof_device.cuh: trace_flat), and it
// costs the per-lane BVH walk nothing.
#pragma once
#include "dtof_device.cuh"

namespace dtof {

enum TraversalMode : int {
    MODE_BVH_GLOBAL = 0,   // BVH walk, scene read through L1/L2 from HBM
    MODE_BVH_SMEM = 1,     // BVH walk, traversal data staged in shared memory
    MODE_FLAT_SMEM = 2     // flat warp-coherent walk over all triangles (tiny scenes), shared memory
};

// pointers a traversal needs, resolved once per CTA (global or shared memory)
struct TravPtrs {
    const float4 *N;    // BVH nodes
    const float4 *T;    // triangles in leaf order (BVH modes)
    const float4 *TF;   // triangles in scene order (flat mode)
    const float4 *I;    // instance records (8 x float4)
    const float4 *B;    // instance world boxes (2 x float4)
    SmemScene S;        // BVH_SMEM mode: shared-window addresses of N / T / I and of this lane's traversal stack
};

template <int MODE, bool STATS, bool ROBUST = false>   // ROBUST: caller-supplied rays (RaySlab::set)
DTOF_DEV bool trace_any_mode(const DeviceScene &S, const TravPtrs &P, bool any, V3 o, V3 d, float tmax, float time,
                             bool lane_active, Hit &hit, Counters &st) {
    if (MODE == MODE_FLAT_SMEM)
        return trace_flat<STATS>(P.TF, P.I, P.B, S.n_insts, any, o, d, tmax, time, lane_active, hit, st);
    if (!lane_active || !S.has_geometry)
        return false;
#ifndef DTOF_OLD_SMEM_WALK   // A/B builds only: the round-1 walk (generic pointers, local-memory stack) on the staged copy
    if (MODE == MODE_BVH_SMEM)
        return trace_bvh_smem<STATS, ROBUST>(P.S, S.root, any, o, d, tmax, time, hit, st);
#endif
    return trace_bvh<STATS, ROBUST>(P.N, P.T, P.I, S.root, any, o, d, tmax, time, hit, st);
}

// VelocityIntegrator::sample (src/integrators/velocity.cpp:113-127): the camera ray is intersected at t = 0 and at
// t = `time`; the sample is (t2 - t1) / time where both hits exist, else 0. Must be called by all lanes of a warp.
template <int MODE, bool STATS>
DTOF_DEV PathOut trace_velocity(const DeviceScene &S, const TravPtrs &TP, const dtof_params &P, bool lane_on, V3 ray_o,
                                V3 ray_d, float ray_maxt, Counters &st) {
    PathOut out{ v3(0, 0, 0), 0.f, 0 };
    const bool active = lane_on && P.max_depth != 0;
    float t[2];
    bool ok[2];
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        Hit h;
        h.gid = 0;
        h.inst = -1;
        h.t = 0.f;
        ok[k] = trace_any_mode<MODE, STATS>(S, TP, false, ray_o, ray_d, ray_maxt, k ? P.time : 0.f, active, h, st);
        t[k] = ok[k] ? h.t : 0.f;
    }
    if (active && ok[0] && ok[1]) {
        float v = (t[1] - t[0]) / P.time;
        out.rgb = v3(v, v, v);
        out.depth = 1;
    }
    return out;
}

// Per-lane state of one path between bounces: the loop-carried variables of DopplerToFPathIntegrator::sample
// (dopplertofpath.cpp:95-121). The fused kernel keeps it in registers for the whole path; the wavefront pipeline
// (dtof_wavefront.cuh) stores it in HBM between the stages of a bounce.
struct PathState {
    V3 throughput, result, prev_p;
    float path_length, prev_bsdf_pdf;
    float eta;   // product of the sampled lobes' relative indices of refraction (:251); only dielectrics change it
    uint32_t depth;
    bool valid_ray, prev_bsdf_delta, active;
    // eta: every BSDF in scope has eta = 1 on a live path (bs.eta is 0 only where the throughput is 0 too and the
    // lane stops), so `path_length += t * eta`, `eta *= bs.eta` and `rr_prob = tmax * eta^2` reduce to eta == 1.
    DTOF_DEV void init(bool on, bool env_visible) {
        throughput = v3(1, 1, 1), result = v3(0, 0, 0), prev_p = v3(0, 0, 0);
        path_length = 0.f, prev_bsdf_pdf = 1.f, eta = 1.f;
        depth = 0;
        valid_ray = env_visible;                  // !hide_emitters && the scene has an environment emitter (:102)
        prev_bsdf_delta = true;
        active = on;
    }
};

// The emitter-sampling term of a bounce, pending until its shadow ray has been traced (:214-226)
struct PendingNee {
    bool want;
    V3 thr, c, o, d;
    float maxt;
};

// One iteration of the reference's loop body (:136-276) for a lane whose path ray has been traced: interaction,
// emission, emitter sampling, BSDF eval + sample, next ray, throughput, Russian roulette. Nothing in it depends on
// the shadow ray except whether the emitter-sampling term is added, and the shadow ray draws no random numbers, so
// that term is returned as a pending contribution. `ray_o/ray_d/ray_maxt` hold the traced ray on entry and the next
// path ray on exit.
// DOPPLER = false is the stock path tracer (src/integrators/path.cpp:103-283): no time wrap, modulation weight 1,
// every draw is Sampler::next_1d / next_2d, i.e. the independent stream only (the path stream does not move).
// ENV compiles the constant environment emitter in (src/emitters/constant.cpp). It is a template parameter because the
// fused kernel runs at 64 registers: the extra live values cost it 4 % even when no scene uses them
// (profiles/r01_tuning.md), so scenes without an environment emitter run the ENV = false instantiation.
template <bool DOPPLER, bool ENV>
DTOF_DEV void shade_bounce(const DeviceScene &S, const float4 *__restrict__ I, const dtof_params &P, const Modulation &mod,
                           LaneSampler &smp, PathState &ps, const bool hit, const Hit &h, V3 &ray_o, V3 &ray_d,
                           float &ray_maxt, const float ray_time, const float emitter_pmf, PendingNee &nee) {
    const uint32_t max_depth = (uint32_t) P.max_depth, rr_depth = (uint32_t) P.rr_depth;   // -1 -> 0xffffffff
    const uint32_t n_em = S.n_emitters;
    V3 &throughput = ps.throughput, &result = ps.result;
    float &path_length = ps.path_length;
    uint32_t &depth = ps.depth;
    nee.want = false;
    const bool valid = hit;
    const bool correlate = DOPPLER && (depth + 1) < P.path_correlation_depth;   // :122
    SI si;
    uint32_t bsdf_flags = 0, bsdf_id = 0;
    V3 refl = v3(0, 0, 0);
    int32_t mesh_emitter = -1;
    if (valid) {
        compute_si(S, I, h, ray_d, ray_time, si);
        const MeshRec &mr = S.meshes[si.mesh];
        mesh_emitter = mr.emitter;
        bsdf_id = mr.bsdf;
        const float4 b0 = *reinterpret_cast<const float4 *>(&S.bsdfs[mr.bsdf]);   // {r, g, b, flags}
        bsdf_flags = __float_as_uint(b0.w);
        refl = v3(b0.x, b0.y, b0.z);
        path_length += ENV ? h.t * ps.eta : h.t;                        // :141 (eta == 1 without dielectrics)
    }
    // ---- direct emission (:150-168)
    if (valid && mesh_emitter >= 0) {
        const MeshRec &em_mesh = S.meshes[si.mesh];
        const EmitterRec em = S.emitters[mesh_emitter];
        V3 rel = si.p - ps.prev_p;                                      // DirectionSample(scene, si, prev_si)
        float dist = fsqrt(dot3(rel, rel));
        V3 dsd = rel / dist;
        float em_pdf = 0.f;
        if (!ps.prev_bsdf_delta) {                                      // AreaLight::pdf_direction, area.cpp:148-166
            float dp = dot3(dsd, si.sh_n);
            float pdf = em_mesh.inv_area, adp = fabsf(dp);
            pdf *= adp != 0.f ? fdiv(dist * dist, adp) : 0.f;
            em_pdf = (dp < 0.f ? pdf : 0.f) * emitter_pmf;
        }
        float mis_bsdf = mis_weight(ps.prev_bsdf_pdf, em_pdf);
        float lw = DOPPLER ? mod.eval(ray_time, path_length) : 1.f;
        V3 Le = (si.wi.z > 0.f && ps.prev_bsdf_pdf > 0.f) ? v3(em.vr, em.vg, em.vb) : v3(0, 0, 0);
        V3 c = Le * mis_bsdf * lw;
        result = v3(fmaf(throughput.x, c.x, result.x), fmaf(throughput.y, c.y, result.y),
                    fmaf(throughput.z, c.z, result.z));
    }
    // ---- the ray left the scene: the environment emitter is "hit" (si.emitter(scene), scene.h:583-594). The path
    // length is not advanced (:141); DirectionSample(scene, si, prev_si).d = -si.wi = ray_d (records.h:173-180)
    if (ENV && !valid && S.env_emitter >= 0) {
        float em_pdf = ps.prev_bsdf_delta ? 0.f : kInvFourPi * emitter_pmf;     // constant.cpp:141-146, scene.cpp:293-299
        float mis_bsdf = mis_weight(ps.prev_bsdf_pdf, em_pdf);
        float lw = DOPPLER ? mod.eval(ray_time, path_length) : 1.f;
        V3 Le = ps.prev_bsdf_pdf > 0.f ? v3(S.env_r, S.env_g, S.env_b) : v3(0, 0, 0);
        V3 c = Le * mis_bsdf * lw;
        result = v3(fmaf(throughput.x, c.x, result.x), fmaf(throughput.y, c.y, result.y),
                    fmaf(throughput.z, c.z, result.z));
    }
    const bool active_next = (depth + 1 < max_depth) && valid;         // :171
    const bool smooth = (bsdf_flags & 2u) != 0, twosided = (bsdf_flags & 1u) != 0;

    // ---- emitter sampling (:187-202): the 2D sample is always consumed, its value only when needed
    uint64_t e1a = DOPPLER ? smp.rng_path.step() : 0ull, e1b = smp.rng.step();
    uint64_t e2a = DOPPLER ? smp.rng_path.step() : 0ull, e2b = smp.rng.step();
    smp.draws += 2;
    bool active_em = active_next && smooth && n_em > 0, ds_delta = false;
    V3 em_weight = v3(0, 0, 0), wo = v3(0, 0, 0);
    float ds_dist = 0.f, ds_pdf = 0.f;
    if (active_em) {
        uint32_t index = 0;
        float sx = 0.f, sy = 0.f;
        V3 ds_p, ds_n, spec, ds_d;
        EmitterRec em = S.emitters[0];
        if (n_em > 1 || (em.kind != DTOF_EMITTER_POINT && em.kind != DTOF_EMITTER_SPOT)) {
            sx = u32_to_float(pcg_output(correlate ? e1a : e1b));
            sy = u32_to_float(pcg_output(correlate ? e2a : e2b));
        }
        if (n_em > 1) {                                                 // Scene::sample_emitter, scene.cpp:171-189
            float scaled = sx * (float) n_em;
            index = min((uint32_t) scaled, n_em - 1u);
            sx = scaled - (float) index;
            em = S.emitters[index];
        }
        if (em.kind == DTOF_EMITTER_POINT) {                            // PointLight::sample_direction, point.cpp:118-147
            ds_p = v3(em.px, em.py, em.pz);
            ds_pdf = 1.f;
            ds_delta = true;
            ds_d = ds_p - si.p;
            float dist2 = dot3(ds_d, ds_d), inv_dist = rsqrt_ieee(dist2);
            ds_dist = fsqrt(dist2);
            ds_d = ds_d * inv_dist;
            float f = inv_dist * inv_dist;
            spec = v3(em.vr * f, em.vg * f, em.vb * f);
        } else if (ENV && em.kind == DTOF_EMITTER_SPOT) {               // SpotLight::sample_direction (spot.cpp:180-215), falloff_curve (:146-154)
            const SpotRec &sr = S.spots[index];
            ds_p = v3(em.px, em.py, em.pz);
            ds_pdf = 1.f;
            ds_delta = true;
            ds_d = ds_p - si.p;
            // IEEE square root / division here: inside the falloff ramp one ulp of cos_theta (~0.99) is 4e-7 rad of angle,
            // i.e. up to 1.5e-4 of a small falloff value; the SFU approximations (2 ulp) would double that
            ds_dist = sqrtf(dot3(ds_d, ds_d));
            const float inv_dist = 1.f / ds_dist;
            ds_d = ds_d * inv_dist;
            const V3 nd = -ds_d;
            const float *M = sr.to_local;
            const V3 local = v3(fmaf(M[2], nd.z, fmaf(M[1], nd.y, M[0] * nd.x)), fmaf(M[5], nd.z, fmaf(M[4], nd.y, M[3] * nd.x)),
                                fmaf(M[8], nd.z, fmaf(M[7], nd.y, M[6] * nd.x)));
            const float cos_theta = local.z * (1.f / sqrtf(dot3(local, local)));
            const float beam_res = cos_theta >= sr.cos_beam ? 1.f : (sr.cutoff_angle - acosf(cos_theta)) * sr.inv_transition;
            const float falloff = cos_theta > sr.cos_cutoff ? beam_res : 0.f;
            const float f = falloff * (inv_dist * inv_dist);
            spec = v3(em.vr * f, em.vg * f, em.vb * f);
        } else if (ENV && em.kind == DTOF_EMITTER_DIRECTIONAL) {        // DirectionalEmitter::sample_direction, directional.cpp:149-176
            const V3 d = v3(em.px, em.py, em.pz);
            V3 rel = si.p - v3(S.env_cx, S.env_cy, S.env_cz);
            float radius = fmaxf(S.env_radius, fsqrt(dot3(rel, rel)));
            ds_dist = 2.f * radius;
            ds_p = si.p - d * ds_dist;
            ds_pdf = 1.f;
            ds_delta = true;
            ds_d = -d;
            spec = v3(em.vr, em.vg, em.vb);
        } else if (ENV && em.kind == DTOF_EMITTER_CONSTANT) {                  // ConstantBackgroundEmitter::sample_direction, constant.cpp:112-139
            ds_d = square_to_uniform_sphere(sx, sy);
            V3 rel = si.p - v3(S.env_cx, S.env_cy, S.env_cz);
            float radius = fmaxf(S.env_radius, fsqrt(dot3(rel, rel)));   // the sphere grows to contain the reference point
            ds_dist = 2.f * radius;
            ds_p = fma3(ds_d, ds_dist, si.p);
            ds_pdf = kInvFourPi;
            spec = v3(em.vr, em.vg, em.vb) / ds_pdf;
        } else {                                                        // AreaLight / Shape::sample_direction
            sample_position(S, S.meshes[em.mesh], sx, sy, ds_p, ds_n, ds_pdf);
            ds_d = ds_p - si.p;
            float dist2 = dot3(ds_d, ds_d);
            ds_dist = fsqrt(dist2);
            ds_d = ds_d / ds_dist;
            float dp = fabsf(dot3(ds_d, ds_n));
            float x = fdiv(dist2, dp);
            ds_pdf *= isfinite(x) ? x : 0.f;
            bool em_active = dot3(ds_d, ds_n) < 0.f && ds_pdf != 0.f;
            spec = em_active ? v3(em.vr, em.vg, em.vb) / ds_pdf : v3(0, 0, 0);
        }
        if (n_em > 1) {
            ds_pdf *= emitter_pmf;
            spec = spec * (float) n_em;
        }
        em_weight = spec;
        if (ds_pdf != 0.f) {                                            // spawn_ray_to, interaction.h:141-148
            nee.o = offset_p(si.p, si.n, ds_p - si.p);
            nee.d = ds_p - nee.o;
            float dist = fsqrt(dot3(nee.d, nee.d));
            nee.d = nee.d / dist;
            nee.maxt = dist * (1.f - kShadowEps);
            nee.want = true;
        }
        wo = v3(dot3(ds_d, si.sh_s), dot3(ds_d, si.sh_t), dot3(ds_d, si.sh_n));
    }
    // an occluded or zero-pdf emitter sample clears active_em (:190): the term is pending iff nee.want
    // ---- BSDF eval + sample (:206-210); sample_1 is drawn but unused by the diffuse lobe
    float s1 = 0.f;                                                     // lobe selection of the dielectric
    if (ENV)
        s1 = smp.template next_1d<DOPPLER>(correlate);
    else
        smp.template skip_1d<DOPPLER>();
    float s2x = smp.template next_1d<DOPPLER>(correlate), s2y = smp.template next_1d<DOPPLER>(correlate);
    V3 bsdf_val = v3(0, 0, 0), bsdf_weight = v3(0, 0, 0), bs_wo = v3(0, 0, 0);
    float bsdf_pdf = 0.f, bs_pdf = 0.f;                                 // zero-initialised BSDFSample3f
    bool sampled_delta = false;
    float bs_eta = 1.f;
    if (ENV && valid && (bsdf_flags & 256u)) {                           // RoughDielectric::eval / pdf / sample, roughdielectric.cpp:240-490
        const BsdfRec &br = S.bsdfs[bsdf_id];
        Microfacet distr;
        distr.ggx = (bsdf_flags & 128u) != 0, distr.au = br.pad0, distr.av = br.pad1;
        const float m_eta = br.eta_r, m_inv_eta = frcp(m_eta);
        const V3 wi = si.wi, wi_up = wi.z >= 0.f ? wi : -wi;             // mulsign(wi, cos_theta_i)
        if (wi.z != 0.f) {
            // ---- eval + pdf for the emitter sample
            const bool reflect = wi.z * wo.z > 0.f;
            const float eta = wi.z > 0.f ? m_eta : m_inv_eta, inv_eta = wi.z > 0.f ? m_inv_eta : m_eta;
            V3 m = normalize3(wi + wo * (reflect ? 1.f : eta));
            if (m.z < 0.f)
                m = -m;
            const float D = distr.eval(m), F = fresnel_r(dot3(wi, m), m_eta);
            const float G = distr.smith_g1(wi, m) * distr.smith_g1(wo, m);
            const float wim = dot3(wi, m), wom = dot3(wo, m), denom = wim + eta * wom;
            if (reflect) {
                bsdf_val = refl * fdiv(F * D * G, 4.f * fabsf(wi.z));
            } else {
                const float value = fabsf(fdiv(inv_eta * inv_eta * (1.f - F) * D * G * eta * eta * wim * wom, wi.z * (denom * denom)));
                bsdf_val = v3(br.k_r * value, br.k_g * value, br.k_b * value);
            }
            if (wim * wi.z > 0.f && wom * wo.z > 0.f) {
                const float dwh_dwo = reflect ? frcp(4.f * wom) : fdiv(eta * eta * wom, denom * denom);
                float prob = fdiv(D * distr.smith_g1(wi_up, m) * fabsf(dot3(wi_up, m)), wi_up.z);
                prob *= reflect ? F : 1.f - F;
                bsdf_pdf = prob * fabsf(dwh_dwo);
            }
            // ---- sample
            float pdf_m;
            const V3 ms = distr.sample(wi_up, s2x, s2y, pdf_m);
            float Fs, cos_theta_t, eta_it, eta_ti;
            const float wims = dot3(wi, ms);
            fresnel_dielectric(wims, m_eta, Fs, cos_theta_t, eta_it, eta_ti);
            const bool selected_r = s1 <= Fs;
            bs_pdf = pdf_m * (selected_r ? Fs : 1.f - Fs);
            float dwh;
            V3 weight;
            if (selected_r) {
                bs_wo = v3(fmaf(2.f * wims, ms.x, -wi.x), fmaf(2.f * wims, ms.y, -wi.y), fmaf(2.f * wims, ms.z, -wi.z));   // reflect(wi, m)
                weight = refl;
                dwh = frcp(4.f * dot3(bs_wo, ms));
            } else {
                const float c = fmaf(wims, eta_ti, cos_theta_t);          // refract(wi, m, cos_theta_t, eta_ti)
                bs_wo = v3(fmaf(ms.x, c, -(wi.x * eta_ti)), fmaf(ms.y, c, -(wi.y * eta_ti)), fmaf(ms.z, c, -(wi.z * eta_ti)));
                bs_eta = eta_it;
                const float f2 = eta_ti * eta_ti;
                weight = v3(br.k_r * f2, br.k_g * f2, br.k_b * f2);
                const float dn = wims + bs_eta * dot3(bs_wo, ms);
                dwh = fdiv(bs_eta * bs_eta * dot3(bs_wo, ms), dn * dn);
            }
            bs_pdf *= fabsf(dwh);
            if (pdf_m != 0.f)
                bsdf_weight = weight * distr.smith_g1(bs_wo, ms);
        }
    } else if (ENV && valid && (bsdf_flags & 64u)) {                            // RoughConductor::eval / pdf / sample, roughconductor.cpp:226-390
        const BsdfRec &br = S.bsdfs[bsdf_id];
        Microfacet distr;
        distr.ggx = (bsdf_flags & 128u) != 0, distr.au = br.pad0, distr.av = br.pad1;
        V3 wi = si.wi, wo_l = wo;
        if (twosided) {
            wo_l.z = mulsign(wo_l.z, wi.z);
            wi.z = fabsf(wi.z);
        }
        if (wi.z > 0.f && wo_l.z > 0.f) {
            const V3 H = normalize3(wo_l + wi);
            const float D = distr.eval(H), g1_i = distr.smith_g1(wi, H), c = dot3(wi, H);
            if (D != 0.f) {
                const float result = fdiv(D * g1_i * distr.smith_g1(wo_l, H), 4.f * wi.z);
                bsdf_val = v3(fresnel_conductor(c, br.eta_r, br.k_r) * (result * refl.x), fresnel_conductor(c, br.eta_g, br.k_g) * (result * refl.y),
                              fresnel_conductor(c, br.eta_b, br.k_b) * (result * refl.z));
            }
            if (c > 0.f && dot3(wo_l, H) > 0.f)
                bsdf_pdf = fdiv(D * g1_i, 4.f * wi.z);
        }
        if (wi.z > 0.f) {
            float pdf_m;
            const V3 m = distr.sample(wi, s2x, s2y, pdf_m);
            const float dwm = dot3(wi, m);
            bs_wo = v3(fmaf(2.f * dwm, m.x, -wi.x), fmaf(2.f * dwm, m.y, -wi.y), fmaf(2.f * dwm, m.z, -wi.z));   // reflect(wi, m)
            const bool ok = pdf_m != 0.f && bs_wo.z > 0.f;
            const float weight = distr.smith_g1(bs_wo, m);
            bs_pdf = fdiv(pdf_m, 4.f * dot3(bs_wo, m));
            if (ok)
                bsdf_weight = v3(fresnel_conductor(dwm, br.eta_r, br.k_r) * (weight * refl.x), fresnel_conductor(dwm, br.eta_g, br.k_g) * (weight * refl.y),
                                 fresnel_conductor(dwm, br.eta_b, br.k_b) * (weight * refl.z));
            if (twosided)
                bs_wo.z = mulsign(bs_wo.z, si.wi.z);
        }
    } else if (ENV && valid && (bsdf_flags & 32u)) {                            // SmoothPlastic::eval / pdf / sample, plastic.cpp:210-345
        const BsdfRec &br = S.bsdfs[bsdf_id];
        const float eta = br.eta_r, fdr_int = br.eta_g, inv_eta_2 = br.eta_b, ssw = br.pad0;
        float wi_z = si.wi.z, wo_z = wo.z;
        if (twosided) {
            wo_z = mulsign(wo_z, wi_z);
            wi_z = fabsf(wi_z);
        }
        // diffuse_reflectance / (1 - fdr_int [* diffuse_reflectance])
        const bool nonlinear = br.pad1 != 0.f;
        const V3 diff = v3(fdiv(refl.x, 1.f - (nonlinear ? refl.x * fdr_int : fdr_int)), fdiv(refl.y, 1.f - (nonlinear ? refl.y * fdr_int : fdr_int)),
                           fdiv(refl.z, 1.f - (nonlinear ? refl.z * fdr_int : fdr_int)));
        const float f_i = fresnel_r(wi_z, eta);
        const float prob_s0 = f_i * ssw, prob_d0 = (1.f - f_i) * (1.f - ssw);
        if (wi_z > 0.f && wo_z > 0.f) {
            const float f_o = fresnel_r(wo_z, eta);
            bsdf_val = diff * (kInvPi * wo_z * inv_eta_2 * (1.f - f_i) * (1.f - f_o));
            bsdf_pdf = kInvPi * wo_z * fdiv(prob_d0, prob_s0 + prob_d0);
        }
        if (wi_z > 0.f) {
            const float prob_specular = fdiv(prob_s0, prob_s0 + prob_d0), prob_diffuse = 1.f - prob_specular;
            if (s1 < prob_specular) {
                bs_wo = v3(-si.wi.x, -si.wi.y, wi_z);                    // reflect(wi) of the (flipped) incident direction
                bs_pdf = prob_specular;
                const float value = fdiv(f_i, bs_pdf);
                bsdf_weight = v3(value * br.k_r, value * br.k_g, value * br.k_b);
                sampled_delta = true;
            } else {
                bs_wo = square_to_cosine_hemisphere(s2x, s2y);
                bs_pdf = prob_diffuse * (kInvPi * bs_wo.z);
                const float f_o = fresnel_r(bs_wo.z, eta);
                bsdf_weight = diff * fdiv(inv_eta_2 * (1.f - f_i) * (1.f - f_o), prob_diffuse);
            }
            if (twosided)
                bs_wo.z = mulsign(bs_wo.z, si.wi.z);
        }
    } else if (valid && smooth) {
        float wi_z = si.wi.z, wo_z = wo.z;
        if (twosided) {                                                 // TwoSidedBRDF, twosided.cpp:111-125,219-235
            wo_z = mulsign(wo_z, wi_z);
            wi_z = fabsf(wi_z);
        }
        if (wi_z > 0.f && wo_z > 0.f) {                                 // SmoothDiffuse::eval_pdf, diffuse.cpp:160-176
            bsdf_val = refl * kInvPi * wo_z;
            bsdf_pdf = kInvPi * wo_z;
        }
        if (wi_z > 0.f) {                                               // SmoothDiffuse::sample, diffuse.cpp:101-125
            bs_wo = square_to_cosine_hemisphere(s2x, s2y);
            bs_pdf = kInvPi * bs_wo.z;
            if (bs_pdf > 0.f)
                bsdf_weight = refl;
            if (twosided)
                bs_wo.z = mulsign(bs_wo.z, si.wi.z);
        }
    }
    if (ENV && valid && (bsdf_flags & 4u)) {                             // SmoothConductor::sample, conductor.cpp:247-300
        float wi_z = twosided ? fabsf(si.wi.z) : si.wi.z;                // TwoSidedBRDF: |wi.z| in, sign restored on wo.z
        if (wi_z > 0.f) {
            const BsdfRec &br = S.bsdfs[bsdf_id];
            bs_wo = v3(-si.wi.x, -si.wi.y, si.wi.z);                     // reflect(wi); twosided: mulsign(|wi.z|, wi.z) = wi.z
            bs_pdf = 1.f;
            bsdf_weight = v3(refl.x * fresnel_conductor(wi_z, br.eta_r, br.k_r), refl.y * fresnel_conductor(wi_z, br.eta_g, br.k_g),
                             refl.z * fresnel_conductor(wi_z, br.eta_b, br.k_b));
            sampled_delta = true;                                        // bs.sampled_type = DeltaReflection
        }
    }
    bool sampled_null = false;
    if (ENV && valid && (bsdf_flags & 8u)) {     // SmoothDielectric::sample (dielectric.cpp:250-366), ThinDielectric (thindielectric.cpp:140-189)
        const BsdfRec &br = S.bsdfs[bsdf_id];
        const bool thin = (bsdf_flags & 16u) != 0;
        float r_i, cos_theta_t, eta_it, eta_ti;
        fresnel_dielectric(thin ? fabsf(si.wi.z) : si.wi.z, br.eta_r, r_i, cos_theta_t, eta_it, eta_ti);
        if (thin)
            r_i *= fdiv(2.f, 1.f + r_i);                                 // internal reflections: r' = r + trt + tr^3t + ..
        const bool selected_r = s1 <= r_i;
        bs_pdf = selected_r ? r_i : 1.f - r_i;
        if (selected_r) {
            bs_wo = v3(-si.wi.x, -si.wi.y, si.wi.z);                     // reflect(wi)
            bsdf_weight = refl;
        } else if (thin) {
            bs_wo = v3(-si.wi.x, -si.wi.y, -si.wi.z);                    // straight on: a Null interaction
            bsdf_weight = v3(br.k_r, br.k_g, br.k_b);
            sampled_null = true;
        } else {
            bs_wo = v3(-eta_ti * si.wi.x, -eta_ti * si.wi.y, cos_theta_t);   // refract(wi, cos_theta_t, eta_ti)
            bs_eta = eta_it;
            const float f2 = eta_ti * eta_ti;                            // radiance is scaled by the solid-angle compression
            bsdf_weight = v3(br.k_r * f2, br.k_g * f2, br.k_b * f2);
        }
        sampled_delta = true;
    }
    // ---- emitter sampling contribution (:214-226), added once the shadow ray is known to be unoccluded
    if (nee.want) {
        float mis_em = ds_delta ? 1.f : mis_weight(ds_pdf, bsdf_pdf);
        float lw = DOPPLER ? mod.eval(ray_time, path_length + ds_dist) : 1.f;
        nee.c = bsdf_val * em_weight * mis_em * lw;
        nee.thr = throughput;
    }
    // ---- BSDF sampling (:230-251)
    if (valid) {
        V3 wd = fma3(si.sh_n, bs_wo.z, fma3(si.sh_t, bs_wo.y, si.sh_s * bs_wo.x));
        ray_o = offset_p(si.p, si.n, wd);
        ray_d = wd;
        ray_maxt = 3.402823466e+38f;
        ps.prev_p = si.p;
    }
    throughput = throughput * bsdf_weight;
    ps.valid_ray = ps.valid_ray || (valid && !(ENV && sampled_null));   // :253-254: not for a Null interaction
    ps.prev_bsdf_pdf = bs_pdf;
    ps.prev_bsdf_delta = sampled_delta;                                 // has_flag(bsdf_sample.sampled_type, Delta), :250
    // ---- stopping criterion (:262-276)
    if (valid)
        depth += 1;
    float tmax = max3(throughput);
    if (ENV)
        ps.eta *= bs_eta;                                               // :251
    float rr_prob = fminf(ENV ? tmax * (ps.eta * ps.eta) : tmax, 0.95f);   // :264
    bool rr_active = depth >= rr_depth;
    float q = smp.template next_1d<DOPPLER>(correlate);                 // always drawn
    bool rr_continue = q < rr_prob;
    if (rr_active)
        throughput = throughput * frcp(rr_prob);
    ps.active = active_next && (!rr_active || rr_continue) && tmax != 0.f;
}

// Must be called by all 32 lanes of a warp; `lane_on` masks lanes without a sample.
//
// The loop body is a two-phase state machine around ONE inlined traversal (a single copy of the traversal code keeps
// the kernel's instruction footprint down):
//   phase 0  trace the path ray (closest hit), then run the reference's WHOLE loop body (shade_bounce).
//   phase 1  trace the shadow ray (any hit; Scene::ray_test inside sample_emitter_direction, scene.cpp:262-268) and
//            add the pending term when it is unoccluded: result = fma(thr_nee, c_nee, result), the very operation
//            and operands of :214-226.
// Only ~a dozen values live across the shadow traversal (instead of the whole surface interaction), which is what
// lets the kernel run at 4+ CTAs per SM. All lanes of a warp are always in the same phase.
template <int MODE, bool STATS, bool DOPPLER, bool ENV>
DTOF_DEV PathOut trace_path(const DeviceScene &S, const TravPtrs &TP, const dtof_params &P, const Modulation &mod,
                            LaneSampler &smp, bool lane_on, V3 ray_o, V3 ray_d, float ray_maxt, float time_in,
                            Counters &st) {
    PathOut out{ v3(0, 0, 0), 0.f, 0 };
    const float ray_time = (!DOPPLER || time_in < P.time) ? time_in : time_in - P.time;    // dopplertofpath.cpp:93
    PathState ps;
    ps.init(lane_on && P.max_depth != 0, ENV && S.env_emitter >= 0 && !P.hide_emitters);
    const float emitter_pmf = S.n_emitters ? 1.f / (float) S.n_emitters : 0.f;   // once per sample: IEEE

    // state handed from phase 0 to phase 1
    bool phase_shadow = false;
    PendingNee nee;
    nee.want = false;
    nee.thr = v3(0, 0, 0), nee.c = v3(0, 0, 0), nee.o = v3(0, 0, 0), nee.d = v3(0, 0, 1);
    nee.maxt = 0.f;

    while (phase_shadow || __any_sync(kFullMask, ps.active)) {
        Hit h;
        h.gid = 0;
        h.inst = -1;
        const bool hit = trace_any_mode<MODE, STATS>(S, TP, phase_shadow, phase_shadow ? nee.o : ray_o, phase_shadow ? nee.d : ray_d,
                                                     phase_shadow ? nee.maxt : ray_maxt, ray_time,
                                                     phase_shadow ? nee.want : ps.active, h, st);
        if (phase_shadow) {   // ---- phase 1: the pending emitter-sampling term (:214-226)
            phase_shadow = false;
            if (nee.want && !hit)
                ps.result = v3(fmaf(nee.thr.x, nee.c.x, ps.result.x), fmaf(nee.thr.y, nee.c.y, ps.result.y),
                               fmaf(nee.thr.z, nee.c.z, ps.result.z));
            continue;
        }
        // ---- phase 0
        phase_shadow = true;
        nee.want = false;
        if (!ps.active)
            continue;
        shade_bounce<DOPPLER, ENV>(S, TP.I, P, mod, smp, ps, hit, h, ray_o, ray_d, ray_maxt, ray_time, emitter_pmf, nee);
    }
    out.rgb = ps.valid_ray ? ps.result : v3(0, 0, 0);
    out.path_length = ps.path_length;
    out.depth = ps.depth;
    return out;
}

} // namespace dtof
