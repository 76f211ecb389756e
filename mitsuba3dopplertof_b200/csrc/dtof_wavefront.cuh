// Wavefront pipeline of the B200 Doppler-ToF path tracer (sm_100a): the same per-lane algorithm as the fused kernel
// (dtof_api.cu: render_kernel), cut into stages that exchange compacted ray / hit queues through HBM.
//
// Why it exists: in a scene whose BVH lives in HBM (millions of triangles) the rays of a warp need wildly different
// numbers of traversal steps, and the fused kernel -- one lane owns one path, the warp waits for its slowest ray --
// runs the walk at ~8 of 32 lanes (profiles/r01_render_c5_v5_ncu_summary.json). Here traversal is a kernel of its own
// with DYNAMIC FETCH (after Aila & Laine 2009): a persistent warp keeps 32 rays in flight and a lane that finishes
// its ray pulls the next one from the queue, so lane utilisation no longer depends on the slowest ray of a fixed
// group. Shading runs over compacted queues at full warps.
//
// Stages per batch of lanes and pass (all on one stream, no host round trip while max_depth is bounded):
//   wf_generate   sampler seeding / time / camera ray (render_sample, src/render/integrator.cpp:476-542)     -> queue 0
//   per bounce b: wf_trace<closest>  queue b -> hits                      (Scene::ray_intersect, scene.cpp:125-137)
//                 wf_shade           hits -> path state, queue b+1, shadow queue   (dopplertofpath.cpp:136-276)
//                 wf_trace<any>      shadow queue -> result += thr * c when unoccluded       (:214-226, scene.cpp:262-268)
//   wf_splat      ImageBlock::put of every lane of the batch, warp-aggregated             (imageblock.cpp:418-531)
// Queues are compacted with one atomicAdd per CTA (ballot + popc per warp, counts combined in shared memory). Per-lane results are bit-identical to the fused
// kernel: both run shade_bounce() / tri_test() / the same slab test; only the order of the film atomics differs.
#pragma once
#include "dtof_path.cuh"

namespace dtof {

constexpr int kWfBlock = 256;
constexpr int kWfMaxSets = 4;         // batches in flight (one stream and one set of queues each)
constexpr int kWfRing = 16;            // bounce slots in the counter ring
constexpr uint32_t kWfMiss = 0xffffffffu;
constexpr uint32_t kWfChunk = 256;    // rays a warp reserves from a queue per atomicAdd
#ifndef DTOF_WF_TRACE_CTAS
#define DTOF_WF_TRACE_CTAS 6
#endif
#ifndef DTOF_WF_SHADE_CTAS
#define DTOF_WF_SHADE_CTAS 3
#endif
constexpr int kWfShadeBlock = 256;

// counters of bounce b live at ring + kWfSlotWords * (b % kWfRing)
//
// RAY BINNING. A queue is filled from both ends: rays that cross the world box of an animated instance (class A: they will
// enter that instance and walk its BLAS -- in a scene with a multi-million-triangle instance that is a 5x longer walk that
// starts with a matrix inverse) are appended at the front, all other rays (class B: the static part only) at the back.
// Entry i of the queue lives at position wf_pos(i): i < nA ? i : n_slots - 1 - (i - nA). A traversal warp reserves chunks of
// consecutive entries, so its 32 lanes hold rays of ONE class: class-A warps enter the instance together (the entry ran at
// 3.6 of 32 lanes in mixed warps) and descend the same deep tree, class-B warps turn over short rays. Per-lane results do not
// depend on queue order. With no classifier boxes (WfArgs::n_cls = 0: shared-memory scenes, many instances) every ray is
// class A and the layout is the plain compacted queue.
constexpr int kWfSlotWords = 8;
enum : int { WF_N_RAY = 0, WF_N_SHADOW = 1, WF_FETCH_CLOSEST = 2, WF_FETCH_SHADOW = 3, WF_N_RAY_B = 4, WF_N_SHADOW_B = 5 };
constexpr int kWfMaxCls = 4;           // animated-instance boxes the classifier tests

struct WfBuffers {
    // per-lane path state, indexed by the lane's slot in the batch
    ulonglong2 *rng, *rng_path;     // PCG32 (state, inc) of the independent / path-correlated streams
    float4 *thr_len;                // throughput.xyz, path_length
    float4 *res_pdf;                // result.xyz, prev_bsdf_pdf
    float4 *prev_meta;              // prev_p.xyz, bits: depth | valid_ray << 30 | prev_bsdf_delta << 31
    float2 *film_pos;               // position handed to ImageBlock::put
    float *eta;                     // accumulated relative index of refraction; touched only in scenes with dielectrics
    // path-ray queues (ping-pong), compacted: {o.xyz, maxt}, {d.xyz, time}, lane slot
    float4 *q_o[2], *q_d[2];
    uint32_t *q_lane[2];
    // closest-hit records by queue entry: {t, u, v, gid} (gid = kWfMiss: no hit), instance
    float4 *hit;
    int32_t *hit_inst;
    // shadow-ray queue with the pending emitter-sampling term
    float4 *s_o, *s_d, *s_thr, *s_c;
    uint32_t *s_lane;
    uint32_t *ring;                 // kWfRing x kWfSlotWords counters
};

struct WfArgs {
    DeviceScene scene;
    dtof_camera cam;
    FilmParams film;
    dtof_params p;
    Modulation mod;
    WfBuffers buf;
    uint32_t spp_per_pass, pass;
    unsigned long long lane_begin, shard_block;
    uint32_t shard_count, shard_index;
    unsigned long long batch_begin;   // local lane index of slot 0
    uint32_t n_slots;                 // lanes of this batch
    uint32_t bounce;
    uint32_t fetch_threshold;         // dynamic fetch: refill when fewer lanes than this still hold a ray
    uint32_t inner_threshold;         // phase scheduling: inner-node steps run while this many lanes want one
    uint32_t double_step;             // ... and two steps per vote while this many want one (33 = never)
    uint32_t nodes_bytes, tris_bytes, insts_bytes;
    // ray binning: padded world boxes of the animated instances (over both keyframes), n_cls = 0 turns binning off
    uint32_t n_cls;
    float4 cls_lo[kWfMaxCls], cls_hi[kWfMaxCls];
};

DTOF_DEV uint32_t *wf_slot(const WfArgs &A, uint32_t bounce) { return A.buf.ring + kWfSlotWords * (bounce % kWfRing); }
// position of queue entry i (see RAY BINNING above)
DTOF_DEV uint32_t wf_pos(uint32_t i, uint32_t n_a, uint32_t n_slots) { return i < n_a ? i : n_slots - 1u - (i - n_a); }
// does the segment o + t d, t in [0, maxt], cross the box of an animated instance? (conservative slab test; class A)
DTOF_DEV bool wf_class_a(const WfArgs &A, V3 o, V3 d, float maxt) {
    if (A.n_cls == 0u)
        return true;
    const V3 id = v3(frcp(d.x), frcp(d.y), frcp(d.z));
    bool hit = false;
#pragma unroll
    for (int g = 0; g < kWfMaxCls; ++g) {
        if ((uint32_t) g < A.n_cls) {
            const float4 lo = A.cls_lo[g], hi = A.cls_hi[g];
            const float ax = (lo.x - o.x) * id.x, bx = (hi.x - o.x) * id.x, ay = (lo.y - o.y) * id.y, by = (hi.y - o.y) * id.y,
                        az = (lo.z - o.z) * id.z, bz = (hi.z - o.z) * id.z;
            const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
            const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), maxt));
            hit = hit || tn <= tf * 1.00001f;
        }
    }
    return hit;
}

// ------------------------------------------------------------------------------------------------
// Stage 1: render_sample() up to the camera ray (Doppler branch src/render/integrator.cpp:476-542, stock branch
// :409-472 for the `path` integrator).
template <int KIND>
__global__ void __launch_bounds__(kWfBlock) wf_generate_kernel(const __grid_constant__ WfArgs A) {
    constexpr bool STOCK_SAMPLE = KIND != DTOF_INTEGRATOR_DOPPLERTOFPATH;
    constexpr bool DOPPLER = !STOCK_SAMPLE;
    const WfBuffers &B = A.buf;
    const bool any_depth = A.p.max_depth != 0;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        wf_slot(A, 0)[WF_N_RAY] = any_depth ? A.n_slots : 0u;
    for (uint32_t s = blockIdx.x * kWfBlock + threadIdx.x; s < A.n_slots; s += gridDim.x * kWfBlock) {
        const unsigned long long li = A.batch_begin + s;
        unsigned long long idx64;
        if (A.shard_block) {   // local lane -> (block of this shard, offset) -> global lane
            unsigned long long j = li / A.shard_block, within = li - j * A.shard_block;
            idx64 = A.lane_begin + (j * A.shard_count + A.shard_index) * A.shard_block + within;
        } else {
            idx64 = A.lane_begin + li;
        }
        const uint32_t idx = (uint32_t) idx64;
        const uint32_t pixel = idx / A.spp_per_pass;
        const uint32_t py = pixel / A.film.width, px = pixel - py * A.film.width;
        LaneSampler smp;
        if (A.pass == 0) {
            smp.seed(A.p, idx);
        } else {   // the streams continue across passes (integrator.cpp:299-308)
            ulonglong2 r = B.rng[s];
            smp.rng.state = r.x, smp.rng.inc = r.y;
            if (DOPPLER) {
                ulonglong2 q = B.rng_path[s];
                smp.rng_path.state = q.x, smp.rng_path.inc = q.y;
            }
            smp.draws = 0;
        }
        const bool correlate_pixel = A.p.path_correlation_depth > 0;
        const float scale_x = 1.f / (float) A.film.width, scale_y = 1.f / (float) A.film.height;
        const float off_x = -(float) A.film.crop_x * scale_x, off_y = -(float) A.film.crop_y * scale_y;
        const float posx = (float) (px + A.film.crop_x), posy = (float) (py + A.film.crop_y);
        float jx, jy;
        if (STOCK_SAMPLE) {
            jx = smp.rng.next_f32(), jy = smp.rng.next_f32();
        } else {
            jx = smp.next_1d(correlate_pixel), jy = smp.next_1d(correlate_pixel);
        }
        float spx = posx + jx, spy = posy + jy;
        float ax = fmaf(spx, scale_x, off_x), ay = fmaf(spy, scale_y, off_y);
        float time = A.cam.shutter_open;
        if (A.cam.shutter_open_time > 0.f) {
            if (STOCK_SAMPLE)
                time += smp.rng.next_f32() * A.cam.shutter_open_time;
            else
                time += smp.next_time(A.p, idx, A.spp_per_pass, A.pass) * A.cam.shutter_open_time;
        }
        V3 o, d;
        float maxt;
        camera_ray(A.cam, ax, ay, o, d, maxt);
        const float ray_time = (!DOPPLER || time < A.p.time) ? time : time - A.p.time;   // dopplertofpath.cpp:93
        if (A.film.rfilter == DTOF_RFILTER_BOX) {
            spx = posx;
            spy = posy;
        }
        B.rng[s] = make_ulonglong2(smp.rng.state, smp.rng.inc);
        if (DOPPLER)
            B.rng_path[s] = make_ulonglong2(smp.rng_path.state, smp.rng_path.inc);
        B.thr_len[s] = make_float4(1.f, 1.f, 1.f, 0.f);
        B.res_pdf[s] = make_float4(0.f, 0.f, 0.f, 1.f);
        // depth 0, prev_bsdf_delta, valid_ray = environment emitter visible (dopplertofpath.cpp:102)
        B.prev_meta[s] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0x80000000u |
                                     ((A.scene.env_emitter >= 0 && !A.p.hide_emitters) ? 0x40000000u : 0u)));
        B.film_pos[s] = make_float2(spx, spy);
        if (A.scene.extended)
            B.eta[s] = 1.f;
        if (any_depth) {
            B.q_o[0][s] = make_float4(o.x, o.y, o.z, maxt);
            B.q_d[0][s] = make_float4(d.x, d.y, d.z, ray_time);
            B.q_lane[0][s] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Stage 2 / 4: closest-hit (ANY = false) or any-hit (ANY = true) traversal.
//
// Two things keep the lanes of a warp busy although their rays need very different work:
//  * DYNAMIC FETCH: a persistent warp keeps up to 32 rays in flight; when `32 - fetch_threshold` lanes have finished
//    their rays they pull new ones from the queue (one atomicAdd per refill).
//  * PHASE SCHEDULING instead of the while-while loop: every iteration the warp votes and runs ONE of two uniform
//    phases -- a box-test step for the lanes that hold an inner node, or leaf processing (triangle tests / instance
//    entry) for the lanes that hold a leaf. Inner steps run while at least `inner_threshold` lanes want one; lanes
//    holding a leaf wait until enough of them have gathered. In a while-while loop the inner loop runs until the LAST
//    lane reaches a leaf: with ~5 inner nodes between leaves on incoherent rays that is ~20 iterations of which a
//    lane uses 5 (measured: 10.8 of 32 lanes active, profiles/r01_tuning.md).
// The walk itself (slab test, child order, triangle test, tie-break, instance entry) is the one of trace_bvh
// (dtof_device.cuh), so hits are identical. The world-space ray is not kept live while a lane is inside an instance;
// it is re-read from the queue when the lane leaves it.
template <int MODE, bool ANY>
__global__ void __launch_bounds__(kWfBlock, DTOF_WF_TRACE_CTAS) wf_trace_kernel(const __grid_constant__ WfArgs A) {
    extern __shared__ float4 wf_smem[];
    const float4 *__restrict__ N = A.scene.nodes, *__restrict__ T = A.scene.tris, *__restrict__ I = A.scene.insts;
    if (MODE == MODE_BVH_SMEM) {
        const uint32_t nn = A.nodes_bytes / 16, nt = A.tris_bytes / 16, ni = A.insts_bytes / 16;
        float4 *sN = wf_smem, *sT = sN + nn, *sI = sT + nt;
        for (uint32_t i = threadIdx.x; i < nn; i += kWfBlock) sN[i] = A.scene.nodes[i];
        for (uint32_t i = threadIdx.x; i < nt; i += kWfBlock) sT[i] = A.scene.tris[i];
        for (uint32_t i = threadIdx.x; i < ni; i += kWfBlock) sI[i] = A.scene.insts[i];
        __syncthreads();
        N = sN, T = sT, I = sI;
    }
    const WfBuffers &B = A.buf;
    uint32_t *slot = wf_slot(A, A.bounce);
    const uint32_t n_a = slot[ANY ? WF_N_SHADOW : WF_N_RAY], n = n_a + slot[ANY ? WF_N_SHADOW_B : WF_N_RAY_B];
    uint32_t *fetch = slot + (ANY ? WF_FETCH_SHADOW : WF_FETCH_CLOSEST);
    const int q = A.bounce & 1;
    const float4 *__restrict__ qo = ANY ? B.s_o : B.q_o[q], *__restrict__ qd = ANY ? B.s_d : B.q_d[q];
    const int lane = threadIdx.x & 31;
    const uint32_t refill_at = 32u - min(A.fetch_threshold, 31u);   // idle lanes that trigger a refill (>= 1)
    const uint32_t inner_min = max(A.inner_threshold, 1u);

    int stack[kStackSize];
    int sp = 0, node = kDone, cur_inst = -1;
    uint32_t k = kWfMiss;
    V3 ro = v3(0, 0, 0), rd = v3(0, 0, 1);
    RaySlab R;
    R.set(ro, v3(0, 0, 0));
    float best = 0.f;
    Hit hit;
    hit.t = 0.f, hit.u = 0.f, hit.v = 0.f, hit.gid = 0, hit.inst = -1;
    bool found = false, exhausted = false;

    // the sentinel that marks the way out of an instance is popped like a leaf and handled in the leaf phase
#define DTOF_WF_POP() node = sp ? stack[--sp] : kDone

    uint32_t w_next = 0, w_end = 0;   // rays this warp has reserved from the queue (warp-uniform)
    for (;;) {
        // ---- vote: what does every lane hold?
        const bool is_inner = (unsigned) node < (unsigned) kDone, is_leaf = node < 0;
        const unsigned m_inner = __ballot_sync(kFullMask, is_inner), m_leaf = __ballot_sync(kFullMask, is_leaf);
        const unsigned m_idle = ~(m_inner | m_leaf);
        if ((uint32_t) __popc(m_idle) >= refill_at) {
            const bool idle = !is_inner && !is_leaf;
            // ---- finished rays hand over their results
            if (idle && k != kWfMiss) {
                if (ANY) {
                    if (!found) {   // unoccluded: result = fma(thr_nee, c_nee, result), dopplertofpath.cpp:214-226
                        const uint32_t s = B.s_lane[k];
                        const float4 thr = B.s_thr[k], c = B.s_c[k];
                        float4 r = B.res_pdf[s];
                        r.x = fmaf(thr.x, c.x, r.x), r.y = fmaf(thr.y, c.y, r.y), r.z = fmaf(thr.z, c.z, r.z);
                        B.res_pdf[s] = r;
                    }
                } else {
                    B.hit[k] = make_float4(hit.t, hit.u, hit.v, __uint_as_float(found ? hit.gid : kWfMiss));
                    B.hit_inst[k] = hit.inst;
                }
                k = kWfMiss;
            }
            // ---- dynamic fetch: the idle lanes take rays from the chunk the warp has reserved (one atomicAdd per
            // kWfChunk rays); lanes the chunk cannot serve wait for the next refill
            if (!exhausted) {
                if (w_next == w_end) {
                    uint32_t base = 0;
                    if (lane == 0)
                        base = atomicAdd(fetch, kWfChunk);
                    base = __shfl_sync(kFullMask, base, 0);
                    w_next = min(base, n), w_end = min(base + kWfChunk, n);
                    exhausted = w_next == w_end;
                }
                const uint32_t give = min((uint32_t) __popc(m_idle), w_end - w_next);
                const uint32_t rank = __popc(m_idle & ((1u << lane) - 1u));
                if (idle && rank < give) {
                    k = wf_pos(w_next + rank, n_a, A.n_slots);
                    const float4 a = qo[k], b = qd[k];
                    ro = v3(a.x, a.y, a.z), rd = v3(b.x, b.y, b.z);
                    best = a.w;
                    R.set(ro, v3(frcp(rd.x), frcp(rd.y), frcp(rd.z)));
                    sp = 0, cur_inst = -1, found = false;
                    hit.gid = 0, hit.inst = -1;
                    node = A.scene.root;
                }
                w_next += give;
                if (give)
                    continue;   // vote again with the new rays
            }
            if (m_idle == kFullMask) {
                if (exhausted)
                    break;
                continue;
            }
        }
        if ((uint32_t) __popc(m_inner) >= inner_min || !m_leaf) {
            // ---- phase A: one inner-node step, two when most of the warp wants one (halves the vote overhead)
            const int steps = (uint32_t) __popc(m_inner) >= A.double_step ? 2 : 1;
#pragma unroll 1
            for (int step = 0; step < steps; ++step)
            if ((unsigned) node < (unsigned) kDone) {
                const float4 *np = N + 4 * (size_t) node;
                const float4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
                bool h0, h1;
                float t0n, t1n;
                node_test(n0, n1, n2, R, best, h0, h1, t0n, t1n);
                int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
                if (h0 && h1) {
                    bool swap = t1n < t0n;
                    stack[sp++] = swap ? c0 : c1;
                    node = swap ? c1 : c0;
                } else if (h0 || h1) {
                    node = h0 ? c0 : c1;
                } else {
                    DTOF_WF_POP();
                }
            }
        } else if (is_leaf) {
            // ---- phase B: the lanes that hold a leaf process it
            const uint32_t code = (uint32_t) ~node;
            const uint32_t count = code & 15u;
            if (node == kSentinel) {   // leave the instance: back to the world-space ray
                const float4 a = qo[k], b = qd[k];
                ro = v3(a.x, a.y, a.z), rd = v3(b.x, b.y, b.z);
                R.set(ro, v3(frcp(rd.x), frcp(rd.y), frcp(rd.z)));
                cur_inst = -1;
                DTOF_WF_POP();
            } else if (count == 0) {   // animated instance: move the ray into its space (Embree semantics, enter_instance)
                cur_inst = (int) (code >> 4);
                const float4 *ip = I + 8 * (size_t) cur_inst;
                const float4 a = qo[k], b = qd[k];
                enter_instance(ip, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), b.w, ro, rd);
                R.set(ro, v3(frcp(rd.x), frcp(rd.y), frcp(rd.z)));
                stack[sp++] = kSentinel;
                node = __float_as_int(ip[6].z);
            } else {
                const uint32_t first = code >> 4;
                bool terminate = false;
                for (uint32_t i = 0; i < count; ++i) {
                    const float4 *tp = T + 3 * (size_t) (first + i);
                    float4 a = tp[0], b = tp[1], c = tp[2];
                    float t, u, v;
                    if (tri_test(a, b, c, ro, rd, best, t, u, v)) {
                        if (ANY) {
                            found = true;
                            terminate = true;
                            break;
                        }
                        uint32_t gid = __float_as_uint(a.w);
                        if (t < best || !found || gid < hit.gid) {
                            best = t;
                            hit.t = t, hit.u = u, hit.v = v;
                            hit.gid = gid;
                            hit.inst = cur_inst;
                            found = true;
                        }
                    }
                }
                if (terminate)
                    node = kDone;
                else
                    DTOF_WF_POP();
            }
        }
    }
#undef DTOF_WF_POP
}

// ------------------------------------------------------------------------------------------------
// Stage 3: one iteration of the bounce loop for every entry of the path-ray queue (full warps).
template <bool DOPPLER, bool ENV>
__global__ void __launch_bounds__(kWfShadeBlock, DTOF_WF_SHADE_CTAS * (kWfBlock / kWfShadeBlock))
wf_shade_kernel(const __grid_constant__ WfArgs A) {
    const WfBuffers &B = A.buf;
    uint32_t *slot = wf_slot(A, A.bounce), *next = wf_slot(A, A.bounce + 1);
    const uint32_t n_a = slot[WF_N_RAY], n = n_a + slot[WF_N_RAY_B];
    const int q = A.bounce & 1, qn = q ^ 1;
    const int lane = threadIdx.x & 31;
    const float emitter_pmf = A.scene.n_emitters ? 1.f / (float) A.scene.n_emitters : 0.f;
    const uint32_t stride = gridDim.x * kWfShadeBlock;
    __shared__ uint32_t sh_count[4][kWfShadeBlock / 32], sh_base[4];   // next path queue A / B, shadow queue A / B
    const int warp = threadIdx.x >> 5;
    for (uint32_t bbase = blockIdx.x * kWfShadeBlock; bbase < n; bbase += stride) {   // block-uniform trip count
        const bool on = bbase + threadIdx.x < n;
        const uint32_t k = wf_pos(bbase + threadIdx.x, n_a, A.n_slots);
        PathState ps;
        PendingNee nee;
        nee.want = false;
        ps.active = false;
        V3 ray_o = v3(0, 0, 0), ray_d = v3(0, 0, 1);
        float ray_maxt = 0.f, ray_time = 0.f;
        uint32_t s = 0;
        if (on) {
            s = B.q_lane[q][k];
            const float4 d4 = B.q_d[q][k], h4 = B.hit[k];
            ray_d = v3(d4.x, d4.y, d4.z);
            ray_time = d4.w;
            Hit h;
            h.t = h4.x, h.u = h4.y, h.v = h4.z;
            h.gid = __float_as_uint(h4.w);
            h.inst = B.hit_inst[k];
            const bool hit = h.gid != kWfMiss;
            const float4 a = B.thr_len[s], r = B.res_pdf[s], c = B.prev_meta[s];
            const uint32_t meta = __float_as_uint(c.w);
            ps.throughput = v3(a.x, a.y, a.z), ps.path_length = a.w;
            ps.result = v3(r.x, r.y, r.z), ps.prev_bsdf_pdf = r.w;
            ps.prev_p = v3(c.x, c.y, c.z);
            ps.depth = meta & 0x3fffffffu;
            ps.valid_ray = (meta & 0x40000000u) != 0;
            ps.prev_bsdf_delta = (meta & 0x80000000u) != 0;
            ps.eta = ENV ? B.eta[s] : 1.f;
            ps.active = true;
            LaneSampler smp;
            const ulonglong2 g = B.rng[s];
            smp.rng.state = g.x, smp.rng.inc = g.y;
            if (DOPPLER) {
                const ulonglong2 gp = B.rng_path[s];
                smp.rng_path.state = gp.x, smp.rng_path.inc = gp.y;
            }
            smp.draws = 0;
            shade_bounce<DOPPLER, ENV>(A.scene, A.scene.insts, A.p, A.mod, smp, ps, hit, h, ray_o, ray_d, ray_maxt, ray_time,
                                  emitter_pmf, nee);
            B.rng[s] = make_ulonglong2(smp.rng.state, smp.rng.inc);
            if (DOPPLER)
                B.rng_path[s] = make_ulonglong2(smp.rng_path.state, smp.rng_path.inc);
            if (ENV)
                B.eta[s] = ps.eta;
            B.thr_len[s] = make_float4(ps.throughput.x, ps.throughput.y, ps.throughput.z, ps.path_length);
            B.res_pdf[s] = make_float4(ps.result.x, ps.result.y, ps.result.z, ps.prev_bsdf_pdf);
            B.prev_meta[s] = make_float4(ps.prev_p.x, ps.prev_p.y, ps.prev_p.z,
                                         __uint_as_float(ps.depth | (ps.valid_ray ? 0x40000000u : 0u) |
                                                         (ps.prev_bsdf_delta ? 0x80000000u : 0u)));
        }
        // ---- compaction: one atomicAdd per CTA and queue. All warps of the machine count into the same two words, and
        // same-address atomics serialise in L2: with one atomicAdd per warp the kernel spent 46 % of its stall samples
        // waiting for them (profiles/r01_tuning.md).
        // class of the two new rays (RAY BINNING): A = crosses an animated instance's box
        const bool want_next = on && ps.active, want_shadow = on && nee.want;
        const bool next_a = want_next && wf_class_a(A, ray_o, ray_d, ray_maxt);
        const bool shadow_a = want_shadow && wf_class_a(A, nee.o, nee.d, nee.maxt);
        const unsigned m[4] = { __ballot_sync(kFullMask, next_a), __ballot_sync(kFullMask, want_next && !next_a),
                                __ballot_sync(kFullMask, shadow_a), __ballot_sync(kFullMask, want_shadow && !shadow_a) };
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                sh_count[c][warp] = __popc(m[c]);
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < kWfShadeBlock / 32; ++w)
                total += sh_count[threadIdx.x][w];
            uint32_t *ctr = threadIdx.x == 0 ? next + WF_N_RAY : threadIdx.x == 1 ? next + WF_N_RAY_B
                          : threadIdx.x == 2 ? slot + WF_N_SHADOW : slot + WF_N_SHADOW_B;
            sh_base[threadIdx.x] = total ? atomicAdd(ctr, total) : 0u;
        }
        __syncthreads();
        uint32_t base[4] = { sh_base[0], sh_base[1], sh_base[2], sh_base[3] };
        for (int w = 0; w < warp; ++w) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                base[c] += sh_count[c][w];
        }
        __syncthreads();   // sh_count / sh_base are rewritten by the next iteration
        const unsigned lt = (1u << lane) - 1u;
        if (want_next) {
            const uint32_t j = next_a ? base[0] + __popc(m[0] & lt) : A.n_slots - 1u - (base[1] + __popc(m[1] & lt));
            B.q_o[qn][j] = make_float4(ray_o.x, ray_o.y, ray_o.z, ray_maxt);
            B.q_d[qn][j] = make_float4(ray_d.x, ray_d.y, ray_d.z, ray_time);
            B.q_lane[qn][j] = s;
        }
        if (want_shadow) {
            const uint32_t j = shadow_a ? base[2] + __popc(m[2] & lt) : A.n_slots - 1u - (base[3] + __popc(m[3] & lt));
            B.s_o[j] = make_float4(nee.o.x, nee.o.y, nee.o.z, nee.maxt);
            B.s_d[j] = make_float4(nee.d.x, nee.d.y, nee.d.z, ray_time);
            B.s_thr[j] = make_float4(nee.thr.x, nee.thr.y, nee.thr.z, 0.f);
            B.s_c[j] = make_float4(nee.c.x, nee.c.y, nee.c.z, 0.f);
            B.s_lane[j] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Stage 5: film. Slots are in lane order (pixel-major), so a warp usually splats one 3x3 footprint.
__global__ void __launch_bounds__(kWfBlock) wf_splat_kernel(const __grid_constant__ WfArgs A) {
    const WfBuffers &B = A.buf;
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * kWfBlock;
    for (uint32_t base = blockIdx.x * kWfBlock + (threadIdx.x & ~31u); base < A.n_slots; base += stride) {
        const uint32_t s = base + lane;
        const bool lane_on = s < A.n_slots;
        V3 rgb = v3(0, 0, 0);
        float spx = 0.f, spy = 0.f;
        if (lane_on) {
            const float4 r = B.res_pdf[s];
            const uint32_t meta = __float_as_uint(B.prev_meta[s].w);
            const float2 fp = B.film_pos[s];
            if (meta & 0x40000000u)   // valid_ray (dopplertofpath.cpp:279)
                rgb = v3(r.x, r.y, r.z);
            spx = fp.x, spy = fp.y;
        }
        const int ix = (int) floorf(spx), iy = (int) floorf(spy);
        const unsigned on_mask = __ballot_sync(kFullMask, lane_on);
        const int leader = __ffs(on_mask) - 1;
        const int lx = __shfl_sync(kFullMask, ix, leader), ly = __shfl_sync(kFullMask, iy, leader);
        const bool uniform = __all_sync(kFullMask, !lane_on || (ix == lx && iy == ly));
        if (uniform && A.film.rfilter == DTOF_RFILTER_TENT && A.film.n == 1)
            splat_tent3_warp(A.film, lx, ly, spx, spy, rgb, lane_on, lane);
        else if (lane_on)
            splat_generic<true>(A.film, spx, spy, rgb);
    }
}

} // namespace dtof
