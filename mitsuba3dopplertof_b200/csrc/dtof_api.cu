// libdtof_b200.so -- C ABI (include/dtof.h) + render kernels for sm_100a.
//
// Kernel design: a persistent, register-resident "fused wavefront": one thread per wavefront
// lane; every WARP pulls work units of kUnit consecutive lanes (part of ONE pixel when spp_per_pass >= 32) from an
// atomic counter -- no block-wide barrier after start-up. Traversal data that fits in shared memory is staged
// there once per CTA with 128-bit copies; tiny scenes are walked flat and warp-coherently, larger ones through the
// BVH, huge ones through L1/L2 with 128-bit loads. Film accumulation is warp-aggregated (36 atomics per warp
// instead of 36 per lane).
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a profiler (nsys / ncu --nvtx) is attached

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dtof.h"
#include "dtof_bvh.h"
#include "dtof_device.cuh"
#include "dtof_layout.h"
#include "dtof_path.cuh"
#include "dtof_wavefront.cuh"

using namespace dtof;

namespace {

// NVTX range per stage (the reference marks its stages with ScopedPhase, include/mitsuba/core/profiler.h:20-49): host-side
// push / pop around the enqueueing calls, which is what nsys needs to attribute the kernels of a stage to it.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// CTA shape of the fused kernel: ONE CTA of 1024 threads per SM (64 registers / thread = the whole register file). Measured
// against 4 x 256 and 2 x 512 at the same occupancy (profiles/r02_tuning.md): +3.2 .. +4.7 % on C1-C3 -- the scene is staged
// once per SM instead of four times and the 32 warps of an SM share one set of shared-memory structures.
constexpr int kBlock = DTOF_BLOCK;   // dtof_device.cuh (1024 unless an A/B build overrides it)
#ifndef DTOF_MIN_CTAS
#define DTOF_MIN_CTAS 1
#endif
constexpr unsigned long long kUnit = 256;       // lanes per work unit fetched by one warp (8 warp iterations)
// traversal data + traversal stacks of one CTA up to this size are staged in shared memory (227 KB opt-in per SM)
constexpr size_t kSmemSceneLimit = (DTOF_MIN_CTAS == 1 ? 200 : 96) * 1024;
constexpr uint32_t kFlatMaxTris = 0;            // flat coherent walk: never automatic (DTOF_MODE=2 only), see r01_tuning.md

struct RenderArgs {
    DeviceScene scene;
    const float4 *tris_flat;                     // triangles in scene order (flat mode)
    const float4 *inst_box;                      // instance world boxes
    dtof_camera cam;
    FilmParams film;
    dtof_params p;
    Modulation mod;
    uint32_t spp_per_pass, n_passes;
    unsigned long long lane_begin, lane_end;     // lanes of every pass handled by this launch
    unsigned long long n_local;                  // number of lanes this launch renders (after interleaved sharding)
    unsigned long long shard_block;
    uint32_t shard_count, shard_index;
    unsigned long long *work_counter;            // work-unit dispenser
    Counters *stats;
    // record mode
    const unsigned long long *rec_lanes;
    dtof_sample_record *rec_out;
    uint32_t n_rec;
    uint32_t nodes_bytes, tris_bytes, insts_bytes, boxes_bytes;
    uint32_t stack_levels;                       // BVH_SMEM mode: rows of the shared-memory traversal stack (BVH depth + 4)
    uint32_t rec_pass;                           // record mode: the pass whose lanes are recorded (passes before it are replayed)
};

template <int MODE, bool STATS, bool RECORD, int KIND, bool ENV>
__global__ void __launch_bounds__(kBlock, DTOF_MIN_CTAS) render_kernel(const __grid_constant__ RenderArgs A) {
    extern __shared__ float4 smem[];
    TravPtrs TP;
    TP.N = A.scene.nodes, TP.T = A.scene.tris, TP.TF = A.tris_flat, TP.I = A.scene.insts, TP.B = A.inst_box;
    TP.S = SmemScene{ 0u, 0u, 0u, 0u };
    if (MODE != MODE_BVH_GLOBAL) {
        // stage the traversal data into shared memory (128-bit copies): [nodes | tris] or [flat tris], insts, boxes
        const uint32_t nn = MODE == MODE_BVH_SMEM ? A.nodes_bytes / 16 : 0, nt = A.tris_bytes / 16,
                       ni = A.insts_bytes / 16, nb = A.boxes_bytes / 16;
        float4 *sN = smem, *sT = sN + nn, *sI = sT + nt, *sB = sI + ni;
        const float4 *gT = MODE == MODE_BVH_SMEM ? A.scene.tris : A.tris_flat;
        if constexpr (MODE == MODE_BVH_SMEM)
            for (uint32_t i = threadIdx.x; i < nn; i += kBlock) sN[i] = A.scene.nodes[i];
        for (uint32_t i = threadIdx.x; i < nt; i += kBlock) sT[i] = gT[i];
        for (uint32_t i = threadIdx.x; i < ni; i += kBlock) sI[i] = A.scene.insts[i];
        for (uint32_t i = threadIdx.x; i < nb; i += kBlock) sB[i] = A.inst_box[i];
        __syncthreads();
        TP.N = sN, TP.T = sT, TP.TF = sT, TP.I = sI, TP.B = sB;
        if (MODE == MODE_BVH_SMEM) {   // after the scene: the traversal stacks, stack_levels rows of 128 B per warp
            const uint32_t base = opaque_u32((uint32_t) __cvta_generic_to_shared(smem));
            TP.S.N = base, TP.S.T = base + A.nodes_bytes, TP.S.I = TP.S.T + A.tris_bytes;
            TP.S.stack = TP.S.I + A.insts_bytes + A.boxes_bytes + (threadIdx.x >> 5) * (A.stack_levels * 128u) + (threadIdx.x & 31) * 4u;
        }
    }
    constexpr bool VELOCITY = KIND == DTOF_INTEGRATOR_VELOCITY;
    constexpr bool STOCK_SAMPLE = KIND != DTOF_INTEGRATOR_DOPPLERTOFPATH;   // stock render_sample branch (integrator.cpp:409-472)
    const int lane = threadIdx.x & 31;
    Counters st = {};
    const unsigned long long n_lanes = RECORD ? (unsigned long long) A.n_rec : A.n_local;
    const unsigned long long n_units = (n_lanes + kUnit - 1) / kUnit;

    for (;;) {
        unsigned long long unit = 0;
        if (lane == 0)
            unit = atomicAdd(A.work_counter, 1ull);
        unit = __shfl_sync(kFullMask, unit, 0);
        if (unit >= n_units)
            break;
        for (unsigned long long li = unit * kUnit + lane; li < (unit + 1) * kUnit; li += 32) {
            if (li - lane >= n_lanes)   // warp-uniform: the whole iteration is past the end
                break;
            const bool lane_on = li < n_lanes;
            unsigned long long idx64 = 0;
            if (lane_on) {
                if (RECORD) {
                    idx64 = A.rec_lanes[li];
                } else if (A.shard_block) {   // local lane -> (block of this shard, offset) -> global lane
                    unsigned long long j = li / A.shard_block, within = li - j * A.shard_block;
                    idx64 = A.lane_begin + (j * A.shard_count + A.shard_index) * A.shard_block + within;
                } else {
                    idx64 = A.lane_begin + li;
                }
            }
            const uint32_t idx = (uint32_t) idx64;
            const uint32_t pixel = idx / A.spp_per_pass;
            const uint32_t py = pixel / A.film.width, px = pixel - py * A.film.width;
            LaneSampler smp;
            smp.seed(A.p, idx);

            for (uint32_t pass = 0; pass < (RECORD ? A.rec_pass + 1u : A.n_passes); ++pass) {
                // render_sample(): Doppler branch (src/render/integrator.cpp:476-542), or the stock branch (:409-472)
                // for the velocity and path integrators -- jitter and time from the independent stream only
                const bool correlate_pixel = A.p.path_correlation_depth > 0;
                const float scale_x = 1.f / (float) A.film.width, scale_y = 1.f / (float) A.film.height;
                const float off_x = -(float) A.film.crop_x * scale_x, off_y = -(float) A.film.crop_y * scale_y;
                const float posx = (float) (px + A.film.crop_x), posy = (float) (py + A.film.crop_y);
                float jx, jy;
                if (STOCK_SAMPLE) {
                    jx = smp.rng.next_f32(), jy = smp.rng.next_f32();
                    smp.draws += 2;
                } else {
                    jx = smp.next_1d(correlate_pixel), jy = smp.next_1d(correlate_pixel);
                }
                float spx = posx + jx, spy = posy + jy;
                float ax = fmaf(spx, scale_x, off_x), ay = fmaf(spy, scale_y, off_y);
                float time = A.cam.shutter_open;
                if (A.cam.shutter_open_time > 0.f) {
                    if (STOCK_SAMPLE) {
                        time += smp.rng.next_f32() * A.cam.shutter_open_time;
                        smp.draws++;
                    } else {
                        time += smp.next_time(A.p, idx, A.spp_per_pass, pass) * A.cam.shutter_open_time;
                    }
                }
                V3 o, d;
                float maxt;
                camera_ray(A.cam, ax, ay, o, d, maxt);
                PathOut r = VELOCITY ? trace_velocity<MODE, STATS>(A.scene, TP, A.p, lane_on, o, d, maxt, st)
                                     : trace_path<MODE, STATS, KIND == DTOF_INTEGRATOR_DOPPLERTOFPATH, ENV>(A.scene, TP, A.p, A.mod, smp,
                                                                                                       lane_on, o, d, maxt, time, st);
                V3 rgb = r.rgb;
                if (A.film.rfilter == DTOF_RFILTER_BOX) {
                    spx = posx;
                    spy = posy;
                }
                if (STATS && lane_on) st.samples++;
                if (RECORD) {
                    if (lane_on && pass == A.rec_pass) {
                        dtof_sample_record &rec = A.rec_out[li];
                        rec.sample_pos[0] = spx, rec.sample_pos[1] = spy;
                        rec.time = time;
                        rec.ray_o[0] = o.x, rec.ray_o[1] = o.y, rec.ray_o[2] = o.z;
                        rec.ray_d[0] = d.x, rec.ray_d[1] = d.y, rec.ray_d[2] = d.z;
                        rec.ray_maxt = maxt;
                        rec.rgb[0] = rgb.x, rec.rgb[1] = rgb.y, rec.rgb[2] = rgb.z;
                        rec.path_length = r.path_length;
                        rec.depth = r.depth;
                        rec.rng_draws = smp.draws;
                    }
                } else {
                    // ---- film: warp-aggregated when the whole warp splats the same 3x3 footprint
                    const int ix = (int) floorf(spx), iy = (int) floorf(spy);
                    const unsigned on_mask = __ballot_sync(kFullMask, lane_on);
                    const int leader = __ffs(on_mask) - 1;
                    const int lx = __shfl_sync(kFullMask, ix, leader), ly = __shfl_sync(kFullMask, iy, leader);
                    const bool uniform = __all_sync(kFullMask, !lane_on || (ix == lx && iy == ly));
                    if (uniform && A.film.rfilter == DTOF_RFILTER_TENT && A.film.n == 1)
                        splat_tent3_warp(A.film, lx, ly, spx, spy, rgb, lane_on, lane);
                    else if (lane_on)
                        splat_generic<ENV>(A.film, spx, spy, rgb);
                }
            }
        }
    }
    if (STATS) {
        atomicAdd(&A.stats->rays_closest, st.rays_closest);
        atomicAdd(&A.stats->rays_shadow, st.rays_shadow);
        atomicAdd(&A.stats->nodes, st.nodes);
        atomicAdd(&A.stats->tris, st.tris);
        atomicAdd(&A.stats->inst, st.inst);
        atomicAdd(&A.stats->samples, st.samples);
    }
}

// dtof_trace_rays: one caller-supplied ray per thread through the traversal the render kernels use
template <int MODE>
__global__ void __launch_bounds__(kBlock) rays_kernel(const __grid_constant__ RenderArgs A, const dtof_ray *__restrict__ rays, uint32_t n,
                                                      int any_hit, dtof_ray_hit *__restrict__ out) {
    extern __shared__ float4 smem[];
    TravPtrs TP;
    TP.N = A.scene.nodes, TP.T = A.scene.tris, TP.TF = A.tris_flat, TP.I = A.scene.insts, TP.B = A.inst_box;
    TP.S = SmemScene{ 0u, 0u, 0u, 0u };
    if (MODE == MODE_BVH_SMEM) {
        const uint32_t nn = A.nodes_bytes / 16, nt = A.tris_bytes / 16, ni = A.insts_bytes / 16;
        float4 *sN = smem, *sT = sN + nn, *sI = sT + nt;
        for (uint32_t i = threadIdx.x; i < nn; i += kBlock) sN[i] = A.scene.nodes[i];
        for (uint32_t i = threadIdx.x; i < nt; i += kBlock) sT[i] = A.scene.tris[i];
        for (uint32_t i = threadIdx.x; i < ni; i += kBlock) sI[i] = A.scene.insts[i];
        __syncthreads();
        const uint32_t base = opaque_u32((uint32_t) __cvta_generic_to_shared(smem));
        TP.S.N = base, TP.S.T = base + A.nodes_bytes, TP.S.I = TP.S.T + A.tris_bytes;
        TP.S.stack = TP.S.I + A.insts_bytes + A.boxes_bytes + (threadIdx.x >> 5) * (A.stack_levels * 128u) + (threadIdx.x & 31) * 4u;
    }
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    const bool on = i < n;
    dtof_ray r{};
    if (on)
        r = rays[i];
    Counters st = {};
    Hit h;
    h.t = 0.f, h.u = 0.f, h.v = 0.f, h.gid = 0, h.inst = -1;
    const bool hit = trace_any_mode<MODE, true, true>(A.scene, TP, any_hit != 0, v3(r.o[0], r.o[1], r.o[2]), v3(r.d[0], r.d[1], r.d[2]), r.tmax,
                                                r.time, on, h, st);
    if (on) {
        dtof_ray_hit o{};
        o.hit = hit ? 1u : 0u;
        if (hit && !any_hit)
            o.t = h.t, o.u = h.u, o.v = h.v, o.prim = h.gid, o.instance = h.inst;
        else
            o.instance = -1;
        o.nodes_visited = (uint32_t) st.nodes, o.tris_tested = (uint32_t) st.tris;
        out[i] = o;
    }
}

// HDRFilm::develop (src/films/hdrfilm.cpp:305-419): RGB / W, W == 0 -> divide by 1
__global__ void develop_kernel(const float4 *__restrict__ rgbw, float *__restrict__ img, size_t n) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float4 v = rgbw[i];
    float w = v.w == 0.f ? 1.f : v.w;
    img[3 * i + 0] = v.x / w;
    img[3 * i + 1] = v.y / w;
    img[3 * i + 2] = v.z / w;
}

// image (+)= RGB / W * scale: one developed render folded into the running multi-pass mean
__global__ void develop_accumulate_kernel(const float4 *__restrict__ rgbw, float *__restrict__ img, size_t n, float scale, int first) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float4 v = rgbw[i];
    float w = v.w == 0.f ? 1.f : v.w;
    float r = v.x / w * scale, g = v.y / w * scale, b = v.z / w * scale;
    if (!first) {
        r += img[3 * i + 0];
        g += img[3 * i + 1];
        b += img[3 * i + 2];
    }
    img[3 * i + 0] = r;
    img[3 * i + 1] = g;
    img[3 * i + 2] = b;
}

// Multi-device contexts: dst += the peers' partial results. The sources are peer-device pointers read over NVLink
// (P2P loads) or staging copies on this device; the summation order is fixed (device 0, 1, 2, ...).
constexpr int kMaxPeers = 15;
struct PeerPtrs {
    const float *p[kMaxPeers];
};
__global__ void peer_reduce_kernel(float *__restrict__ dst, const __grid_constant__ PeerPtrs src, int n_src, size_t n) {
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    const size_t n4 = n / 4;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 a = reinterpret_cast<float4 *>(dst)[i];
        for (int k = 0; k < n_src; ++k) {
            const float4 b = reinterpret_cast<const float4 *>(src.p[k])[i];
            a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        }
        reinterpret_cast<float4 *>(dst)[i] = a;
    }
    for (size_t i = n4 * 4 + (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float a = dst[i];
        for (int k = 0; k < n_src; ++k)
            a += src.p[k][i];
        dst[i] = a;
    }
}

} // namespace

// ==================================================================================================
struct dtof_ctx {
    int device = 0;
    std::string error;
    // scene
    bool has_scene = false;
    dtof_camera cam{};
    dtof_film film{};
    DeviceScene ds{};
    void *d_tris_flat = nullptr, *d_boxes = nullptr;
    size_t flat_bytes = 0, boxes_bytes = 0;
    uint32_t n_tris_total = 0;
    int forced_mode = -1;       // DTOF_MODE env / stats runs: -1 = automatic
    int last_mode = -1;
    void *d_nodes = nullptr, *d_tris = nullptr, *d_shade = nullptr, *d_insts = nullptr, *d_meshes = nullptr,
         *d_bsdfs = nullptr, *d_emitters = nullptr, *d_spots = nullptr, *d_cdf = nullptr, *d_pmf = nullptr;
    std::vector<InstRec> h_insts;
    std::vector<InstBox> h_boxes;    // padded world bounds per instance (host copy: ray binning of the wavefront pipeline)
    std::vector<TlasEntry> h_tlas;   // per instance: motion, object-space bounds, BLAS root (TLAS rebuild on keyframe updates)
    uint32_t tlas_begin = 0, n_tlas_nodes = 0;
    int blas_depth = 0;
    size_t nodes_bytes = 0, tris_bytes = 0, insts_bytes = 0;
    int bvh_depth = 0;
    // render state
    unsigned long long *d_counter = nullptr;
    Counters *d_stats = nullptr;
    float *d_rgbw = nullptr, *d_img = nullptr;
    size_t film_px = 0;
    bool stats_enabled = false;
    dtof_stats last_stats{};
    uint64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t last_stream = nullptr;
    bool have_timing = false;
    int sm_count = 148;
    size_t smem_optin = 0;
    // wavefront pipeline (dtof_wavefront.cuh): queues and per-lane path state for one batch of lanes
    // Two sets, so that two batches are in flight on two streams: the memory-bound shading of one overlaps the
    // issue-bound traversal of the other.
    WfBuffers wf[kWfMaxSets]{};
    void *wf_block[kWfMaxSets] = {};
    size_t wf_cap = 0;
    int wf_sets = 0;
    cudaStream_t wf_stream[kWfMaxSets] = {};
    cudaEvent_t wf_ev_start = nullptr, wf_ev_done[kWfMaxSets] = {};
    uint32_t *wf_host_count = nullptr;   // pinned, one word per set
    int last_pipeline = 0;               // 0 = fused kernel, 1 = wavefront
    // multi-device context (dtof_create_multi): this context is device 0 of the group and owns one full context per
    // further device; the scene is replicated, a render is sharded over all of them and summed here
    std::vector<dtof_ctx *> peers;
    std::vector<int> peer_access;        // per peer: 1 = device 0 reads its film directly over NVLink (P2P), 0 = staged copy
    void *d_stage = nullptr;             // staging film on device 0 for peers without P2P access
    cudaEvent_t ev_done = nullptr;       // "this device's share of the render is complete"
    int last_shard_mode = -1;            // 0 = sample slots, 1 = pixel tiles (multi-device renders)
};

namespace {

dtof_status fail(dtof_ctx *c, dtof_status s, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c)
        c->error = buf;
    return s;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, DTOF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));       \
    } while (0)

void free_scene(dtof_ctx *c) {
    void **ptrs[] = { &c->d_tris_flat, &c->d_boxes, &c->d_nodes, &c->d_tris, &c->d_shade, &c->d_insts, &c->d_meshes, &c->d_bsdfs,
                      &c->d_emitters, &c->d_spots, &c->d_cdf, &c->d_pmf };
    for (void **p : ptrs) {
        if (*p)
            cudaFree(*p);
        *p = nullptr;
    }
    if (c->d_rgbw) cudaFree(c->d_rgbw);
    if (c->d_img) cudaFree(c->d_img);
    c->d_rgbw = c->d_img = nullptr;
    c->has_scene = false;
}

template <typename Tv> dtof_status upload_vec(dtof_ctx *ctx, const std::vector<Tv> &v, void **dst) {
    size_t bytes = std::max<size_t>(v.size() * sizeof(Tv), 16);
    CU(cudaMalloc(dst, bytes));
    CU(cudaMemset(*dst, 0, bytes));
    if (!v.empty())
        CU(cudaMemcpy(*dst, v.data(), v.size() * sizeof(Tv), cudaMemcpyHostToDevice));
    return DTOF_OK;
}

// host copies of the oracle-identical float helpers needed at upload time (rectangle frame, areas)
struct HV3 {
    float x, y, z;
};
inline float hdot(HV3 a, HV3 b) { return std::fmaf(a.z, b.z, std::fmaf(a.y, b.y, a.x * b.x)); }
inline HV3 hcross(HV3 a, HV3 b) {
    return HV3{ std::fmaf(a.y, b.z, -(a.z * b.y)), std::fmaf(a.z, b.x, -(a.x * b.z)), std::fmaf(a.x, b.y, -(a.y * b.x)) };
}

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
    if (spp % spp_per_pass != 0)   // Sampler::set_samples_per_wavefront, src/render/sampler.cpp:81-82
        return 2;
    out->spp_per_pass = spp_per_pass;
    out->n_passes = n_passes;
    out->wavefront_size = wavefront;
    return 0;
}

dtof_status check_params(dtof_ctx *ctx, const dtof_params *p) {
    if (!p)
        return fail(ctx, DTOF_ERR_INVALID, "params is NULL");
    if (p->time_sampling_method > DTOF_TIME_ANTITHETIC_MIRROR)
        return fail(ctx, DTOF_ERR_INVALID, "unknown time_sampling_method %u", p->time_sampling_method);
    if (p->wave_function_type > DTOF_WAVE_TRAPEZOIDAL)
        return fail(ctx, DTOF_ERR_INVALID, "unknown wave_function_type %u", p->wave_function_type);
    if (p->time_correlate_number == 0 || p->path_correlate_number == 0)
        return fail(ctx, DTOF_ERR_INVALID, "correlate numbers must be >= 1");
    if (p->time_sampling_method == DTOF_TIME_ANTITHETIC_MIRROR && p->time_correlate_number != 2)
        return fail(ctx, DTOF_ERR_INVALID, "antithetic_mirror requires time_correlate_number == 2 (correlated.cpp:141-142)");
    if (p->max_depth < -1)
        return fail(ctx, DTOF_ERR_INVALID, "max_depth must be -1 or >= 0");
    if (p->rr_depth <= 0)
        return fail(ctx, DTOF_ERR_INVALID, "rr_depth must be > 0");
    if (p->sample_count == 0)
        return fail(ctx, DTOF_ERR_INVALID, "sample_count must be > 0");
    if (p->use_stratified_sampling_for_each_interval && p->time_sampling_method != DTOF_TIME_UNIFORM &&
        p->sample_count / p->time_correlate_number == 0)
        return fail(ctx, DTOF_ERR_INVALID, "sample_count < time_correlate_number");
    if (!(p->time > 0.f) && p->integrator != DTOF_INTEGRATOR_PATH)   // `path` has no time property
        return fail(ctx, DTOF_ERR_INVALID, "time must be > 0");
    if (p->integrator > DTOF_INTEGRATOR_PATH || p->reserved != 0)
        return fail(ctx, DTOF_ERR_INVALID, "unknown integrator kind %u", p->integrator);
    return DTOF_OK;
}

Modulation make_modulation(const dtof_params &p) {
    Modulation m;
    double mhz = (double) p.w_g;
    m.w_g = (float) (2.0 * M_PI * mhz * 1e6);
    m.w_d = (float) (2.0 * M_PI / (double) p.time * (double) p.hetero_frequency);
    m.k_phi = (float) ((2.0 * M_PI * mhz) / 300.0);
    m.phase = p.sensor_phase_offset;
    m.half_g1 = (float) (0.5 * (double) p.g_1);
    m.g1 = p.g_1;
    m.g0 = p.g_0;
    m.w_gd = m.w_g + m.w_d;
    m.type = p.wave_function_type;
    m.lowpass = p.low_frequency_component_only != 0;
    return m;
}

template <int MODE, bool STATS, bool RECORD, int KIND, bool ENV>
dtof_status launch_variant(dtof_ctx *ctx, RenderArgs &A, int grid, cudaStream_t stream) {
    size_t smem = 0;
    if (MODE == MODE_BVH_SMEM)
        smem = (size_t) A.nodes_bytes + A.tris_bytes + A.insts_bytes + A.boxes_bytes + (size_t) A.stack_levels * 128u * (kBlock / 32);
    else if (MODE == MODE_FLAT_SMEM)
        smem = (size_t) A.tris_bytes + A.insts_bytes + A.boxes_bytes;
    auto k = render_kernel<MODE, STATS, RECORD, KIND, ENV>;
    if (smem)
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    k<<<grid, kBlock, smem, stream>>>(A);
    ctx->launches++;
    CU(cudaGetLastError());
    return DTOF_OK;
}

// ENV = the scene has a constant environment emitter (compiled in only for the BVH modes; launch_render never picks
// the flat walk for such a scene). The velocity integrator does not shade, so it has no ENV instantiation.
template <int MODE, bool ENV>
dtof_status launch_mode(dtof_ctx *ctx, RenderArgs &A, bool record, int grid, cudaStream_t stream) {
    if (A.p.integrator == DTOF_INTEGRATOR_VELOCITY)   // counters are not instrumented for the velocity / path variants
        return record ? launch_variant<MODE, false, true, DTOF_INTEGRATOR_VELOCITY, false>(ctx, A, grid, stream)
                      : launch_variant<MODE, false, false, DTOF_INTEGRATOR_VELOCITY, false>(ctx, A, grid, stream);
    if (A.p.integrator == DTOF_INTEGRATOR_PATH)
        return record ? launch_variant<MODE, false, true, DTOF_INTEGRATOR_PATH, ENV>(ctx, A, grid, stream)
                      : launch_variant<MODE, false, false, DTOF_INTEGRATOR_PATH, ENV>(ctx, A, grid, stream);
    if (record)
        return launch_variant<MODE, false, true, DTOF_INTEGRATOR_DOPPLERTOFPATH, ENV>(ctx, A, grid, stream);
    return ctx->stats_enabled ? launch_variant<MODE, true, false, DTOF_INTEGRATOR_DOPPLERTOFPATH, ENV>(ctx, A, grid, stream)
                              : launch_variant<MODE, false, false, DTOF_INTEGRATOR_DOPPLERTOFPATH, ENV>(ctx, A, grid, stream);
}

// ---- wavefront pipeline (dtof_wavefront.cuh) ------------------------------------------------------------------
constexpr size_t kWfDefaultBatch = 1u << 24;   // lanes per batch: 16 Mi lanes x ~260 B of queues and state = 4.2 GiB per set

void free_wavefront(dtof_ctx *c) {
    for (int i = 0; i < kWfMaxSets; ++i) {
        if (c->wf_block[i])
            cudaFree(c->wf_block[i]);
        c->wf_block[i] = nullptr;
        c->wf[i] = WfBuffers{};
    }
    c->wf_cap = 0;
    c->wf_sets = 0;
}

// One allocation per set, carved into the SoA arrays of WfBuffers (every array 256-byte aligned).
dtof_status ensure_wavefront(dtof_ctx *ctx, size_t cap, int sets) {
    if (!ctx->wf_stream[0]) {
        for (int i = 0; i < kWfMaxSets; ++i) {
            CU(cudaStreamCreateWithFlags(&ctx->wf_stream[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&ctx->wf_ev_done[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&ctx->wf_ev_start, cudaEventDisableTiming));
        CU(cudaMallocHost(&ctx->wf_host_count, 64));
    }
    if (ctx->wf_cap >= cap && ctx->wf_sets >= sets)
        return DTOF_OK;
    free_wavefront(ctx);
    size_t off = 0;
    auto carve = [&](size_t elem) {
        size_t at = off;
        off += (cap * elem + 255) / 256 * 256;
        return at;
    };
    const size_t o_rng = carve(16), o_rngp = carve(16), o_thr = carve(16), o_res = carve(16), o_prev = carve(16),
                 o_film = carve(8), o_qo0 = carve(16), o_qo1 = carve(16), o_qd0 = carve(16), o_qd1 = carve(16),
                 o_ql0 = carve(4), o_ql1 = carve(4), o_hit = carve(16), o_hi = carve(4), o_so = carve(16), o_sd = carve(16),
                 o_st = carve(16), o_sc = carve(16), o_sl = carve(4), o_eta = carve(4);
    const size_t o_ring = off;
    off += (kWfRing * kWfSlotWords * sizeof(uint32_t) + 255) / 256 * 256;
    for (int i = 0; i < sets; ++i) {
        if (cudaMalloc(&ctx->wf_block[i], off) != cudaSuccess) {
            ctx->wf_block[i] = nullptr;
            free_wavefront(ctx);
            return fail(ctx, DTOF_ERR_NOMEM, "cudaMalloc of %zu bytes of wavefront queues failed", off);
        }
        char *b = (char *) ctx->wf_block[i];
        WfBuffers &W = ctx->wf[i];
        W.rng = (ulonglong2 *) (b + o_rng), W.rng_path = (ulonglong2 *) (b + o_rngp);
        W.thr_len = (float4 *) (b + o_thr), W.res_pdf = (float4 *) (b + o_res), W.prev_meta = (float4 *) (b + o_prev);
        W.film_pos = (float2 *) (b + o_film);
        W.eta = (float *) (b + o_eta);
        W.q_o[0] = (float4 *) (b + o_qo0), W.q_o[1] = (float4 *) (b + o_qo1);
        W.q_d[0] = (float4 *) (b + o_qd0), W.q_d[1] = (float4 *) (b + o_qd1);
        W.q_lane[0] = (uint32_t *) (b + o_ql0), W.q_lane[1] = (uint32_t *) (b + o_ql1);
        W.hit = (float4 *) (b + o_hit), W.hit_inst = (int32_t *) (b + o_hi);
        W.s_o = (float4 *) (b + o_so), W.s_d = (float4 *) (b + o_sd), W.s_thr = (float4 *) (b + o_st), W.s_c = (float4 *) (b + o_sc);
        W.s_lane = (uint32_t *) (b + o_sl);
        W.ring = (uint32_t *) (b + o_ring);
    }
    ctx->wf_cap = cap;
    ctx->wf_sets = sets;
    return DTOF_OK;
}

template <int MODE, bool ANY> dtof_status launch_wf_trace(dtof_ctx *ctx, const WfArgs &W, int grid, size_t smem, cudaStream_t stream) {
    auto k = wf_trace_kernel<MODE, ANY>;
    if (smem)
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    k<<<grid, kWfBlock, smem, stream>>>(W);
    ctx->launches++;
    CU(cudaGetLastError());
    return DTOF_OK;
}

// Renders the lanes of `A` (all passes) through the wavefront pipeline. `mode` is MODE_BVH_GLOBAL or MODE_BVH_SMEM.
// Batches alternate between two internal streams (fork from / join into `stream` with events).
dtof_status launch_wavefront(dtof_ctx *ctx, const RenderArgs &A, int mode, cudaStream_t stream) {
    size_t batch = kWfDefaultBatch;
    if (const char *e = getenv("DTOF_WF_BATCH"))
        batch = std::max<size_t>(1024, (size_t) atoll(e));
    uint32_t threshold = 20;   // refill when 12 lanes are idle (re-swept in round 2: 16 / 20 / 24 / 28 -> 825 / 846 / 834 / 806, with binning 868 at 20)
    if (const char *e = getenv("DTOF_WF_THRESHOLD"))
        threshold = (uint32_t) atoi(e);
    int n_streams = kWfMaxSets;
    if (const char *e = getenv("DTOF_WF_STREAMS"))
        n_streams = std::min(std::max(atoi(e), 1), kWfMaxSets);
    size_t cap = (size_t) std::min<unsigned long long>(A.n_local, batch);
    if (cap == 0)
        return DTOF_OK;
    const int want_streams = n_streams;
    dtof_status s;
    for (;;) {   // a GPU that cannot hold 4 x 4.2 GiB of queues next to the scene gets smaller batches
        n_streams = (int) std::min<unsigned long long>((unsigned long long) want_streams, (A.n_local + cap - 1) / cap);
        s = ensure_wavefront(ctx, cap, n_streams);
        if (s != DTOF_ERR_NOMEM || cap <= (1u << 16))
            break;
        cudaGetLastError();   // clear the allocation error
        cap /= 2;
    }
    if (s != DTOF_OK)
        return s;
    WfArgs W{};
    W.scene = A.scene, W.cam = A.cam, W.film = A.film, W.p = A.p, W.mod = A.mod;
    W.spp_per_pass = A.spp_per_pass;
    W.lane_begin = A.lane_begin, W.shard_block = A.shard_block, W.shard_count = A.shard_count, W.shard_index = A.shard_index;
    W.fetch_threshold = threshold;
    W.inner_threshold = 16;
    if (const char *e = getenv("DTOF_WF_INNER"))
        W.inner_threshold = (uint32_t) atoi(e);
    W.double_step = 20;
    if (const char *e = getenv("DTOF_WF_DOUBLE"))
        W.double_step = (uint32_t) atoi(e);
    W.nodes_bytes = A.nodes_bytes, W.tris_bytes = A.tris_bytes, W.insts_bytes = A.insts_bytes;
    // ray binning (dtof_wavefront.cuh): only where the walk is from HBM and a few animated instances carry the geometry
    W.n_cls = 0;
    {
        bool bins = mode == MODE_BVH_GLOBAL;
        if (const char *e = getenv("DTOF_WF_BINS"))   // 0: off, 1: default rule, 2: also for shared-memory scenes (experiments)
            bins = atoi(e) == 2 || (bins && atoi(e) != 0);
        uint32_t n_anim = 0;
        for (const InstRec &r : ctx->h_insts)
            n_anim += r.animated ? 1u : 0u;
        if (bins && n_anim >= 1 && n_anim <= (uint32_t) kWfMaxCls && ctx->h_boxes.size() == ctx->h_insts.size()) {
            for (size_t g = 0; g < ctx->h_insts.size(); ++g) {
                if (!ctx->h_insts[g].animated)
                    continue;
                const InstBox &b = ctx->h_boxes[g];
                W.cls_lo[W.n_cls] = make_float4(b.lox, b.loy, b.loz, 0.f);
                W.cls_hi[W.n_cls] = make_float4(b.hix, b.hiy, b.hiz, 0.f);
                W.n_cls++;
            }
        }
    }
    const bool doppler = A.p.integrator == DTOF_INTEGRATOR_DOPPLERTOFPATH;
    const size_t smem = mode == MODE_BVH_SMEM ? (size_t) A.nodes_bytes + A.tris_bytes + A.insts_bytes : 0;
    // with two batches in flight a traversal kernel leaves room for the other batch's CTAs
    // (measured on the 4.2 M-triangle scene, profiles/r01_tuning.md: 4 batches in flight x 2 CTAs/SM per traversal kernel)
    int trace_ctas = n_streams >= 3 ? 2 : n_streams == 2 ? DTOF_WF_TRACE_CTAS - 1 : DTOF_WF_TRACE_CTAS;
    if (const char *e = getenv("DTOF_WF_TRACE_GRID"))
        trace_ctas = std::max(1, atoi(e));
    const int trace_grid = ctx->sm_count * trace_ctas, shade_grid = ctx->sm_count * DTOF_WF_SHADE_CTAS * (kWfBlock / kWfShadeBlock),
              stream_grid = ctx->sm_count * 8;
    const bool bounded = A.p.max_depth >= 0;
    CU(cudaEventRecord(ctx->wf_ev_start, stream));
    for (int i = 0; i < n_streams; ++i)
        CU(cudaStreamWaitEvent(ctx->wf_stream[i], ctx->wf_ev_start, 0));
    unsigned batch_index = 0;
    for (unsigned long long begin = 0; begin < A.n_local; begin += cap, ++batch_index) {
        const int set = (int) (batch_index % (unsigned) n_streams);
        cudaStream_t st = ctx->wf_stream[set];
        W.buf = ctx->wf[set];
        W.batch_begin = begin;
        W.n_slots = (uint32_t) std::min<unsigned long long>(cap, A.n_local - begin);
        for (uint32_t pass = 0; pass < A.n_passes; ++pass) {
            W.pass = pass;
            W.bounce = 0;
            CU(cudaMemsetAsync(W.buf.ring, 0, kWfRing * kWfSlotWords * sizeof(uint32_t), st));
            {
            NvtxRange r_gen("dtof.wf_generate");
            if (doppler)
                wf_generate_kernel<DTOF_INTEGRATOR_DOPPLERTOFPATH><<<stream_grid, kWfBlock, 0, st>>>(W);
            else
                wf_generate_kernel<DTOF_INTEGRATOR_PATH><<<stream_grid, kWfBlock, 0, st>>>(W);
            }
            ctx->launches++;
            CU(cudaGetLastError());
            for (uint32_t b = 0; !bounded || b < (uint32_t) A.p.max_depth; ++b) {
                W.bounce = b;
                if (b + 1 >= (uint32_t) kWfRing)   // recycle the ring slot the next bounce will count into
                    CU(cudaMemsetAsync(W.buf.ring + kWfSlotWords * ((b + 1) % kWfRing), 0, kWfSlotWords * sizeof(uint32_t), st));
                {
                    NvtxRange r_tc("dtof.wf_trace_closest");
                    s = mode == MODE_BVH_SMEM ? launch_wf_trace<MODE_BVH_SMEM, false>(ctx, W, trace_grid, smem, st)
                                              : launch_wf_trace<MODE_BVH_GLOBAL, false>(ctx, W, trace_grid, 0, st);
                }
                if (s != DTOF_OK)
                    return s;
                nvtxRangePushA("dtof.wf_shade");
                // ENV: environment emitter / non-diffuse BSDFs compiled in only for the scenes that use them
                if (A.scene.extended) {
                    if (doppler)
                        wf_shade_kernel<true, true><<<shade_grid, kWfShadeBlock, 0, st>>>(W);
                    else
                        wf_shade_kernel<false, true><<<shade_grid, kWfShadeBlock, 0, st>>>(W);
                } else {
                    if (doppler)
                        wf_shade_kernel<true, false><<<shade_grid, kWfShadeBlock, 0, st>>>(W);
                    else
                        wf_shade_kernel<false, false><<<shade_grid, kWfShadeBlock, 0, st>>>(W);
                }
                nvtxRangePop();
                ctx->launches++;
                CU(cudaGetLastError());
                {
                    NvtxRange r_ts("dtof.wf_trace_shadow");
                    s = mode == MODE_BVH_SMEM ? launch_wf_trace<MODE_BVH_SMEM, true>(ctx, W, trace_grid, smem, st)
                                              : launch_wf_trace<MODE_BVH_GLOBAL, true>(ctx, W, trace_grid, 0, st);
                }
                if (s != DTOF_OK)
                    return s;
                // bounded depth: every bounce is enqueued blind (an empty queue costs one idle launch). Unbounded or
                // deep paths (Russian roulette ends them): from the 8th bounce on, look at the next queue's length.
                if (b >= 7 && (!bounded || (uint32_t) A.p.max_depth > 8)) {
                    uint32_t *h = ctx->wf_host_count + 2 * set;   // both classes of the next path-ray queue
                    const uint32_t *nx = W.buf.ring + kWfSlotWords * ((b + 1) % kWfRing);
                    CU(cudaMemcpyAsync(h, nx + WF_N_RAY, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                    CU(cudaMemcpyAsync(h + 1, nx + WF_N_RAY_B, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                    CU(cudaStreamSynchronize(st));
                    if (h[0] + h[1] == 0)
                        break;
                }
            }
            {
                NvtxRange r_sp("dtof.wf_splat");
                wf_splat_kernel<<<stream_grid, kWfBlock, 0, st>>>(W);
            }
            ctx->launches++;
            CU(cudaGetLastError());
        }
    }
    for (int i = 0; i < n_streams; ++i) {
        CU(cudaEventRecord(ctx->wf_ev_done[i], ctx->wf_stream[i]));
        CU(cudaStreamWaitEvent(stream, ctx->wf_ev_done[i], 0));
    }
    return DTOF_OK;
}

dtof_status launch_render(dtof_ctx *ctx, const dtof_params *p, float *d_rgbw, cudaStream_t stream,
                          const unsigned long long *d_lanes, dtof_sample_record *d_rec, uint32_t n_rec, uint32_t rec_pass = 0) {
    NvtxRange r_render(d_rec ? "dtof.trace_samples" : "dtof.render");
    dtof_pass_info pi;
    int rc = pass_info(ctx->film, *p, &pi);
    if (rc)
        return fail(ctx, DTOF_ERR_INVALID,
                    rc == 2 ? "sample_count should be a multiple of samples_per_wavefront!" : "invalid sample_count");
    RenderArgs A{};
    A.scene = ctx->ds;
    A.cam = ctx->cam;
    A.p = *p;
    A.mod = make_modulation(*p);
    A.spp_per_pass = pi.spp_per_pass;
    A.n_passes = pi.n_passes;
    A.lane_begin = p->lane_begin;
    A.lane_end = p->lane_end ? p->lane_end : pi.wavefront_size;
    if (A.lane_end > pi.wavefront_size || A.lane_begin > A.lane_end)
        return fail(ctx, DTOF_ERR_INVALID, "lane range [%llu, %llu) outside the wavefront of %llu lanes", A.lane_begin,
                    A.lane_end, (unsigned long long) pi.wavefront_size);
    A.n_local = A.lane_end - A.lane_begin;
    A.shard_block = 0;
    if (p->shard_block) {
        if (p->shard_count == 0 || p->shard_index >= p->shard_count)
            return fail(ctx, DTOF_ERR_INVALID, "shard_index %u outside shard_count %u", p->shard_index, p->shard_count);
        unsigned long long total = A.lane_end - A.lane_begin, B = p->shard_block, R = p->shard_count, r = p->shard_index;
        unsigned long long full_blocks = total / B, tail = total % B;
        unsigned long long mine = full_blocks / R + (r < full_blocks % R ? 1 : 0);
        A.n_local = mine * B + ((tail && full_blocks % R == r) ? tail : 0);
        A.shard_block = B;
        A.shard_count = p->shard_count;
        A.shard_index = p->shard_index;
    }
    FilmParams &F = A.film;
    F.rgbw = d_rgbw;
    F.width = ctx->film.width, F.height = ctx->film.height;
    F.crop_x = ctx->film.crop_offset_x, F.crop_y = ctx->film.crop_offset_y;
    F.rfilter = ctx->film.rfilter;
    F.radius = ctx->film.rfilter_radius;
    F.inv_radius = 1.f / F.radius;
    F.g_alpha = -1.f / (2.f * ctx->film.gaussian_stddev * ctx->film.gaussian_stddev);
    F.g_bias = expf(F.g_alpha * F.radius * F.radius);
    F.n = (int) ceilf(F.radius - .5f);
    F.mitchell_b = ctx->film.mitchell_b, F.mitchell_c = ctx->film.mitchell_c;
    A.work_counter = ctx->d_counter;
    A.stats = ctx->d_stats;
    A.rec_lanes = d_lanes;
    A.rec_out = d_rec;
    A.n_rec = n_rec;
    A.tris_flat = (const float4 *) ctx->d_tris_flat;
    A.inst_box = (const float4 *) ctx->d_boxes;
    const bool record = d_rec != nullptr;
    // ---- traversal mode: flat coherent walk for tiny scenes, BVH in shared memory while it fits, else BVH from HBM
    A.stack_levels = (uint32_t) ctx->bvh_depth + 4u;
#ifdef DTOF_LOCAL_STACK
    A.stack_levels = 0;   // A/B builds: no shared-memory stack rows
#endif
    A.rec_pass = rec_pass;
    const size_t stack_bytes = (size_t) A.stack_levels * 128u * (kBlock / 32);   // shared-memory traversal stacks of one CTA
    const size_t bvh_bytes = ctx->nodes_bytes + ctx->tris_bytes + ctx->insts_bytes + ctx->boxes_bytes + stack_bytes;
    const size_t flat_bytes = ctx->flat_bytes + ctx->insts_bytes + ctx->boxes_bytes;
    int mode = MODE_BVH_GLOBAL;
    if (bvh_bytes <= kSmemSceneLimit && bvh_bytes + 1024 <= ctx->smem_optin)
        mode = MODE_BVH_SMEM;
    if (ctx->n_tris_total <= kFlatMaxTris && ctx->n_tris_total > 0 && flat_bytes + 1024 <= ctx->smem_optin)
        mode = MODE_FLAT_SMEM;
    int want = ctx->forced_mode;
    if (const char *e = getenv("DTOF_MODE"))
        want = atoi(e);
    if (want == MODE_BVH_GLOBAL || (want == MODE_BVH_SMEM && bvh_bytes + 1024 <= ctx->smem_optin) ||
        (want == MODE_FLAT_SMEM && flat_bytes + 1024 <= ctx->smem_optin))
        mode = want;
    const bool env = ctx->ds.extended != 0;   // environment emitter or conductor: the ENV = true instantiations
    if (env && mode == MODE_FLAT_SMEM)   // the flat walk has no such instantiation
        mode = bvh_bytes + 1024 <= ctx->smem_optin ? MODE_BVH_SMEM : MODE_BVH_GLOBAL;
    // the traversal counters are defined on the BVH walk (algorithmic work of the scene), whatever
    // mode the production launch of this scene picks
    if (ctx->stats_enabled && !record && mode == MODE_FLAT_SMEM)
        mode = bvh_bytes + 1024 <= ctx->smem_optin ? MODE_BVH_SMEM : MODE_BVH_GLOBAL;
    A.nodes_bytes = (uint32_t) ctx->nodes_bytes;
    A.tris_bytes = (uint32_t) (mode == MODE_FLAT_SMEM ? ctx->flat_bytes : ctx->tris_bytes);
    A.insts_bytes = (uint32_t) ctx->insts_bytes;
    A.boxes_bytes = (uint32_t) ctx->boxes_bytes;
    unsigned long long n_lanes = record ? n_rec : A.n_local;
    unsigned long long n_units = (n_lanes + kUnit - 1) / kUnit;
    unsigned long long want_ctas = (n_units + (kBlock / 32) - 1) / (kBlock / 32);
    int grid = (int) std::min<unsigned long long>(std::max<unsigned long long>(want_ctas, 1), (unsigned long long) ctx->sm_count * DTOF_MIN_CTAS);
    CU(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), stream));
    if (ctx->stats_enabled)
        CU(cudaMemsetAsync(ctx->d_stats, 0, sizeof(Counters), stream));
    CU(cudaEventRecord(ctx->ev0, stream));
    dtof_status s;
    // ---- pipeline: the wavefront pipeline with dynamic ray fetch where the BVH is walked from HBM (divergent
    // traversal lengths), the fused kernel for shared-memory scenes, record / stats runs and the velocity integrator.
    // DTOF_WAVEFRONT=0 / 1 forces the choice (1 also for shared-memory scenes).
    int wavefront = mode == MODE_BVH_GLOBAL ? 1 : 0;
    if (const char *e = getenv("DTOF_WAVEFRONT"))
        wavefront = atoi(e);
    if (record || ctx->stats_enabled || p->integrator == DTOF_INTEGRATOR_VELOCITY || mode == MODE_FLAT_SMEM ||
        !ctx->ds.has_geometry)
        wavefront = 0;
    ctx->last_pipeline = wavefront;
    if (wavefront) {
        if ((s = launch_wavefront(ctx, A, mode, stream)) != DTOF_OK)
            return s;
        ctx->last_mode = mode;
        CU(cudaEventRecord(ctx->ev1, stream));
        ctx->last_stream = stream;
        ctx->have_timing = true;
        return DTOF_OK;
    }
    if (mode == MODE_FLAT_SMEM)
        s = launch_mode<MODE_FLAT_SMEM, false>(ctx, A, record, grid, stream);
    else if (mode == MODE_BVH_SMEM)
        s = env ? launch_mode<MODE_BVH_SMEM, true>(ctx, A, record, grid, stream)
                : launch_mode<MODE_BVH_SMEM, false>(ctx, A, record, grid, stream);
    else
        s = env ? launch_mode<MODE_BVH_GLOBAL, true>(ctx, A, record, grid, stream)
                : launch_mode<MODE_BVH_GLOBAL, false>(ctx, A, record, grid, stream);
    if (s != DTOF_OK)
        return s;
    ctx->last_mode = mode;
    CU(cudaEventRecord(ctx->ev1, stream));
    ctx->last_stream = stream;
    ctx->have_timing = true;
    return DTOF_OK;
}

// ---- multi-device contexts (dtof_create_multi) -------------------------------------------------------------------------
// dst (on device 0, `n_floats` floats) += the same buffer of every peer, after the peer has signalled ev_done. Peers the
// device can address are read in place over NVLink; the others are copied into a staging buffer first.
dtof_status reduce_from_peers(dtof_ctx *ctx, float *dst, size_t n_floats, bool image, cudaStream_t stream) {
    NvtxRange r_reduce("dtof.film_reduce_peers");
    CU(cudaSetDevice(ctx->device));
    PeerPtrs src{};
    size_t n_staged = 0;
    for (size_t k = 0; k < ctx->peers.size(); ++k)
        n_staged += ctx->peer_access[k] ? 0 : 1;
    if (n_staged && !ctx->d_stage)
        CU(cudaMalloc(&ctx->d_stage, n_staged * ctx->film_px * 4 * sizeof(float)));
    size_t staged = 0;
    for (size_t k = 0; k < ctx->peers.size(); ++k) {
        dtof_ctx *p = ctx->peers[k];
        const float *buf = image ? p->d_img : p->d_rgbw;
        CU(cudaStreamWaitEvent(stream, p->ev_done, 0));
        if (ctx->peer_access[k]) {
            src.p[k] = buf;
        } else {
            float *st = (float *) ctx->d_stage + staged++ * ctx->film_px * 4;
            CU(cudaMemcpyPeerAsync(st, ctx->device, buf, p->device, n_floats * sizeof(float), stream));
            src.p[k] = st;
        }
    }
    const int grid = (int) std::min<size_t>((n_floats / 4 + 255) / 256 + 1, (size_t) ctx->sm_count * 8);
    peer_reduce_kernel<<<grid, 256, 0, stream>>>(dst, src, (int) ctx->peers.size(), n_floats);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_done, stream));   // on device 0: "the peers' buffers have been read"
    return DTOF_OK;
}

// Runs job(i, context of device i) for every device of the group: device 0 on the calling thread, the peers on one
// thread each (enqueueing a render is thousands of launches for the wavefront pipeline; the devices must not wait for
// each other's host work). Every job starts by waiting for device 0's last reduce, which read the peer's buffers.
template <typename Job> dtof_status for_each_device(dtof_ctx *ctx, Job job) {
    const size_t n = 1 + ctx->peers.size();
    std::vector<dtof_status> st(n, DTOF_OK);
    auto run = [&](size_t i) {
        dtof_ctx *c = i ? ctx->peers[i - 1] : ctx;
        if (cudaSetDevice(c->device) != cudaSuccess) {
            st[i] = DTOF_ERR_CUDA;
            return;
        }
        if (i)
            cudaStreamWaitEvent(0, ctx->ev_done, 0);
        st[i] = job(i, c);
        if (i && st[i] == DTOF_OK && cudaEventRecord(c->ev_done, 0) != cudaSuccess)
            st[i] = DTOF_ERR_CUDA;
    };
    std::vector<std::thread> th;
    for (size_t i = 1; i < n; ++i)
        th.emplace_back(run, i);
    run(0);
    for (auto &t : th)
        t.join();
    cudaSetDevice(ctx->device);
    for (size_t i = 0; i < n; ++i)
        if (st[i] != DTOF_OK)
            return i ? fail(ctx, st[i], "device %d: %s", ctx->peers[i - 1]->device, ctx->peers[i - 1]->error.c_str()) : st[i];
    return DTOF_OK;
}

// ONE render sharded over the devices of the group: by sample slots -- every device renders
// spp_per_pass / n slots of every pixel, whole correlate groups -- when spp_per_pass divides that way, else by
// interleaved 64-pixel tiles. All passes of a lane stay on one device. Device 0 accumulates into `d_rgbw0` (zeroed first if
// `zero0`), the peers into their own films; then device 0 adds the peers' films to `d_rgbw0`. Asynchronous.
dtof_status render_sharded(dtof_ctx *ctx, const dtof_params *params, float *d_rgbw0, bool zero0, cudaStream_t stream0) {
    const uint32_t n = 1u + (uint32_t) ctx->peers.size();
    if (params->shard_block)
        return fail(ctx, DTOF_ERR_INVALID, "a multi-device context shards the render itself: shard_block must be 0");
    dtof_pass_info pi;
    if (pass_info(ctx->film, *params, &pi))
        return fail(ctx, DTOF_ERR_INVALID, "sample_count should be a multiple of samples_per_wavefront!");
    uint64_t a = params->time_correlate_number, b = params->path_correlate_number, g = a, h = b;
    while (h) {
        const uint64_t t = g % h;
        g = h, h = t;
    }
    const uint64_t group = a / g * b;   // lcm(tcn, pcn): antithetic / correlated groups stay on one device
    dtof_params base = *params;
    base.shard_count = n;
    if (pi.spp_per_pass % (n * group) == 0) {
        base.shard_block = pi.spp_per_pass / n;
        ctx->last_shard_mode = 0;
    } else {
        base.shard_block = (uint64_t) pi.spp_per_pass * 64u;
        ctx->last_shard_mode = 1;
    }
    dtof_status s = for_each_device(ctx, [&](size_t i, dtof_ctx *c) -> dtof_status {
        dtof_ctx *ctx = c;   // for CU()
        dtof_params p = base;
        p.shard_index = (uint32_t) i;
        float *film = i ? c->d_rgbw : d_rgbw0;
        cudaStream_t st = i ? (cudaStream_t) 0 : stream0;
        if (i || zero0)
            CU(cudaMemsetAsync(film, 0, c->film_px * 4 * sizeof(float), st));
        return launch_render(c, &p, film, st, nullptr, nullptr, 0);
    });
    if (s != DTOF_OK)
        return s;
    if ((s = reduce_from_peers(ctx, d_rgbw0, ctx->film_px * 4, false, stream0)) != DTOF_OK)
        return s;
    CU(cudaEventRecord(ctx->ev1, stream0));   // dtof_last_kernel_ms: device 0's share + waiting for the peers + the reduce
    return DTOF_OK;
}

} // namespace

namespace {

// Host half of dtof_upload_scene: validate, flatten, build the two-level BVH. No CUDA call in here.
struct HostScene {
    std::vector<MeshRec> meshes;
    std::vector<BsdfRec> bsdfs;
    std::vector<EmitterRec> emitters;
    std::vector<SpotRec> spots;   // empty unless the scene has a spot light
    std::vector<TriShade> shade;
    std::vector<float> cdf, pmf;
    std::vector<InstRec> insts;
    std::vector<TriIsect> tris_flat;
    BuiltScene built;
    uint32_t n_tris = 0;
    // constant environment emitter: index (-1: none) and the scene's bounding sphere (constant.cpp:73-82)
    int32_t env_emitter = -1;
    float env_center[3] = { 0.f, 0.f, 0.f }, env_radius = 1.f;
    bool extended = false;   // the scene needs the ENV = true kernel instantiations (environment emitter or conductor)
};

dtof_status prepare_scene(dtof_ctx *ctx, const dtof_scene_desc *sc, HostScene &H) {
    if (sc->film.width == 0 || sc->film.height == 0)
        return fail(ctx, DTOF_ERR_INVALID, "empty film");
    if (sc->film.rfilter > DTOF_RFILTER_LANCZOS)
        return fail(ctx, DTOF_ERR_INVALID, "unknown rfilter %u", sc->film.rfilter);
    if (sc->film.rfilter != DTOF_RFILTER_BOX && !(sc->film.rfilter_radius > 0.f && sc->film.rfilter_radius <= 7.5f))
        return fail(ctx, DTOF_ERR_INVALID, "rfilter radius out of range");
    H.extended = sc->film.rfilter >= DTOF_RFILTER_MITCHELL;   // filters with negative lobes live in the ENV = true instantiations
    // ---- validate + flatten
    std::vector<MeshRec> &meshes = H.meshes;
    std::vector<BsdfRec> &bsdfs = H.bsdfs;
    std::vector<EmitterRec> &emitters = H.emitters;
    std::vector<TriShade> &shade = H.shade;
    std::vector<float> &cdf = H.cdf, &pmf = H.pmf;
    meshes.assign(sc->n_meshes, MeshRec{});
    bsdfs.assign(sc->n_bsdfs, BsdfRec{});
    emitters.assign(sc->n_emitters, EmitterRec{});
    std::vector<GroupInput> groups(sc->n_instances);
    std::vector<uint32_t> mesh_first_gid(sc->n_meshes, 0xffffffffu);
    for (uint32_t i = 0; i < sc->n_bsdfs; ++i) {
        const dtof_bsdf &b = sc->bsdfs[i];
        if (b.kind > DTOF_BSDF_ROUGHDIELECTRIC)
            return fail(ctx, DTOF_ERR_UNSUPPORTED, "bsdf kind %u is outside the hot-path scope", b.kind);
        const bool dielectric = b.kind == DTOF_BSDF_DIELECTRIC || b.kind == DTOF_BSDF_THINDIELECTRIC;
        if (b.kind == DTOF_BSDF_ROUGHDIELECTRIC && (b.twosided || !(b.eta[0] > 0.f) || b.eta[0] == 1.f))
            return fail(ctx, DTOF_ERR_INVALID, b.twosided ? "Only materials without a transmission component can be nested!"
                                                          : "The interior and exterior indices of refraction must be positive and differ!");
        if (dielectric && b.twosided)
            return fail(ctx, DTOF_ERR_INVALID, "Only materials without a transmission component can be nested!");   // twosided.cpp:102-103
        if (dielectric && !(b.eta[0] > 0.f))
            return fail(ctx, DTOF_ERR_INVALID, "The interior and exterior indices of refraction must be positive!");
        bsdfs[i] = BsdfRec{ b.reflectance[0], b.reflectance[1], b.reflectance[2],
                            (b.twosided ? 1u : 0u) | (b.kind == DTOF_BSDF_DIFFUSE ? 2u : 0u) |
                                (b.kind == DTOF_BSDF_CONDUCTOR ? 4u : 0u) | (dielectric ? 8u : 0u) |
                                (b.kind == DTOF_BSDF_THINDIELECTRIC ? 16u : 0u),
                            b.eta[0], b.eta[1], b.eta[2], 0.f, b.k[0], b.k[1], b.k[2], 0.f };
        H.extended = H.extended || b.kind == DTOF_BSDF_CONDUCTOR || dielectric || b.kind == DTOF_BSDF_PLASTIC;
        if (b.kind == DTOF_BSDF_ROUGHDIELECTRIC) {
            if (b.distribution > 1u)
                return fail(ctx, DTOF_ERR_INVALID, "Specified an invalid distribution, must be \"beckmann\" or \"ggx\"!");
            BsdfRec &r = bsdfs[i];
            r.flags |= 2u | 256u | (b.distribution == 1u ? 128u : 0u);
            r.pad0 = std::max(b.alpha[0], 1e-4f), r.pad1 = std::max(b.alpha[1], 1e-4f);
            H.extended = true;
        }
        if (b.kind == DTOF_BSDF_ROUGHCONDUCTOR) {   // MicrofacetDistribution::configure: alpha >= 1e-4 (microfacet.h:420-423)
            if (b.distribution > 1u)
                return fail(ctx, DTOF_ERR_INVALID, "Specified an invalid distribution, must be \"beckmann\" or \"ggx\"!");
            BsdfRec &r = bsdfs[i];
            r.flags |= 2u | 64u | (b.distribution == 1u ? 128u : 0u);
            r.pad0 = std::max(b.alpha[0], 1e-4f), r.pad1 = std::max(b.alpha[1], 1e-4f);
            H.extended = true;
        }
        if (b.kind == DTOF_BSDF_PLASTIC) {   // SmoothPlastic::parameters_changed, plastic.cpp:193-208
            if (!(b.eta[0] > 0.f))
                return fail(ctx, DTOF_ERR_INVALID, "The interior and exterior indices of refraction must be positive!");
            const float eta = b.eta[0], inv_eta = 1.f / (1.f / eta);   // fresnel_diffuse_reflectance(1 / eta): its inv_eta
            const float e = 1.f / eta;                                  // the argument
            float fdr;
            if (e < 1.f) {
                fdr = std::fmaf(0.0636f, inv_eta, std::fmaf(e, std::fmaf(e, -1.4399f, 0.7099f), 0.6681f));
            } else {
                const float c[6] = { 0.919317f, -3.4793f, 6.75335f, -7.80989f, 4.98554f, -1.36881f };
                fdr = c[5];
                for (int j = 4; j >= 0; --j)
                    fdr = std::fmaf(inv_eta, fdr, c[j]);
            }
            const float d_mean = (b.reflectance[0] + b.reflectance[1] + b.reflectance[2]) * (1.f / 3.f);
            const float s_mean = (b.k[0] + b.k[1] + b.k[2]) * (1.f / 3.f);
            BsdfRec &r = bsdfs[i];
            r.flags |= 2u | 32u;
            r.eta_r = eta, r.eta_g = fdr, r.eta_b = 1.f / (eta * eta), r.pad0 = s_mean / (d_mean + s_mean);
            r.pad1 = b.eta[1] != 0.f ? 1.f : 0.f;
        }
    }
    uint32_t gid = 0;
    for (uint32_t g = 0; g < sc->n_instances; ++g) {
        const dtof_instance &in = sc->instances[g];
        if ((uint64_t) in.first_mesh + in.n_meshes > sc->n_meshes)
            return fail(ctx, DTOF_ERR_INVALID, "instance %u references meshes out of range", g);
        if (in.animated && !(in.t1 > in.t0))
            return fail(ctx, DTOF_ERR_INVALID, "instance %u: keyframe times must be increasing", g);
        GroupInput &G = groups[g];
        G.animated = in.animated != 0;
        memcpy(G.m0, in.m0, sizeof(G.m0));
        memcpy(G.m1, in.m1, sizeof(G.m1));
        for (uint32_t mi = in.first_mesh; mi < in.first_mesh + in.n_meshes; ++mi) {
            const dtof_mesh &m = sc->meshes[mi];
            if (mesh_first_gid[mi] != 0xffffffffu)
                return fail(ctx, DTOF_ERR_UNSUPPORTED, "mesh %u belongs to more than one instance", mi);
            if (!m.positions || !m.faces)
                return fail(ctx, DTOF_ERR_INVALID, "mesh %u has no positions/faces", mi);
            if (m.bsdf >= sc->n_bsdfs)
                return fail(ctx, DTOF_ERR_INVALID, "mesh %u: bsdf index out of range", mi);
            if (m.emitter >= (int32_t) sc->n_emitters)
                return fail(ctx, DTOF_ERR_INVALID, "mesh %u: emitter index out of range", mi);
            if (m.emitter >= 0 && in.animated)
                return fail(ctx, DTOF_ERR_UNSUPPORTED, "Instancing of emitters is not supported (shapegroup.cpp:27-30)");
            mesh_first_gid[mi] = gid;
            uint32_t flags = (m.normals ? TRI_HAS_NORMALS : 0u) | (m.texcoords ? TRI_HAS_UV : 0u) | (m.flip_normals ? TRI_FLIP : 0u);
            for (uint32_t f = 0; f < m.n_faces; ++f, ++gid) {
                uint32_t i0 = m.faces[3 * f], i1 = m.faces[3 * f + 1], i2 = m.faces[3 * f + 2];
                if (i0 >= m.n_vertices || i1 >= m.n_vertices || i2 >= m.n_vertices)
                    return fail(ctx, DTOF_ERR_INVALID, "mesh %u face %u: vertex index out of range", mi, f);
                const float *p0 = m.positions + 3 * i0, *p1 = m.positions + 3 * i1, *p2 = m.positions + 3 * i2;
                TriIsect t{};
                t.p0x = p0[0], t.p0y = p0[1], t.p0z = p0[2];
                t.gid = gid;
                t.e1x = p1[0] - p0[0], t.e1y = p1[1] - p0[1], t.e1z = p1[2] - p0[2];
                t.e2x = p2[0] - p0[0], t.e2y = p2[1] - p0[1], t.e2z = p2[2] - p0[2];
                G.tris.push_back(t);
                G.p1.insert(G.p1.end(), p1, p1 + 3);
                G.p2.insert(G.p2.end(), p2, p2 + 3);
                TriShade s{};
                s.p0x = p0[0], s.p0y = p0[1], s.p0z = p0[2];
                s.mesh = mi;
                s.p1x = p1[0], s.p1y = p1[1], s.p1z = p1[2];
                s.flags = flags;
                s.p2x = p2[0], s.p2y = p2[1], s.p2z = p2[2];
                if (m.normals) {
                    const float *n0 = m.normals + 3 * i0, *n1 = m.normals + 3 * i1, *n2 = m.normals + 3 * i2;
                    s.n0x = n0[0], s.n0y = n0[1], s.n0z = n0[2];
                    s.n1x = n1[0], s.n1y = n1[1], s.n1z = n1[2];
                    s.n2x = n2[0], s.n2y = n2[1], s.n2z = n2[2];
                }
                if (m.texcoords) {
                    const float *u0 = m.texcoords + 2 * i0, *u1 = m.texcoords + 2 * i1, *u2 = m.texcoords + 2 * i2;
                    s.uv0x = u0[0], s.uv0y = u0[1], s.uv1x = u1[0], s.uv1y = u1[1], s.uv2x = u2[0], s.uv2y = u2[1];
                }
                shade.push_back(s);
            }
        }
    }
    if ((uint64_t) gid >= (1ull << 27))
        return fail(ctx, DTOF_ERR_UNSUPPORTED, "more than 2^27 triangles");
    for (uint32_t mi = 0; mi < sc->n_meshes; ++mi) {
        const dtof_mesh &m = sc->meshes[mi];
        MeshRec &r = meshes[mi];
        memset(&r, 0, sizeof(r));
        r.bsdf = m.bsdf;
        r.emitter = m.emitter;
        r.kind = m.kind;
        r.n_faces = m.n_faces;
        r.first_gid = mesh_first_gid[mi];
        if (m.emitter >= 0 && mesh_first_gid[mi] == 0xffffffffu)
            return fail(ctx, DTOF_ERR_INVALID, "emitter mesh %u is not part of any instance", mi);
        if (m.kind == DTOF_SHAPE_RECTANGLE) {   // Rectangle::update, src/shapes/rectangle.cpp:101-113
            memcpy(r.rect_to_world, m.rect_to_world, sizeof(r.rect_to_world));
            const float *M = m.rect_to_world;
            HV3 s{ M[0] * 2.f, M[4] * 2.f, M[8] * 2.f }, t{ M[1] * 2.f, M[5] * 2.f, M[9] * 2.f };
            // normal = normalize(inverse_transpose * (0,0,1)) = normalize(third row of the cofactor matrix)
            float c02 = std::fmaf(M[4], M[9], -(M[5] * M[8])), c12 = std::fmaf(M[1], M[8], -(M[0] * M[9])),
                  c22 = std::fmaf(M[0], M[5], -(M[1] * M[4]));
            float c00 = std::fmaf(M[5], M[10], -(M[6] * M[9])), c01 = std::fmaf(M[6], M[8], -(M[4] * M[10]));
            float det = std::fmaf(M[0], c00, std::fmaf(M[1], c01, M[2] * c02));
            float id = 1.f / det;
            HV3 nn{ c02 * id, c12 * id, c22 * id };
            float il = 1.f / std::sqrt(hdot(nn, nn));
            r.rect_n[0] = nn.x * il, r.rect_n[1] = nn.y * il, r.rect_n[2] = nn.z * il;
            HV3 c = hcross(s, t);
            r.inv_area = 1.f / std::sqrt(hdot(c, c));
        } else if (m.emitter >= 0) {   // Mesh::build_pmf (mesh.cpp:361-393) + DiscreteDistribution::compute_cdf
            r.cdf_offset = (uint32_t) cdf.size();
            double sum = 0.0;
            bool any = false;
            for (uint32_t f = 0; f < m.n_faces; ++f) {
                const float *p0 = m.positions + 3 * m.faces[3 * f], *p1 = m.positions + 3 * m.faces[3 * f + 1],
                            *p2 = m.positions + 3 * m.faces[3 * f + 2];
                HV3 a{ p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2] }, b{ p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2] };
                HV3 c = hcross(a, b);
                float area = .5f * std::sqrt(hdot(c, c));
                pmf.push_back(area);
                sum += (double) area;
                cdf.push_back((float) sum);
                if (area > 0.f) {
                    if (!any)
                        r.valid_lo = f;
                    r.valid_hi = f;
                    any = true;
                }
            }
            if (!any)
                return fail(ctx, DTOF_ERR_INVALID, "emitter mesh %u has no surface area", mi);
            r.area_sum = (float) sum;
            r.inv_area = (float) (1.0 / sum);
        }
    }
    bool all_point = true, need_bsphere = false;
    for (uint32_t i = 0; i < sc->n_emitters; ++i) {
        const dtof_emitter &e = sc->emitters[i];
        if (e.kind > DTOF_EMITTER_DIRECTIONAL)
            return fail(ctx, DTOF_ERR_UNSUPPORTED, "emitter kind %u is outside the hot-path scope", e.kind);
        if (e.kind == DTOF_EMITTER_SPOT) {   // SpotLight ctor, spot.cpp:102-112
            if (!(e.cutoff_angle >= e.beam_width) || !(e.cutoff_angle > 0.f))
                return fail(ctx, DTOF_ERR_INVALID, "spot emitter %u: cutoff_angle must be positive and >= beam_width", i);
            if (H.spots.empty())
                H.spots.assign(sc->n_emitters, SpotRec{});
            SpotRec &sr = H.spots[i];
            memcpy(sr.to_local, e.to_local, sizeof(sr.to_local));
            sr.cutoff_angle = e.cutoff_angle;
            sr.cos_cutoff = std::cos(e.cutoff_angle), sr.cos_beam = std::cos(e.beam_width);
            sr.inv_transition = 1.f / (e.cutoff_angle - e.beam_width);
            H.extended = true;
        }
        if (e.kind == DTOF_EMITTER_DIRECTIONAL)   // distant light: needs the scene's bounding sphere like the environment emitter
            H.extended = need_bsphere = true;
        if (e.kind == DTOF_EMITTER_CONSTANT) {
            need_bsphere = true;
            if (H.env_emitter >= 0)
                return fail(ctx, DTOF_ERR_INVALID, "Only one environment emitter can be specified per scene.");   // scene.cpp:53-55
            H.env_emitter = (int32_t) i;
            H.extended = true;
        }
        if (e.kind == DTOF_EMITTER_AREA && (e.mesh >= sc->n_meshes || sc->meshes[e.mesh].emitter != (int32_t) i))
            return fail(ctx, DTOF_ERR_INVALID, "area emitter %u and its mesh do not reference each other", i);
        all_point = all_point && e.kind == DTOF_EMITTER_POINT;
        emitters[i] = EmitterRec{ e.kind, e.mesh, e.position[0], e.position[1], e.position[2], e.value[0], e.value[1], e.value[2] };
    }
    (void) all_point;

    // ---- BVH
    BuiltScene &built = H.built;
    build_scene_bvh(groups, built);
    if (built.max_depth + built.tlas_depth + 4 > kStackSize)
        return fail(ctx, DTOF_ERR_UNSUPPORTED, "BVH too deep for the traversal stack (%d + %d)", built.max_depth, built.tlas_depth);
    if (need_bsphere) {
        // Scene::bbox (scene.cpp:36): union of the shapes' boxes -- static shapes: their vertices; an instance: the 8
        // corners of its group's box under both keyframes (instance.cpp:101-114). Then ConstantBackgroundEmitter::
        // set_scene: bounding sphere, radius * (1 + RayEpsilon), at least RayEpsilon; an empty scene gives (0, 1).
        float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
        auto grow = [&](const float *q) {
            for (int a = 0; a < 3; ++a)
                lo[a] = std::min(lo[a], q[a]), hi[a] = std::max(hi[a], q[a]);
        };
        for (uint32_t g = 0; g < sc->n_instances; ++g) {
            const InstBox &ob = built.group_box[g];
            if (!(ob.lox <= ob.hix))
                continue;
            for (int c = 0; c < 8; ++c) {
                const float p[3] = { (c & 1) ? ob.hix : ob.lox, (c & 2) ? ob.hiy : ob.loy, (c & 4) ? ob.hiz : ob.loz };
                if (!sc->instances[g].animated) {
                    grow(p);
                    continue;
                }
                for (const float *M : { sc->instances[g].m0, sc->instances[g].m1 }) {   // Transform * Point, column by column
                    float q[3];
                    for (int r = 0; r < 3; ++r)
                        q[r] = std::fmaf(M[4 * r + 3], 1.f, std::fmaf(M[4 * r + 2], p[2], std::fmaf(M[4 * r + 1], p[1], M[4 * r] * p[0])));
                    grow(q);
                }
            }
        }
        if (lo[0] <= hi[0]) {
            float c[3], d2 = 0.f;
            for (int a = 0; a < 3; ++a) {
                c[a] = (lo[a] + hi[a]) * .5f;
                H.env_center[a] = c[a];
            }
            HV3 dv{ c[0] - hi[0], c[1] - hi[1], c[2] - hi[2] };
            d2 = hdot(dv, dv);
            const float eps = 1500.f * 5.9604644775390625e-08f;   // math::RayEpsilon<float>
            H.env_radius = std::max(eps, std::sqrt(d2) * (1.f + eps));
        }
    }
    std::vector<InstRec> &insts = H.insts;
    insts.assign(sc->n_instances, InstRec{});
    std::vector<TriIsect> &tris_flat = H.tris_flat;   // scene (gid) order, for the flat traversal of tiny scenes
    for (uint32_t g = 0; g < sc->n_instances; ++g) {
        const dtof_instance &in = sc->instances[g];
        InstRec &r = insts[g];
        memcpy(r.m0, in.m0, sizeof(r.m0));
        memcpy(r.m1, in.m1, sizeof(r.m1));
        r.t0 = in.t0, r.t1 = in.t1;
        r.root = built.inst_root[g];
        r.animated = in.animated ? 1u : 0u;
        r.first_tri = (uint32_t) tris_flat.size();
        r.n_tris = (uint32_t) groups[g].tris.size();
        r.pad0 = r.pad1 = 0;
        if (gid <= 4096)               // only tiny scenes ever use the flat walk
            tris_flat.insert(tris_flat.end(), groups[g].tris.begin(), groups[g].tris.end());
    }

    H.n_tris = gid;
    return DTOF_OK;
}

} // namespace

// ==================================================================================================
extern "C" {

uint32_t dtof_abi_version(void) { return DTOF_ABI_VERSION; }

dtof_status dtof_create_multi(dtof_ctx **out, const int *devices, uint32_t n_devices) {
    if (!out || !devices || n_devices == 0 || n_devices > (uint32_t) kMaxPeers + 1)
        return DTOF_ERR_INVALID;
    *out = nullptr;
    for (uint32_t i = 0; i < n_devices; ++i)
        for (uint32_t j = 0; j < i; ++j)
            if (devices[i] == devices[j])
                return DTOF_ERR_INVALID;
    dtof_ctx *ctx = nullptr;
    dtof_status s = dtof_create(&ctx, devices[0]);
    if (s != DTOF_OK)
        return s;
    for (uint32_t i = 1; i < n_devices; ++i) {
        dtof_ctx *p = nullptr;
        if ((s = dtof_create(&p, devices[i])) != DTOF_OK) {
            dtof_destroy(ctx);
            return s;
        }
        ctx->peers.push_back(p);
        // device 0 reads the peer's film directly when the two can address each other (NVLink / NVSwitch)
        int can = 0;
        cudaSetDevice(devices[0]);
        if (cudaDeviceCanAccessPeer(&can, devices[0], devices[i]) == cudaSuccess && can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[i], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                can = 0;
            cudaGetLastError();
        }
        ctx->peer_access.push_back(can ? 1 : 0);
    }
    cudaSetDevice(devices[0]);
    *out = ctx;
    return DTOF_OK;
}

uint32_t dtof_device_count(const dtof_ctx *ctx) { return ctx ? 1u + (uint32_t) ctx->peers.size() : 0u; }

dtof_status dtof_create(dtof_ctx **out, int device) {
    if (!out)
        return DTOF_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n)
        return e != cudaSuccess ? DTOF_ERR_CUDA : DTOF_ERR_INVALID;
    dtof_ctx *ctx = new dtof_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaMalloc(&ctx->d_counter, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&ctx->d_stats, sizeof(Counters)) != cudaSuccess || cudaEventCreate(&ctx->ev0) != cudaSuccess ||
        cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return DTOF_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    *out = ctx;
    return DTOF_OK;
}

void dtof_destroy(dtof_ctx *ctx) {
    if (!ctx)
        return;
    for (dtof_ctx *p : ctx->peers)
        dtof_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    free_scene(ctx);
    free_wavefront(ctx);
    if (ctx->wf_host_count) cudaFreeHost(ctx->wf_host_count);
    for (int i = 0; i < kWfMaxSets; ++i) {
        if (ctx->wf_stream[i]) cudaStreamDestroy(ctx->wf_stream[i]);
        if (ctx->wf_ev_done[i]) cudaEventDestroy(ctx->wf_ev_done[i]);
    }
    if (ctx->wf_ev_start) cudaEventDestroy(ctx->wf_ev_start);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    delete ctx;
}

const char *dtof_last_error(const dtof_ctx *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

// Device half of dtof_upload_scene: copies a prepared (flattened, BVH built) scene to this context's GPU.
static dtof_status upload_prepared(dtof_ctx *ctx, const dtof_scene_desc *sc, HostScene &H) {
    dtof_status s;
    CU(cudaSetDevice(ctx->device));
    std::vector<MeshRec> &meshes = H.meshes;
    std::vector<BsdfRec> &bsdfs = H.bsdfs;
    std::vector<EmitterRec> &emitters = H.emitters;
    std::vector<TriShade> &shade = H.shade;
    std::vector<float> &cdf = H.cdf, &pmf = H.pmf;
    std::vector<InstRec> &insts = H.insts;
    std::vector<TriIsect> &tris_flat = H.tris_flat;
    BuiltScene &built = H.built;
    const uint32_t gid = H.n_tris;
    // ---- upload (asynchronous renders of the previous scene must have finished, see dtof_update_instances)
    CU(cudaDeviceSynchronize());
    free_scene(ctx);
    if ((s = upload_vec(ctx, built.nodes, &ctx->d_nodes)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, built.tris, &ctx->d_tris)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, shade, &ctx->d_shade)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, insts, &ctx->d_insts)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, tris_flat, &ctx->d_tris_flat)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, built.inst_box, &ctx->d_boxes)) != DTOF_OK) return s;
    ctx->flat_bytes = tris_flat.size() * sizeof(TriIsect);
    ctx->boxes_bytes = built.inst_box.size() * sizeof(InstBox);
    ctx->n_tris_total = gid;
    if ((s = upload_vec(ctx, meshes, &ctx->d_meshes)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, bsdfs, &ctx->d_bsdfs)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, emitters, &ctx->d_emitters)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, H.spots, &ctx->d_spots)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, cdf, &ctx->d_cdf)) != DTOF_OK) return s;
    if ((s = upload_vec(ctx, pmf, &ctx->d_pmf)) != DTOF_OK) return s;
    ctx->h_insts = insts;
    ctx->h_boxes = built.inst_box;
    ctx->nodes_bytes = built.nodes.size() * sizeof(BvhNode);
    ctx->tris_bytes = built.tris.size() * sizeof(TriIsect);
    ctx->insts_bytes = insts.size() * sizeof(InstRec);
    ctx->bvh_depth = built.max_depth + built.tlas_depth;
    ctx->blas_depth = built.max_depth;
    ctx->tlas_begin = built.tlas_begin;
    ctx->n_tlas_nodes = (uint32_t) built.nodes.size() - built.tlas_begin;
    ctx->h_tlas.assign(sc->n_instances, TlasEntry{});
    for (uint32_t g = 0; g < sc->n_instances; ++g) {
        TlasEntry &E = ctx->h_tlas[g];
        E.animated = sc->instances[g].animated != 0;
        memcpy(E.m0, sc->instances[g].m0, sizeof(E.m0));
        memcpy(E.m1, sc->instances[g].m1, sizeof(E.m1));
        E.object_box = built.group_box[g];
        E.blas_root = built.inst_root[g];
    }
    DeviceScene &D = ctx->ds;
    D.nodes = (const float4 *) ctx->d_nodes;
    D.tris = (const float4 *) ctx->d_tris;
    D.shade = (const float4 *) ctx->d_shade;
    D.insts = (const float4 *) ctx->d_insts;
    D.meshes = (const MeshRec *) ctx->d_meshes;
    D.bsdfs = (const BsdfRec *) ctx->d_bsdfs;
    D.emitters = (const EmitterRec *) ctx->d_emitters;
    D.spots = (const SpotRec *) ctx->d_spots;
    D.area_cdf = (const float *) ctx->d_cdf;
    D.area_pmf = (const float *) ctx->d_pmf;
    D.root = built.root;
    D.n_emitters = sc->n_emitters;
    D.n_insts = sc->n_instances;
    D.n_nodes = (uint32_t) built.nodes.size();
    D.n_tris = (uint32_t) built.tris.size();
    D.has_geometry = built.has_geometry ? 1u : 0u;
    D.extended = H.extended ? 1u : 0u;
    D.env_emitter = H.env_emitter;
    if (H.env_emitter >= 0) {
        const dtof_emitter &e = sc->emitters[H.env_emitter];
        D.env_r = e.value[0], D.env_g = e.value[1], D.env_b = e.value[2];
    }
    D.env_cx = H.env_center[0], D.env_cy = H.env_center[1], D.env_cz = H.env_center[2];
    D.env_radius = H.env_radius;
    ctx->cam = sc->camera;
    ctx->film = sc->film;
    ctx->film_px = (size_t) sc->film.width * sc->film.height;
    CU(cudaMalloc(&ctx->d_rgbw, ctx->film_px * 4 * sizeof(float)));
    CU(cudaMalloc(&ctx->d_img, ctx->film_px * 3 * sizeof(float)));
    ctx->has_scene = true;
    return DTOF_OK;
}

dtof_status dtof_upload_scene(dtof_ctx *ctx, const dtof_scene_desc *sc) {
    if (!ctx || !sc)
        return DTOF_ERR_INVALID;
    NvtxRange r_upload("dtof.upload_scene");
    HostScene H;   // flattened + BVH built ONCE, then copied to every device of the context
    dtof_status s;
    {
        NvtxRange r_bvh("dtof.bvh_build");
        s = prepare_scene(ctx, sc, H);
    }
    if (s != DTOF_OK)
        return s;
    NvtxRange r_h2d("dtof.scene_h2d");
    if ((s = upload_prepared(ctx, sc, H)) != DTOF_OK)
        return s;
    for (dtof_ctx *p : ctx->peers)
        if ((s = upload_prepared(p, sc, H)) != DTOF_OK)
            return fail(ctx, s, "device %d: %s", p->device, p->error.c_str());
    return DTOF_OK;
}

dtof_status dtof_scene_info_for(const dtof_scene_desc *sc, dtof_scene_info *out, char *err, uint32_t err_len) {
    if (!sc || !out)
        return DTOF_ERR_INVALID;
    dtof_ctx tmp;   // host-only: carries the error string, never touches CUDA
    HostScene H;
    auto t0 = std::chrono::steady_clock::now();
    dtof_status s = prepare_scene(&tmp, sc, H);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (s != DTOF_OK) {
        if (err && err_len)
            snprintf(err, err_len, "%s", tmp.error.c_str());
        return s;
    }
    out->n_triangles = H.n_tris;
    out->n_nodes = (uint32_t) H.built.nodes.size();
    out->n_instances = sc->n_instances;
    out->bvh_depth = (uint32_t) (H.built.max_depth + H.built.tlas_depth);
    out->traversal_bytes = (uint64_t) H.built.nodes.size() * sizeof(BvhNode) + (uint64_t) H.built.tris.size() * sizeof(TriIsect) +
                           (uint64_t) H.insts.size() * sizeof(InstRec);
    out->shading_bytes = (uint64_t) H.shade.size() * sizeof(TriShade);
    out->build_ms = (float) ms;
    return DTOF_OK;
}

dtof_status dtof_update_instances(dtof_ctx *ctx, uint32_t first, uint32_t n, const dtof_instance *instances) {
    if (!ctx || !instances)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    if ((uint64_t) first + n > ctx->h_insts.size())
        return fail(ctx, DTOF_ERR_INVALID, "instance range out of bounds");
    CU(cudaSetDevice(ctx->device));
    // The keyframes move, the BLASes stay: the TLAS (k - 1 nodes over k instances) is rebuilt on the host from the
    // groups' object-space bounds and overwrites the old TLAS nodes in place -- an animation loop never re-uploads
    // geometry (doppler_tutorials/src/main_animation.py:61-157 reloads the whole scene per frame).
    for (uint32_t i = 0; i < n; ++i) {
        const InstRec &r = ctx->h_insts[first + i];
        if ((r.animated != 0) != (instances[i].animated != 0))
            return fail(ctx, DTOF_ERR_INVALID, "cannot change the animated flag of instance %u", first + i);
        if (instances[i].animated && !(instances[i].t1 > instances[i].t0))
            return fail(ctx, DTOF_ERR_INVALID, "instance %u: keyframe times must be increasing", first + i);
    }
    for (uint32_t i = 0; i < n; ++i) {
        InstRec &r = ctx->h_insts[first + i];
        memcpy(r.m0, instances[i].m0, sizeof(r.m0));
        memcpy(r.m1, instances[i].m1, sizeof(r.m1));
        r.t0 = instances[i].t0, r.t1 = instances[i].t1;
        TlasEntry &E = ctx->h_tlas[first + i];
        memcpy(E.m0, instances[i].m0, sizeof(E.m0));
        memcpy(E.m1, instances[i].m1, sizeof(E.m1));
    }
    // renders launched with dtof_render_device are asynchronous and the wavefront pipeline runs on internal
    // non-blocking streams, which a plain cudaMemcpy does not wait for: finish them before the scene changes
    CU(cudaDeviceSynchronize());
    BuiltScene tl;
    std::vector<BvhNode> tlas;
    build_tlas(ctx->h_tlas, (int32_t) ctx->tlas_begin, tlas, tl);
    if (tlas.size() != ctx->n_tlas_nodes)
        return fail(ctx, DTOF_ERR_STATE, "TLAS rebuild produced %zu nodes, expected %u", tlas.size(), ctx->n_tlas_nodes);
    if (ctx->blas_depth + tl.tlas_depth + 4 > kStackSize)
        return fail(ctx, DTOF_ERR_UNSUPPORTED, "BVH too deep for the traversal stack (%d + %d)", ctx->blas_depth, tl.tlas_depth);
    CU(cudaMemcpy((char *) ctx->d_insts + first * sizeof(InstRec), ctx->h_insts.data() + first, n * sizeof(InstRec),
                  cudaMemcpyHostToDevice));
    if (!tlas.empty())
        CU(cudaMemcpy((char *) ctx->d_nodes + (size_t) ctx->tlas_begin * sizeof(BvhNode), tlas.data(), tlas.size() * sizeof(BvhNode),
                      cudaMemcpyHostToDevice));
    if (!tl.inst_box.empty())
        CU(cudaMemcpy(ctx->d_boxes, tl.inst_box.data(), tl.inst_box.size() * sizeof(InstBox), cudaMemcpyHostToDevice));
    ctx->h_boxes = tl.inst_box;
    ctx->ds.root = tl.root;
    ctx->bvh_depth = ctx->blas_depth + tl.tlas_depth;
    for (dtof_ctx *p : ctx->peers) {
        dtof_status ps = dtof_update_instances(p, first, n, instances);
        if (ps != DTOF_OK)
            return fail(ctx, ps, "device %d: %s", p->device, p->error.c_str());
    }
    CU(cudaSetDevice(ctx->device));
    return DTOF_OK;
}

dtof_status dtof_pass_info_for(const dtof_ctx *cctx, const dtof_params *params, dtof_pass_info *out) {
    dtof_ctx *ctx = const_cast<dtof_ctx *>(cctx);
    if (!ctx || !params || !out)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    int rc = pass_info(ctx->film, *params, out);
    if (rc)
        return fail(ctx, DTOF_ERR_INVALID,
                    rc == 2 ? "sample_count should be a multiple of samples_per_wavefront!" : "invalid sample_count");
    return DTOF_OK;
}

dtof_status dtof_render_device(dtof_ctx *ctx, const dtof_params *params, float *d_rgbw, void *stream) {
    if (!ctx)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    if (!d_rgbw)
        return fail(ctx, DTOF_ERR_INVALID, "d_rgbw is NULL");
    dtof_status s = check_params(ctx, params);
    if (s != DTOF_OK)
        return s;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->peers.empty())
        return render_sharded(ctx, params, d_rgbw, false, (cudaStream_t) stream);
    return launch_render(ctx, params, d_rgbw, (cudaStream_t) stream, nullptr, nullptr, 0);
}

dtof_status dtof_develop_device(dtof_ctx *ctx, const float *d_rgbw, float *d_image, void *stream) {
    if (!ctx)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    if (!d_rgbw || !d_image)
        return fail(ctx, DTOF_ERR_INVALID, "NULL tensor");
    CU(cudaSetDevice(ctx->device));
    size_t n = ctx->film_px;
    develop_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, (cudaStream_t) stream>>>((const float4 *) d_rgbw, d_image, n);
    ctx->launches++;
    CU(cudaGetLastError());
    return DTOF_OK;
}

dtof_status dtof_render(dtof_ctx *ctx, const dtof_params *params, float *rgbw_out, float *image_out) {
    if (!ctx)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    dtof_status s = check_params(ctx, params);
    if (s != DTOF_OK)
        return s;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->peers.empty()) {
        if ((s = render_sharded(ctx, params, ctx->d_rgbw, true, 0)) != DTOF_OK)
            return s;
    } else {
        CU(cudaMemsetAsync(ctx->d_rgbw, 0, ctx->film_px * 4 * sizeof(float), 0));
        if ((s = launch_render(ctx, params, ctx->d_rgbw, 0, nullptr, nullptr, 0)) != DTOF_OK)
            return s;
    }
    if (image_out && (s = dtof_develop_device(ctx, ctx->d_rgbw, ctx->d_img, nullptr)) != DTOF_OK)
        return s;
    if (rgbw_out)
        CU(cudaMemcpyAsync(rgbw_out, ctx->d_rgbw, ctx->film_px * 4 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    if (image_out)
        CU(cudaMemcpyAsync(image_out, ctx->d_img, ctx->film_px * 3 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    CU(cudaStreamSynchronize(0));
    return DTOF_OK;
}

dtof_status dtof_render_accumulate(dtof_ctx *ctx, const dtof_params *params, int zero_first) {
    if (!ctx)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    dtof_status s = check_params(ctx, params);
    if (s != DTOF_OK)
        return s;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->peers.empty()) {
        // the peers' partial films are summed into device 0's film after every chunk (their films restart from zero)
        if (zero_first)
            CU(cudaMemsetAsync(ctx->d_rgbw, 0, ctx->film_px * 4 * sizeof(float), 0));
        if ((s = render_sharded(ctx, params, ctx->d_rgbw, false, 0)) != DTOF_OK)
            return s;
    } else {
        if (zero_first)
            CU(cudaMemsetAsync(ctx->d_rgbw, 0, ctx->film_px * 4 * sizeof(float), 0));
        if ((s = launch_render(ctx, params, ctx->d_rgbw, 0, nullptr, nullptr, 0)) != DTOF_OK)
            return s;
    }
    CU(cudaStreamSynchronize(0));
    return DTOF_OK;
}

dtof_status dtof_read_film(dtof_ctx *ctx, float *rgbw_out, float *image_out) {
    if (!ctx)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    CU(cudaSetDevice(ctx->device));
    dtof_status s;
    if (image_out && (s = dtof_develop_device(ctx, ctx->d_rgbw, ctx->d_img, nullptr)) != DTOF_OK)
        return s;
    if (rgbw_out)
        CU(cudaMemcpyAsync(rgbw_out, ctx->d_rgbw, ctx->film_px * 4 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    if (image_out)
        CU(cudaMemcpyAsync(image_out, ctx->d_img, ctx->film_px * 3 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    CU(cudaStreamSynchronize(0));
    return DTOF_OK;
}

dtof_status dtof_render_multi_pass(dtof_ctx *ctx, const dtof_params *params, uint32_t n_renders, float *image_out) {
    if (!ctx)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    if (!image_out || n_renders == 0)
        return fail(ctx, DTOF_ERR_INVALID, "image_out is NULL or n_renders is 0");
    dtof_status s = check_params(ctx, params);
    if (s != DTOF_OK)
        return s;
    CU(cudaSetDevice(ctx->device));
    const size_t n = ctx->film_px;
    const float scale = 1.f / (float) n_renders;
    // a multi-device context deals the renders (seeds) round-robin: device i develops and averages renders i, i + D, ...
    // into its own image, device 0 then adds the partial means (mean of DEVELOPED images, as the tutorials' driver does)
    const uint32_t n_dev = 1u + (uint32_t) ctx->peers.size();
    s = for_each_device(ctx, [&](size_t dev, dtof_ctx *c) -> dtof_status {
        dtof_ctx *ctx = c;   // for CU()
        bool first = true;
        for (uint32_t i = (uint32_t) dev; i < n_renders; i += n_dev, first = false) {
            dtof_params p = *params;
            p.seed = params->seed + i;
            CU(cudaMemsetAsync(c->d_rgbw, 0, n * 4 * sizeof(float), 0));
            dtof_status r = launch_render(c, &p, c->d_rgbw, 0, nullptr, nullptr, 0);
            if (r != DTOF_OK)
                return r;
            develop_accumulate_kernel<<<(unsigned) ((n + 255) / 256), 256>>>((const float4 *) c->d_rgbw, c->d_img, n, scale, first);
            c->launches++;
            CU(cudaGetLastError());
        }
        if (first)   // more devices than renders: this one contributes nothing
            CU(cudaMemsetAsync(c->d_img, 0, n * 3 * sizeof(float), 0));
        return DTOF_OK;
    });
    if (s != DTOF_OK)
        return s;
    if (n_dev > 1 && (s = reduce_from_peers(ctx, ctx->d_img, n * 3, true, 0)) != DTOF_OK)
        return s;
    CU(cudaMemcpyAsync(image_out, ctx->d_img, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    CU(cudaStreamSynchronize(0));
    return DTOF_OK;
}

dtof_status dtof_trace_samples(dtof_ctx *ctx, const dtof_params *params, const uint64_t *lanes, uint32_t n,
                               dtof_sample_record *out) {
    return dtof_trace_samples_pass(ctx, params, lanes, n, 0, out);
}

dtof_status dtof_trace_samples_pass(dtof_ctx *ctx, const dtof_params *params, const uint64_t *lanes, uint32_t n, uint32_t pass,
                                    dtof_sample_record *out) {
    if (!ctx)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    if (n == 0)
        return DTOF_OK;
    if (!lanes || !out)
        return fail(ctx, DTOF_ERR_INVALID, "NULL lanes/out");
    dtof_status s = check_params(ctx, params);
    if (s != DTOF_OK)
        return s;
    dtof_pass_info pi;
    if (pass_info(ctx->film, *params, &pi))
        return fail(ctx, DTOF_ERR_INVALID, "sample_count should be a multiple of samples_per_wavefront!");
    if (pass >= pi.n_passes)
        return fail(ctx, DTOF_ERR_INVALID, "pass %u outside the %u passes of this render", pass, pi.n_passes);
    for (uint32_t i = 0; i < n; ++i)
        if (lanes[i] >= pi.wavefront_size)
            return fail(ctx, DTOF_ERR_INVALID, "lane %llu outside the wavefront", (unsigned long long) lanes[i]);
    CU(cudaSetDevice(ctx->device));
    unsigned long long *d_lanes = nullptr;
    dtof_sample_record *d_rec = nullptr;
    CU(cudaMalloc(&d_lanes, n * sizeof(unsigned long long)));
    cudaError_t e = cudaMalloc(&d_rec, n * sizeof(dtof_sample_record));
    if (e != cudaSuccess) {
        cudaFree(d_lanes);
        return fail(ctx, DTOF_ERR_NOMEM, "cudaMalloc failed");
    }
    cudaMemcpy(d_lanes, lanes, n * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    dtof_params p = *params;
    p.lane_begin = 0;
    p.lane_end = 0;
    p.shard_block = 0;
    s = launch_render(ctx, &p, ctx->d_rgbw, 0, d_lanes, d_rec, n, pass);
    if (s == DTOF_OK) {
        e = cudaMemcpy(out, d_rec, n * sizeof(dtof_sample_record), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess)
            s = fail(ctx, DTOF_ERR_CUDA, "readback failed: %s", cudaGetErrorString(e));
    }
    cudaFree(d_lanes);
    cudaFree(d_rec);
    return s;
}

dtof_status dtof_trace_rays(dtof_ctx *ctx, const dtof_ray *rays, uint32_t n, int any_hit, dtof_ray_hit *out) {
    if (!ctx)
        return DTOF_ERR_INVALID;
    if (!ctx->has_scene)
        return fail(ctx, DTOF_ERR_STATE, "no scene uploaded");
    if (n == 0)
        return DTOF_OK;
    if (!rays || !out)
        return fail(ctx, DTOF_ERR_INVALID, "NULL rays/out");
    CU(cudaSetDevice(ctx->device));
    RenderArgs A{};
    A.scene = ctx->ds;
    A.tris_flat = (const float4 *) ctx->d_tris_flat;
    A.inst_box = (const float4 *) ctx->d_boxes;
    A.stack_levels = (uint32_t) ctx->bvh_depth + 4u;
    A.nodes_bytes = (uint32_t) ctx->nodes_bytes, A.tris_bytes = (uint32_t) ctx->tris_bytes;
    A.insts_bytes = (uint32_t) ctx->insts_bytes, A.boxes_bytes = (uint32_t) ctx->boxes_bytes;
    const size_t smem = ctx->nodes_bytes + ctx->tris_bytes + ctx->insts_bytes + ctx->boxes_bytes +
                        (size_t) A.stack_levels * 128u * (kBlock / 32);
    int mode = (smem <= kSmemSceneLimit && smem + 1024 <= ctx->smem_optin) ? MODE_BVH_SMEM : MODE_BVH_GLOBAL;
    if (const char *e = getenv("DTOF_MODE"))
        if (atoi(e) == MODE_BVH_GLOBAL)
            mode = MODE_BVH_GLOBAL;
    dtof_ray *d_rays = nullptr;
    dtof_ray_hit *d_out = nullptr;
    CU(cudaMalloc(&d_rays, n * sizeof(dtof_ray)));
    if (cudaMalloc(&d_out, n * sizeof(dtof_ray_hit)) != cudaSuccess) {
        cudaFree(d_rays);
        return fail(ctx, DTOF_ERR_NOMEM, "cudaMalloc failed");
    }
    cudaMemcpy(d_rays, rays, n * sizeof(dtof_ray), cudaMemcpyHostToDevice);
    const unsigned grid = (n + kBlock - 1) / kBlock;
    if (mode == MODE_BVH_SMEM) {
        cudaFuncSetAttribute(rays_kernel<MODE_BVH_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        rays_kernel<MODE_BVH_SMEM><<<grid, kBlock, smem>>>(A, d_rays, n, any_hit, d_out);
    } else {
        rays_kernel<MODE_BVH_GLOBAL><<<grid, kBlock>>>(A, d_rays, n, any_hit, d_out);
    }
    ctx->launches++;
    ctx->last_mode = mode;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpy(out, d_out, n * sizeof(dtof_ray_hit), cudaMemcpyDeviceToHost);
    cudaFree(d_rays);
    cudaFree(d_out);
    if (e != cudaSuccess)
        return fail(ctx, DTOF_ERR_CUDA, "dtof_trace_rays: %s", cudaGetErrorString(e));
    return DTOF_OK;
}

dtof_status dtof_set_stats(dtof_ctx *ctx, int enabled) {
    if (!ctx)
        return DTOF_ERR_INVALID;
    ctx->stats_enabled = enabled != 0;
    return DTOF_OK;
}

dtof_status dtof_get_stats(dtof_ctx *ctx, dtof_stats *out) {
    if (!ctx || !out)
        return DTOF_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    Counters c;
    CU(cudaMemcpy(&c, ctx->d_stats, sizeof(c), cudaMemcpyDeviceToHost));
    out->samples = c.samples;
    out->rays_closest = c.rays_closest;
    out->rays_shadow = c.rays_shadow;
    out->nodes_visited = c.nodes;
    out->tris_tested = c.tris;
    out->inst_visits = c.inst;
    return DTOF_OK;
}

int dtof_last_traversal_mode(const dtof_ctx *ctx) { return ctx ? ctx->last_mode : -1; }

int dtof_last_pipeline(const dtof_ctx *ctx) { return ctx && ctx->have_timing ? ctx->last_pipeline : -1; }

uint64_t dtof_launch_count(const dtof_ctx *ctx) { return ctx ? ctx->launches : 0; }

dtof_status dtof_last_kernel_ms(dtof_ctx *ctx, float *ms) {
    if (!ctx || !ms)
        return DTOF_ERR_INVALID;
    if (!ctx->have_timing)
        return fail(ctx, DTOF_ERR_STATE, "no render has been launched");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev1));
    CU(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return DTOF_OK;
}

} // extern "C"
