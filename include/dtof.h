/*
 * dtof.h -- C ABI of libdtof_b200.so, the B200-native (sm_100a) Doppler time-of-flight path tracer.
 *
 * This is the drop-in boundary for ONE path of juhyeonkim95/Mitsuba3DopplerToF: the
 * `dopplertofpath` integrator driven by the `correlated` sampler over motion-blurred scenes.
 * A host plugin (C++ Mitsuba plugin, or the Python mirror in mitsuba3dopplertof_b200/) flattens
 * a loaded scene into the plain arrays below and calls these entry points; nothing else crosses
 * the boundary (no torch / Dr.Jit / Mitsuba types, no C++ exceptions).
 *
 * Reference interfaces replaced (file:line relative to the reference repository):
 *   dtof_upload_scene   <- Scene ctor + accel build: src/render/scene.cpp:22-100,
 *                          src/render/scene_embree.inl:84-128, Instance::embree_geometry
 *                          src/shapes/instance.cpp:295-310 (2-keyframe matrix motion),
 *                          ShapeGroup src/render/shapegroup.cpp:7-69
 *   dtof_render         <- SamplingIntegrator::render (JIT branch) src/render/integrator.cpp:104-347,
 *                          render_sample Doppler branch :476-542, DopplerToFPathIntegrator::sample
 *                          src/integrators/dopplertofpath.cpp:79-283, CorrelatedSampler
 *                          src/samplers/correlated.cpp:38-167, ImageBlock::put
 *                          src/render/imageblock.cpp:418-531, HDRFilm::develop src/films/hdrfilm.cpp:305-419
 *   dtof_trace_samples  <- SamplingIntegrator::sample (public virtual, include/mitsuba/render/integrator.h:200-205)
 *                          evaluated for chosen wavefront lanes; the per-sample parity hook
 *   dtof_params         <- the property surface parsed in src/render/integrator.cpp:54-100,568-585,
 *                          src/integrators/dopplertofpath.cpp:19-57, src/render/sampler.cpp:13-14,
 *                          src/samplers/correlated.cpp:17-23
 *
 * Conventions: all matrices are row-major; 3x4 affine = rows of [R | t]; all floats are IEEE
 * binary32; the caller owns every host buffer it passes; the library owns all device memory.
 * One context per GPU; a context is thread-compatible (calls on one context must not overlap).
 * Every function returns DTOF_OK or an error code; dtof_last_error() gives the message.
 */
#ifndef DTOF_H
#define DTOF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTOF_ABI_VERSION 8

typedef struct dtof_ctx dtof_ctx;

typedef enum dtof_status {
    DTOF_OK = 0,
    DTOF_ERR_INVALID = 1,     /* bad argument / unsupported property value (reference: Throw(...)) */
    DTOF_ERR_CUDA = 2,        /* CUDA runtime failure */
    DTOF_ERR_NOMEM = 3,
    DTOF_ERR_UNSUPPORTED = 4, /* feature outside the hot-path scope */
    DTOF_ERR_STATE = 5        /* e.g. render before upload */
} dtof_status;

/* ---- enums mirroring the reference ---------------------------------------------------- */

/* ETimeSampling, include/mitsuba/render/sampler.h:27-34 */
typedef enum dtof_time_sampling {
    DTOF_TIME_UNIFORM = 0,
    DTOF_TIME_STRATIFIED = 1,
    DTOF_TIME_ANTITHETIC = 2,
    DTOF_TIME_ANTITHETIC_MIRROR = 3
} dtof_time_sampling;

/* EWaveformType, include/mitsuba/render/waveform_utils.h:11-16 */
typedef enum dtof_waveform {
    DTOF_WAVE_SINUSOIDAL = 0,
    DTOF_WAVE_RECTANGULAR = 1,
    DTOF_WAVE_TRIANGULAR = 2,
    DTOF_WAVE_TRAPEZOIDAL = 3
} dtof_waveform;

/* Which SamplingIntegrator::sample() the lanes run. VELOCITY = the reference's ground-truth radial-velocity
 * integrator (src/integrators/velocity.cpp:113-127): two closest-hit queries of the camera ray at t = 0 and t = `time`,
 * value (t2 - t1) / time in all three channels; it is NOT a Doppler integrator, so render_sample takes the stock
 * branch (src/render/integrator.cpp:409-472): jitter and time come from the sampler's independent stream only.
 * PATH = the stock path tracer (src/integrators/path.cpp:103-283) the tutorials use for the radiance pass
 * (doppler_tutorials/src/program_runner.py:57-80): the same bounce loop as dopplertofpath without the modulation
 * weight and without the time wrap, every draw from Sampler::next_1d/next_2d = the independent stream
 * (src/samplers/correlated.cpp:78-90), stock render_sample branch. */
typedef enum dtof_integrator_kind {
    DTOF_INTEGRATOR_DOPPLERTOFPATH = 0,
    DTOF_INTEGRATOR_VELOCITY = 1,
    DTOF_INTEGRATOR_PATH = 2
} dtof_integrator_kind;

typedef enum dtof_rfilter {
    DTOF_RFILTER_BOX = 0, DTOF_RFILTER_TENT = 1, DTOF_RFILTER_GAUSSIAN = 2,
    DTOF_RFILTER_MITCHELL = 3, DTOF_RFILTER_CATMULLROM = 4, DTOF_RFILTER_LANCZOS = 5
} dtof_rfilter;
typedef enum dtof_shape_kind { DTOF_SHAPE_MESH = 0, DTOF_SHAPE_RECTANGLE = 1 } dtof_shape_kind;
typedef enum dtof_bsdf_kind {
    DTOF_BSDF_DIFFUSE = 0, DTOF_BSDF_NULL_BLACK = 1, DTOF_BSDF_CONDUCTOR = 2, DTOF_BSDF_DIELECTRIC = 3,
    DTOF_BSDF_THINDIELECTRIC = 4, DTOF_BSDF_PLASTIC = 5, DTOF_BSDF_ROUGHCONDUCTOR = 6, DTOF_BSDF_ROUGHDIELECTRIC = 7
} dtof_bsdf_kind;
typedef enum dtof_emitter_kind {
    DTOF_EMITTER_POINT = 0, DTOF_EMITTER_AREA = 1, DTOF_EMITTER_CONSTANT = 2, DTOF_EMITTER_SPOT = 3,
    DTOF_EMITTER_DIRECTIONAL = 4
} dtof_emitter_kind;

/* ---- scene description ------------------------------------------------------------------ */

/* One triangle mesh (Mesh, src/render/mesh.cpp; Cube = 24 verts/12 faces, src/shapes/cube.cpp:109-165;
 * Rectangle = 4 verts/2 faces plus the analytic parametrisation kept for emitter sampling,
 * src/shapes/rectangle.cpp:101-113,152-166). Static meshes are given in world space; meshes that
 * belong to an animated instance are given in the shape group's object space. */
typedef struct dtof_mesh {
    uint32_t n_vertices;
    uint32_t n_faces;
    const float *positions;   /* 3 * n_vertices */
    const float *normals;     /* 3 * n_vertices, or NULL (flat shading) */
    const float *texcoords;   /* 2 * n_vertices, or NULL */
    const uint32_t *faces;    /* 3 * n_faces */
    uint32_t bsdf;            /* index into dtof_scene_desc::bsdfs */
    int32_t emitter;          /* index of the attached area emitter, or -1 */
    uint32_t flip_normals;    /* Mesh::m_flip_normals */
    uint32_t kind;            /* dtof_shape_kind */
    float rect_to_world[12];  /* kind == RECTANGLE: to_world (maps [-1,1]^2 x {0}), else ignored */
} dtof_mesh;

/* One top-level instance = one shape group placed by a (possibly animated) transform.
 * Motion model: AnimatedTransform::eval, include/mitsuba/core/transform.h:440-466 -- the 4x4 matrix is
 * linearly interpolated between keyframe 0 and keyframe 1. `animated == 0` marks the static group:
 * its meshes are already in world space and both matrices are ignored. */
typedef struct dtof_instance {
    uint32_t first_mesh;      /* range of meshes forming the shape group */
    uint32_t n_meshes;
    uint32_t animated;
    float t0, t1;             /* keyframe times (AnimatedTransform::get_min/max_time) */
    float m0[12];             /* to_world at t0, row-major 3x4 */
    float m1[12];             /* to_world at t1 */
} dtof_instance;

/* SmoothDiffuse (src/bsdfs/diffuse.cpp), SmoothConductor (src/bsdfs/conductor.cpp: perfect specular reflection
 * weighted by specular_reflectance * fresnel_conductor(cos_theta_i, eta + i k), include/mitsuba/render/fresnel.h:93-117;
 * material "none" is eta = 0, k = 1) or SmoothDielectric (src/bsdfs/dielectric.cpp: Fresnel-weighted choice between
 * specular reflection and refraction, never two-sided) or ThinDielectric (src/bsdfs/thindielectric.cpp: a thin slab,
 * r' = 2r / (1 + r), the transmitted ray goes straight on as a Null interaction; same fields as DIELECTRIC) or
 * SmoothPlastic (src/bsdfs/plastic.cpp: a diffuse base under a smooth dielectric coat) or RoughConductor
 * (src/bsdfs/roughconductor.cpp: Beckmann or GGX microfacets with visible-normal sampling,
 * include/mitsuba/render/microfacet.h); diffuse, conductor, plastic and roughconductor optionally wrapped by
 * TwoSidedBRDF (src/bsdfs/twosided.cpp). */
typedef struct dtof_bsdf {
    uint32_t kind;            /* dtof_bsdf_kind */
    uint32_t twosided;
    float reflectance[3];     /* DIFFUSE: reflectance; CONDUCTOR, DIELECTRIC: specular_reflectance; PLASTIC: diffuse_reflectance */
    float eta[3];             /* CONDUCTOR: real part of the index of refraction (RGB); DIELECTRIC, PLASTIC: eta[0] =
                               * int_ior / ext_ior; PLASTIC: eta[1] != 0 = `nonlinear` */
    float k[3];               /* CONDUCTOR: extinction coefficient (RGB); DIELECTRIC: specular_transmittance;
                               * PLASTIC: specular_reflectance */
    float alpha[2];           /* ROUGHCONDUCTOR, ROUGHDIELECTRIC (src/bsdfs/roughdielectric.cpp; other fields as for
                               * CONDUCTOR / DIELECTRIC): alpha_u, alpha_v */
    uint32_t distribution;    /* ROUGHCONDUCTOR, ROUGHDIELECTRIC: 0 = beckmann, 1 = ggx (MicrofacetType) */
    uint32_t reserved;        /* 0 */
} dtof_bsdf;

/* PointLight (src/emitters/point.cpp), AreaLight (src/emitters/area.cpp) on mesh `mesh`, or the constant environment
 * emitter (src/emitters/constant.cpp; at most one per scene, scene.cpp:52-56). The library derives the environment's
 * bounding sphere from the uploaded geometry as ConstantBackgroundEmitter::set_scene does (constant.cpp:73-82). The
 * emitters' order is the order of Scene::m_emitters (scene.cpp:34-50): it decides which one a sample selects. */
typedef struct dtof_emitter {
    uint32_t kind;            /* dtof_emitter_kind */
    uint32_t mesh;            /* AREA: index of the emitting mesh (must be in the static group) */
    float position[3];        /* POINT, SPOT: position; DIRECTIONAL (src/emitters/directional.cpp): the direction the light
                               * travels in, to_world * (0, 0, 1); like CONSTANT it uses the scene's bounding sphere */
    float value[3];           /* POINT, SPOT: intensity; AREA, CONSTANT: radiance; DIRECTIONAL: irradiance */
    /* SPOT (src/emitters/spot.cpp; no projection texture): `position` = translation of to_world; the linear part of
     * to_world^-1, row-major, takes world directions into the light's frame (it looks along +z); angles in radians:
     * the intensity ramps linearly from 0 at cutoff_angle to its full value at beam_width (spot.cpp:146-154) */
    float to_local[9];
    float cutoff_angle, beam_width;
} dtof_emitter;

/* PerspectiveCamera, src/sensors/perspective.cpp:172-279. sample_to_camera is
 * perspective_projection(...).inverse() (include/mitsuba/render/sensor.h:227-262), computed by the host. */
typedef struct dtof_camera {
    float to_world[12];
    float sample_to_camera[16];
    float near_clip, far_clip;
    float shutter_open, shutter_open_time;
} dtof_camera;

/* HDRFilm geometry + reconstruction filter (src/films/hdrfilm.cpp:235-297,
 * src/rfilters/{box,tent,gaussian,mitchell,catmullrom,lanczos}.cpp). */
typedef struct dtof_film {
    uint32_t width, height;            /* crop size == size of the rendered tensor */
    uint32_t crop_offset_x, crop_offset_y;
    uint32_t rfilter;                  /* dtof_rfilter */
    float rfilter_radius;              /* tent: radius; gaussian: 4*stddev; box: 0.5; mitchell, catmullrom: 2; lanczos: lobes */
    float gaussian_stddev;
    float mitchell_b, mitchell_c;      /* MITCHELL only (defaults 1/3, 1/3; mitchell.cpp:52-60) */
} dtof_film;

typedef struct dtof_scene_desc {
    uint32_t n_meshes;
    const dtof_mesh *meshes;
    uint32_t n_instances;
    const dtof_instance *instances;
    uint32_t n_bsdfs;
    const dtof_bsdf *bsdfs;
    uint32_t n_emitters;
    const dtof_emitter *emitters;
    dtof_camera camera;
    dtof_film film;
} dtof_scene_desc;

/* ---- render parameters: the reference's property surface ------------------------------- */

typedef struct dtof_params {
    /* DopplerToFPathIntegrator, src/integrators/dopplertofpath.cpp:19-57 */
    float time;                       /* "time", 0.0015 */
    float w_g;                        /* "w_g" illumination modulation frequency [MHz], 30 */
    float g_1, g_0;                   /* illumination modulation scale / offset, 0.5 / 0.5 */
    float sensor_phase_offset;        /* "sensor_phase_offset", or hetero_offset * 2*pi */
    float hetero_frequency;           /* resolved by the host exactly as :32-38 */
    uint32_t wave_function_type;      /* dtof_waveform */
    uint32_t low_frequency_component_only;
    /* MonteCarloIntegrator, src/render/integrator.cpp:568-585 */
    int32_t max_depth;                /* -1 = unbounded */
    int32_t rr_depth;                 /* 5 */
    uint32_t hide_emitters;
    /* SamplingIntegrator Doppler members, src/render/integrator.cpp:54-100 */
    uint32_t time_sampling_method;    /* dtof_time_sampling, default ANTITHETIC */
    float antithetic_shift;
    uint32_t use_stratified_sampling_for_each_interval;
    uint32_t path_correlation_depth;
    /* Sampler / CorrelatedSampler, src/render/sampler.cpp:13-14, src/samplers/correlated.cpp:17-23 */
    uint32_t sample_count;            /* spp of this render() call */
    uint32_t base_seed;               /* sampler "seed" property */
    uint32_t time_correlate_number;
    uint32_t path_correlate_number;
    /* render() arguments */
    uint32_t seed;
    /* Sharding (multi-GPU): this call renders wavefront lanes [lane_begin, lane_end) of every pass;
     * lane_end == 0 means "all lanes". Lanes are idx = pixel * spp_per_pass + slot as in
     * src/render/integrator.cpp:273-290. */
    uint64_t lane_begin, lane_end;
    /* Interleaved sharding inside [lane_begin, lane_end): the range is cut into blocks of `shard_block` lanes and
     * this call renders the blocks b with b % shard_count == shard_index. shard_block == 0 disables it.
     *   sample-slot (spp) sharding over G GPUs : shard_block = spp_per_pass / G  (a multiple of lcm(tcn, pcn))
     *   interleaved tile sharding              : shard_block = spp_per_pass * pixels_per_tile */
    uint64_t shard_block;
    uint32_t shard_count, shard_index;
    uint32_t integrator;              /* dtof_integrator_kind; 0 = dopplertofpath */
    uint32_t reserved;                /* must be 0 */
} dtof_params;

/* Per-lane record returned by dtof_trace_samples (and by the CPU oracle): everything the
 * reference computes for one wavefront lane of pass 0 before the film splat. */
typedef struct dtof_sample_record {
    float sample_pos[2];   /* film-space position handed to ImageBlock::put */
    float time;            /* sampled time (before the wrap of dopplertofpath.cpp:93) */
    float ray_o[3], ray_d[3], ray_maxt;
    float rgb[3];          /* radiance returned by sample() times ray weight (1) */
    float path_length;     /* accumulated path length at exit */
    uint32_t depth;        /* number of valid surface interactions */
    uint32_t rng_draws;    /* draws consumed from the independent stream */
} dtof_sample_record;

/* Traversal work counters of the last render (filled after dtof_set_stats(ctx, 1)). They count the BVH walk:
 * a scene that normally runs the flat shared-memory walk is rendered through its BVH while stats are on. */
typedef struct dtof_stats {
    uint64_t samples;        /* camera samples traced */
    uint64_t rays_closest, rays_shadow;
    uint64_t nodes_visited;  /* BVH nodes popped (binary-BVH nodes, 64 B each) */
    uint64_t tris_tested;    /* triangles tested (48 B each) */
    uint64_t inst_visits;    /* animated-instance entries (112 B each) */
} dtof_stats;

/* Geometry of a render() call, derived exactly like src/render/integrator.cpp:121-134,227-245. */
typedef struct dtof_pass_info {
    uint32_t spp_per_pass;
    uint32_t n_passes;
    uint64_t wavefront_size; /* lanes per pass */
} dtof_pass_info;

/* Result of the host half of dtof_upload_scene (validation + flattening + BVH build), no GPU needed. */
typedef struct dtof_scene_info {
    uint32_t n_triangles, n_nodes, n_instances, bvh_depth;
    uint64_t traversal_bytes;   /* nodes + leaf-order triangles + instance records */
    uint64_t shading_bytes;     /* per-hit triangle records */
    float build_ms;
} dtof_scene_info;

/* ---- entry points ------------------------------------------------------------------------ */

uint32_t dtof_abi_version(void);

/* Create / destroy a context bound to CUDA device `device`. */
dtof_status dtof_create(dtof_ctx **out, int device);
void dtof_destroy(dtof_ctx *ctx);

/* One context over SEVERAL GPUs of one node, driven by one host thread (what a plugin inside the reference's single
 * `mitsuba` process needs to use the whole box; Integrator::render is one call, src/render/integrator.cpp:104-347).
 * devices[0] is the primary device: it receives the result. The scene is replicated (dtof_upload_scene flattens and builds
 * the BVH once, then copies it to every device); dtof_render / dtof_render_device shard ONE render over the devices --
 * by sample slots in whole correlate groups when spp_per_pass divides by n_devices * lcm(tcn, pcn), else by interleaved
 * 64-pixel tiles; all passes of a lane stay on one device -- and device 0 sums the partial films, reading the peers'
 * films over NVLink (P2P) where the devices can address each other, through a staging copy otherwise.
 * dtof_render_multi_pass deals its renders (seeds) round-robin to the devices and sums the partial means.
 * dtof_update_instances applies to every device. Per-lane records, counters and timings refer to device 0.
 * params->shard_block must be 0 (the context shards by itself); a lane window (lane_begin / lane_end) is honoured.
 * n_devices == 1 is the same as dtof_create. At most 16 devices. */
dtof_status dtof_create_multi(dtof_ctx **out, const int *devices, uint32_t n_devices);
uint32_t dtof_device_count(const dtof_ctx *ctx);
const char *dtof_last_error(const dtof_ctx *ctx);

/* Flatten, build the two-level BVH and upload everything to HBM. May be called again to replace the scene. */
dtof_status dtof_upload_scene(dtof_ctx *ctx, const dtof_scene_desc *scene);

/* Validate a scene description and build its BVH on the host WITHOUT touching the GPU (what the Scene constructor's
 * checks do in the reference, src/render/scene.cpp:22-100, src/render/shapegroup.cpp:27-30). `err` (may be NULL)
 * receives the message on failure. */
dtof_status dtof_scene_info_for(const dtof_scene_desc *scene, dtof_scene_info *out, char *err, uint32_t err_len);

/* Replace the keyframes of instances [first, first+n) without rebuilding bottom-level BVHs
 * (animation loops, doppler_tutorials/src/main_animation.py:61-157). The top-level BVH is rebuilt on the host from
 * the groups' object-space bounds and replaces the old one in place, so the new motion may leave the bounds the scene
 * was uploaded with; geometry, materials and the `animated` flag of an instance cannot change. */
dtof_status dtof_update_instances(dtof_ctx *ctx, uint32_t first, uint32_t n, const dtof_instance *instances);

/* Pass split of a render with these parameters; DTOF_ERR_INVALID where the reference throws
 * (sample_count % spp_per_pass != 0, src/render/sampler.cpp:81-82). */
dtof_status dtof_pass_info_for(const dtof_ctx *ctx, const dtof_params *params, dtof_pass_info *out);

/* Render into HOST buffers. rgbw_out: height*width*4 floats (R,G,B,W accumulation, the ImageBlock tensor);
 * image_out: height*width*3 floats (developed RGB = RGB/W); either may be NULL. Copies are inside. */
dtof_status dtof_render(dtof_ctx *ctx, const dtof_params *params, float *rgbw_out, float *image_out);

/* A render in pieces, for callers that must be able to stop between them (Integrator::cancel / the `timeout` property,
 * include/mitsuba/render/integrator.h:106-108): dtof_render_accumulate renders the lanes [lane_begin, lane_end) of
 * `params` (all their passes) INTO the context's own device film -- zeroed first if `zero_first` -- and returns when
 * they are done; dtof_read_film develops and copies that film to host buffers (either may be NULL). dtof_render is
 * accumulate(all lanes, zero_first = 1) + read_film. */
dtof_status dtof_render_accumulate(dtof_ctx *ctx, const dtof_params *params, int zero_first);
dtof_status dtof_read_film(dtof_ctx *ctx, float *rgbw_out, float *image_out);

/* The tutorials' multi-pass driver (render_image_multi_pass, doppler_tutorials/src/program_runner.py:11-31):
 * n_renders renders with seed = params->seed + i (i = 0 .. n_renders-1), each developed (RGB / W), averaged ON THE
 * DEVICE; the scene stays resident and only the final H*W*3 image crosses to the host. This is also how a 16k-spp
 * image is produced (16 renders of 1024 spp; a single 16k-spp render() throws in the reference). */
dtof_status dtof_render_multi_pass(dtof_ctx *ctx, const dtof_params *params, uint32_t n_renders, float *image_out);

/* Render, ACCUMULATING into a caller-provided DEVICE tensor d_rgbw (height*width*4 floats, on the context's
 * device) on CUDA stream `stream` (a cudaStream_t, NULL = default stream). Asynchronous. The caller zeroes the
 * tensor, reduces it across GPUs (ncclAllReduce sum) and calls dtof_develop_device. */
dtof_status dtof_render_device(dtof_ctx *ctx, const dtof_params *params, float *d_rgbw, void *stream);

/* d_image[h*w*3] = d_rgbw[...,0:3] / d_rgbw[...,3] (W == 0 -> divide by 1), asynchronous on `stream`. */
dtof_status dtof_develop_device(dtof_ctx *ctx, const float *d_rgbw, float *d_image, void *stream);

/* Evaluate n wavefront lanes of pass 0 (host arrays in, host records out). */
dtof_status dtof_trace_samples(dtof_ctx *ctx, const dtof_params *params, const uint64_t *lanes, uint32_t n,
                               dtof_sample_record *out);

/* The same for the lanes' samples of pass `pass` (0 <= pass < n_passes): the sampler streams of a lane continue from
 * pass to pass (src/render/integrator.cpp:299-308, Sampler::advance / current_sample_index src/render/sampler.cpp:52-55,
 * 94-103), so passes 0 .. pass - 1 of each lane are replayed first and the record describes the last one. */
dtof_status dtof_trace_samples_pass(dtof_ctx *ctx, const dtof_params *params, const uint64_t *lanes, uint32_t n,
                                    uint32_t pass, dtof_sample_record *out);

/* Scene::ray_intersect_preliminary / Scene::ray_test (src/render/scene.cpp:125-154) for caller-supplied rays, with the
 * per-ray time of the motion-blurred instances: closest hit (any_hit == 0; ties resolve to the lowest triangle id) or any
 * hit (any_hit != 0; only `hit` is meaningful). Host arrays in and out. `nodes_visited` / `tris_tested` count the BVH
 * walk of that ray (a traversal-quality probe: an axis-parallel ray must not cost more than its neighbours). */
typedef struct dtof_ray {
    float o[3], tmax;
    float d[3], time;
} dtof_ray;
typedef struct dtof_ray_hit {
    float t, u, v;            /* distance and barycentric coordinates (b1, b2) of the hit */
    uint32_t prim;            /* global triangle id = position in the order of the scene description */
    int32_t instance;         /* animated instance the triangle belongs to, -1 = static geometry */
    uint32_t hit;             /* 1 = hit / occluded */
    uint32_t nodes_visited, tris_tested;
} dtof_ray_hit;
dtof_status dtof_trace_rays(dtof_ctx *ctx, const dtof_ray *rays, uint32_t n, int any_hit, dtof_ray_hit *out);

/* Enable/disable traversal counters (slower kernel variant) and read them back after a render. */
dtof_status dtof_set_stats(dtof_ctx *ctx, int enabled);
dtof_status dtof_get_stats(dtof_ctx *ctx, dtof_stats *out);

/* Traversal variant the last render kernel ran with: 0 = BVH read from HBM through L1/L2, 1 = BVH staged in shared
 * memory, 2 = flat warp-coherent walk over all triangles in shared memory (tiny scenes); -1 = nothing rendered yet. */
int dtof_last_traversal_mode(const dtof_ctx *ctx);

/* Pipeline the last render ran through: 0 = fused kernel (one lane owns one path in registers), 1 = wavefront
 * pipeline (generate / trace with dynamic ray fetch / shade / shadow / splat kernels exchanging compacted queues
 * through HBM; chosen when the BVH is walked from HBM). The environment variable DTOF_WAVEFRONT=0|1 overrides the
 * choice; per-lane results are identical in both. -1 = nothing rendered yet. */
int dtof_last_pipeline(const dtof_ctx *ctx);

/* Number of kernels this library launched on this context since creation (bench.py's gpu_launches). */
uint64_t dtof_launch_count(const dtof_ctx *ctx);

/* Device time (ms, CUDA events on the render stream) of the main render kernel(s) of the last
 * dtof_render / dtof_render_device call; synchronises the stream. */
dtof_status dtof_last_kernel_ms(dtof_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* DTOF_H */
