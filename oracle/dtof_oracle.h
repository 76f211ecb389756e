/*
 * TEST INFRASTRUCTURE -- CPU oracle for the dopplertofpath / correlated-sampler hot path.
 *
 * A plain scalar C++ restatement of the reference's algorithm (JIT-variant semantics,
 * SURVEY.md Appendix A), each function citing the reference file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product (libdtof_b200.so) never links or calls it.
 *
 * Parity status: PINNED. (1) integer streams and waveform/warp/sincos values against vectors
 * generated from the reference's own headers (tests/golden/header_vectors.json); (2) per-lane
 * radiance against the reference's compiled scalar_rgb integrator + Embree driven with
 * JIT-identical sample streams (oracle/ref_harness/replay_harness.cpp -> tests/golden/lanes_*.txt);
 * (3) the reference's only committed artefact configs_example/scene.exr (tests/golden/scene_exr.npy).
 * Not pinned by a reference run: pass >= 1 of multi-pass renders (the scalar reference exits the
 * bounce loop early and consumes fewer draws than the JIT variants, dopplertofpath.cpp:171-174).
 *
 * Scene / parameter structs are the public C ABI ones (include/dtof.h), so tests feed the CUDA
 * path and the oracle byte-identical inputs.
 */
#ifndef DTOF_ORACLE_H
#define DTOF_ORACLE_H

#include "../include/dtof.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- unit hooks (integer / scalar functions, pinned by header vectors) ---- */
void dtof_oracle_tea32(uint32_t v0, uint32_t v1, int rounds, uint32_t out[2]);
/* seeds PCG32 exactly like dr::PCG32::seed(1, initstate, initseq) and returns n uint32 draws */
void dtof_oracle_pcg32(uint64_t initstate, uint64_t initseq, uint32_t n, uint32_t *out_u32, float *out_f32,
                       uint64_t *state_inc_after_seed /* [2] */);
uint32_t dtof_oracle_permute_kensler(uint32_t index, uint32_t sample_count, uint32_t seed);
void dtof_oracle_sincos(float x, float *s, float *c);
/* ReconstructionFilter::eval of the film's filter (src/rfilters/*.cpp); box filters are not evaluated by the film */
float dtof_oracle_rfilter_eval(const dtof_film *film, float x);
float dtof_oracle_waveform_lowpass(float t, uint32_t type);
float dtof_oracle_waveform(float t, uint32_t type);
float dtof_oracle_modulation_weight(const dtof_params *p, float ray_time, float path_length);
void dtof_oracle_square_to_cosine_hemisphere(float u, float v, float out[3]);
void dtof_oracle_coordinate_system(const float n[3], float s[3], float t[3]);
/* lane seeding + first draws: out = {rng.state, rng.inc, time.state, time.inc, path.state, path.inc, perm_seed} */
void dtof_oracle_seed_lane(const dtof_params *p, uint64_t idx, uint32_t spp_per_pass, uint64_t out[7]);
/* time sample of lane idx in pass `pass` given fresh streams (advances nothing outside) */
float dtof_oracle_time_sample(const dtof_params *p, uint64_t idx, uint32_t spp_per_pass, uint32_t pass);
/* camera ray for film-space uv in [0,1]^2 */
void dtof_oracle_camera_ray(const dtof_camera *cam, float u, float v, float o[3], float d[3], float *maxt);
/* film splat of one sample (sample_pos, rgbw) into a double accumulation tensor h*w*4 */
void dtof_oracle_film_put(const dtof_film *film, float px, float py, const float rgbw[4], double *accum);

/* pass split, src/render/integrator.cpp:121-134,227-245; returns nonzero where the reference throws */
int dtof_oracle_pass_info(const dtof_scene_desc *scene, const dtof_params *p, dtof_pass_info *out);

/* ---- the path ---- */
typedef struct dtof_oracle_scene dtof_oracle_scene;
dtof_oracle_scene *dtof_oracle_scene_create(const dtof_scene_desc *scene, int use_bvh /* 0 brute force, 1 BVH, -1 auto */);
void dtof_oracle_scene_destroy(dtof_oracle_scene *s);

/* per-lane evaluation of pass 0 (pass > 0 lanes are replayed from pass 0 internally when pass_out >= 1) */
int dtof_oracle_trace_samples(const dtof_oracle_scene *s, const dtof_params *p, const uint64_t *lanes, uint32_t n,
                              dtof_sample_record *out);
/* Scene::ray_intersect_preliminary / ray_test (src/render/scene.cpp:125-154) for caller-supplied rays; prim = global
 * triangle id in scene order, instance = index of the animated instance or -1 */
int dtof_oracle_trace_rays(const dtof_oracle_scene *s, const dtof_ray *rays, uint32_t n, int any_hit, dtof_ray_hit *out);
/* the lanes' samples of pass `pass` (earlier passes replayed; streams persist, integrator.cpp:299-308) */
int dtof_oracle_trace_samples_pass(const dtof_oracle_scene *s, const dtof_params *p, const uint64_t *lanes, uint32_t n,
                                   uint32_t pass, dtof_sample_record *out);

/* full render: rgbw_out[h*w*4] (accumulated in double, rounded to float at the end), image_out[h*w*3] or NULL.
 * n_threads <= 0 -> all hardware threads. lane range from p->lane_begin/lane_end. */
int dtof_oracle_render(const dtof_oracle_scene *s, const dtof_params *p, int n_threads, float *rgbw_out,
                       float *image_out);

/* traversal work counters accumulated by the last dtof_oracle_render on this scene */
void dtof_oracle_get_stats(const dtof_oracle_scene *s, dtof_stats *out);

#ifdef __cplusplus
}
#endif
#endif
