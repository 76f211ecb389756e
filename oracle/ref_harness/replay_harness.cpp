// TEST INFRASTRUCTURE (fixture generator) -- not part of the product path.
//
// Drives the REFERENCE's own compiled scalar_rgb code (libmitsuba.so + plugins, built from
// /root/reference by the survey recipe; runtime copied to oracle/_ref/)
// with the random streams of the reference's JIT variants, so that
// `DopplerToFPathIntegrator::sample()` (src/integrators/dopplertofpath.cpp:79-283), the
// perspective sensor, Embree and the BSDF/emitter plugins produce the per-lane radiance that
// `llvm_rgb` would produce for wavefront lane `idx`.
//
// What is restated here (and only here) is the *sampler*: the JIT branch of
//   PCG32Sampler::seed            src/render/sampler.cpp:115-134
//   CorrelatedSampler::seed       src/samplers/correlated.cpp:38-64
//   compute_per_sequence_seed     src/render/sampler.cpp:85-92
//   current_sample_index          src/render/sampler.cpp:94-103
//   next_1d_time                  src/samplers/correlated.cpp:92-153
//   next_{1,2}d_correlate         src/samplers/correlated.cpp:156-167
// using the reference's own PCG32 / sample_tea_32 / permute_kensler templates, plus the lane ->
// pixel mapping and jitter/time prologue of render()/render_sample()
// (src/render/integrator.cpp:273-290,476-509).
//
// Passes >= 1 (the streams of a lane persist across passes, integrator.cpp:299-308): the JIT variants consume the six
// values of a loop iteration [emitter 2D, bsdf 1D, bsdf 2D, roulette 1D] whenever the iteration is ENTERED (the sampler
// calls carry the loop mask `active`, dopplertofpath.cpp:187-270), the scalar build leaves the loop before any of them at
// `if (dr::none_or<false>(active_next)) break;` (:173-174). Every entered iteration performs exactly one closest-hit
// query, so this file interposes Embree's rtcIntersect1 (an undefined dynamic symbol of libmitsuba.so), counts the
// queries of one sample() call and, when one more iteration was entered than was completed, burns the six values the
// JIT variants would have drawn. With that the later passes of a lane see the JIT variants' streams.
//
// Output: one text line per (lane, pass):
//   idx pass px py sample_pos.x sample_pos.y time ray.o(3) ray.d(3) ray.maxt R G B
// floats printed as %.9g (round-trip exact for float32).
//
// Build/run: see oracle/ref_harness/Makefile and tests/golden/make_golden.py.

#define protected public   // read SamplingIntegrator's Doppler members (m_time_sampling_method, ...)
#include <mitsuba/core/fwd.h>
#include <mitsuba/core/argparser.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/filesystem.h>
#include <mitsuba/core/fresolver.h>
#include <mitsuba/core/jit.h>
#include <mitsuba/core/logger.h>
#include <mitsuba/core/profiler.h>
#include <mitsuba/core/random.h>
#include <mitsuba/core/thread.h>
#include <mitsuba/core/xml.h>
#include <mitsuba/render/film.h>
#include <mitsuba/render/integrator.h>
#include <mitsuba/render/sampler.h>
#include <mitsuba/render/scene.h>
#include <mitsuba/render/sensor.h>
#undef protected

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

namespace mi = mitsuba;
namespace dr = drjit;

// closest-hit queries since the last reset (see the header comment); the real function is Embree's
static int g_closest_queries = 0;
struct RTCSceneTy;
struct RTCIntersectContext;
struct RTCRayHit;
extern "C" __attribute__((visibility("default"))) void rtcIntersect1(RTCSceneTy *scene, RTCIntersectContext *context,
                                                                    RTCRayHit *rayhit) {
    using Fn = void (*)(RTCSceneTy *, RTCIntersectContext *, RTCRayHit *);
    static Fn real = (Fn) dlsym(RTLD_NEXT, "rtcIntersect1");
    ++g_closest_queries;
    real(scene, context, rayhit);
}

using F = float;
using S = mi::Color<float, 3>;

// Sampler that replays the JIT-variant streams of CorrelatedSampler for one wavefront lane.
class ReplaySampler final : public mi::Sampler<F, S> {
public:
    using Base = mi::Sampler<F, S>;
    using PCG = mi::PCG32<uint32_t>;
    using Point2f = mi::Point<float, 2>;

    ReplaySampler(uint32_t base_seed, uint32_t sample_count, uint32_t tcn, uint32_t pcn)
        : Base(mi::Properties()), m_tcn(tcn), m_pcn(pcn) {
        m_base_seed = base_seed;
        m_sample_count = sample_count;
    }

    // (re)seed for lane idx: JIT branch of sampler.cpp:115-134 + correlated.cpp:38-64
    void seed_lane(uint32_t seed, uint32_t idx, uint32_t spp_per_pass) {
        uint32_t S_ = m_base_seed + seed;
        auto [v0, v1] = mi::sample_tea_32<uint32_t>(S_, idx);
        auto [t0, t1] = mi::sample_tea_32<uint32_t>(S_ + 1, idx / m_tcn);
        auto [p0, p1] = mi::sample_tea_32<uint32_t>(S_ + 2, idx / m_pcn);
        m_rng.seed(1, v0, v1);
        m_rng_time.seed(1, t0, t1);
        m_rng_path.seed(1, p0, p1);
        // sampler.cpp:85-92
        uint32_t sequence_idx = spp_per_pass * (idx / spp_per_pass);
        m_perm_seed = mi::sample_tea_32<uint32_t>(m_base_seed, sequence_idx + seed).first;
        m_spp_pp = spp_per_pass;
        m_idx = idx;
        m_pass = 0;
        m_dim = 0;
    }
    void advance() override { m_pass++; m_dim = 0; }          // sampler.cpp:52-55
    uint32_t sample_index() const {                             // sampler.cpp:94-103
        uint32_t off = m_spp_pp > 1 ? m_idx % m_spp_pp : 0;
        return m_pass * m_spp_pp + off;
    }

    mi::ref<Base> fork() override { return this; }
    mi::ref<Base> clone() override { return this; }
    void seed(uint32_t, uint32_t) override {}

    // ---- JIT-faithful draw consumption inside sample() ------------------------------------------------------------
    // Every iteration of the path loop draws, in the JIT variants, [emitter 2D, bsdf 1D, bsdf 2D, roulette 1D]
    // (dopplertofpath.cpp:187-210,262-270; path.cpp likewise). The SCALAR build evaluates `dr::any_or<true>(active_em)`
    // on the lane's real mask and skips the emitter 2D draw on a surface without a smooth lobe (a mirror), which the
    // JIT variants never do. begin_path() arms a small state machine that notices the missing 2D draw (a 1D request
    // arrives where the emitter 2D was expected) and burns the two values the JIT variants would have consumed.
    enum Expect { EM_2D, BSDF_1D, BSDF_2D, RR_1D, FREE };
    void begin_path() { m_expect = EM_2D; m_completed = 0; }
    // `entered` loop iterations (= closest-hit queries) against the completed ones: the JIT variants draw the six values
    // of an iteration the scalar build left at its early `break`
    void end_path(int entered, bool correlated) {
        m_expect = FREE;
        if (entered < 0)
            return;
        if (entered != m_completed && entered != m_completed + 1) {
            fprintf(stderr, "replay_harness: %d iterations entered, %d completed\n", entered, m_completed);
            exit(3);
        }
        if (entered == m_completed + 1)
            for (int i = 0; i < 6; ++i) {
                if (correlated)
                    raw_1d_correlate(false);
                else
                    raw_1d();
            }
    }
    float raw_1d() { return m_rng.next_float32(); }
    float raw_1d_correlate(bool correlate) {
        float r1 = m_rng_path.next_float32();
        float r2 = m_rng.next_float32();
        return correlate ? r1 : r2;
    }
    template <bool TWO> void track(bool correlated) {
        if (m_expect == FREE)
            return;
        if (!TWO && m_expect == EM_2D) {   // the scalar branch skipped the emitter sample: realign with the JIT streams
            for (int i = 0; i < 2; ++i) {
                if (correlated)
                    raw_1d_correlate(false);
                else
                    raw_1d();
            }
            m_expect = BSDF_1D;
        }
        if (m_expect == RR_1D)
            ++m_completed;
        m_expect = m_expect == EM_2D ? BSDF_1D : m_expect == BSDF_1D ? BSDF_2D : m_expect == BSDF_2D ? RR_1D : EM_2D;
    }

    float next_1d(bool = true) override {
        track<false>(false);
        return raw_1d();
    }
    Point2f next_2d(bool = true) override {
        track<true>(false);
        float a = raw_1d(), b = raw_1d();
        return Point2f(a, b);
    }
    float next_1d_correlate(bool = true, bool correlate = false) override {
        track<false>(true);
        return raw_1d_correlate(correlate);
    }
    Point2f next_2d_correlate(bool = true, bool correlate = false) override {
        track<true>(true);
        float a = raw_1d_correlate(correlate), b = raw_1d_correlate(correlate);
        return Point2f(a, b);
    }
    // correlated.cpp:92-153 (scalar restatement, same statement order)
    float next_1d_time(bool = true, mi::ETimeSampling strategy = mi::ETimeSampling::TIME_SAMPLING_UNIFORM,
                       float antithetic_shift = 0.f, bool strat = false) override {
        if (strategy == mi::TIME_SAMPLING_UNIFORM)
            return m_rng.next_float32();
        uint32_t si = sample_index();
        float r;
        if (strategy == mi::TIME_SAMPLING_STRATIFIED)
            r = m_rng.next_float32();
        else
            r = m_rng_time.next_float32();
        if (strat) {
            int n_stratum = m_sample_count / m_tcn;
            if (strategy == mi::TIME_SAMPLING_STRATIFIED) {
                uint32_t ps = m_perm_seed + m_dim++;
                uint32_t p1 = mi::permute_kensler<uint32_t>(si / m_tcn, n_stratum, ps);
                ps = m_perm_seed + m_dim++;
                uint32_t p2 = mi::permute_kensler<uint32_t>(si / m_tcn, n_stratum, ps);
                uint32_t p = (si % m_tcn != 0) ? p1 : p2;
                r = (p + r) / n_stratum;
            } else {
                uint32_t p = si / m_tcn;
                r = (p + r) / n_stratum;
            }
        }
        if (strategy == mi::TIME_SAMPLING_STRATIFIED) {
            uint32_t p = si % m_tcn;
            return (p + r) * dr::rcp(float(m_tcn));
        } else if (strategy == mi::TIME_SAMPLING_ANTITHETIC) {
            uint32_t rem = si % m_tcn;
            if (m_tcn == 2)
                return rem != 1 ? r : r + antithetic_shift;
            return r + float(rem) / float(m_tcn);
        } else if (strategy == mi::TIME_SAMPLING_ANTITHETIC_MIRROR) {
            uint32_t rem = si % m_tcn;
            float r2 = 1.0f - r + antithetic_shift;
            return rem != 1 ? r : r2;
        } else if (strategy == mi::TIME_SAMPLING_PERIODIC) {
            uint32_t rem = si % m_tcn;
            return r + float(rem) / float(m_tcn);
        }
        return r;
    }

    const mi::Class *class_() const override { return s_class; }
    static mi::Class *s_class;

private:
    PCG m_rng, m_rng_time, m_rng_path;
    uint32_t m_tcn, m_pcn, m_perm_seed = 0, m_spp_pp = 1, m_idx = 0, m_pass = 0, m_dim = 0;
    Expect m_expect = FREE;
    int m_completed = 0;
};
mi::Class *ReplaySampler::s_class = new mi::Class("ReplaySampler", "Sampler", "scalar_rgb", nullptr, nullptr);

static void usage() {
    fprintf(stderr,
            "usage: replay_harness scene.xml --seed S --tcn N --pcn N --lanes FILE [--spp N] [-Dk=v ...]\n"
            "  FILE: one wavefront lane index per line\n");
    exit(2);
}

int main(int argc, char **argv) {
    std::string scene_path, lanes_path;
    uint32_t seed = 0, tcn = 2, pcn = 0, spp_override = 0;
    mi::xml::ParameterList params;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--seed") seed = (uint32_t) atoll(argv[++i]);
        else if (a == "--tcn") tcn = (uint32_t) atoll(argv[++i]);
        else if (a == "--pcn") pcn = (uint32_t) atoll(argv[++i]);
        else if (a == "--spp") spp_override = (uint32_t) atoll(argv[++i]);
        else if (a == "--lanes") lanes_path = argv[++i];
        else if (a.rfind("-D", 0) == 0) {
            auto eq = a.find('=');
            params.emplace_back(a.substr(2, eq - 2), a.substr(eq + 1), false);
        } else scene_path = a;
    }
    if (scene_path.empty() || lanes_path.empty()) usage();
    if (pcn == 0) pcn = tcn;

    mi::Jit::static_initialization();
    mi::Class::static_initialization();
    mi::Thread::static_initialization();
    mi::Logger::static_initialization();
    mi::Bitmap::static_initialization();
    mi::Profiler::static_initialization();
    mi::Thread::thread()->logger()->set_log_level(mi::Warn);

    using Scene = mi::Scene<F, S>;
    using Integ = mi::SamplingIntegrator<F, S>;
    using Sensor = mi::Sensor<F, S>;
    using Film = mi::Film<F, S>;
    using Vector2f = mi::Vector<float, 2>;
    using Point2f = mi::Point<float, 2>;

    mi::color_management_static_initialization(false, false);
    Scene::static_accel_initialization();

    {
        mi::ref<mi::FileResolver> fr = mi::Thread::thread()->file_resolver();
        mi::fs::path exe_dir = mi::fs::path(argv[0]).parent_path();
        const char *ref_dir = getenv("DTOF_REF_DIR");
        fr->append(ref_dir ? mi::fs::path(ref_dir) : exe_dir);
        fr->append(mi::fs::path(scene_path).parent_path());

        std::vector<mi::ref<mi::Object>> parsed =
            mi::xml::load_file(scene_path, "scalar_rgb", params, false, false);
        Scene *scene = dynamic_cast<Scene *>(parsed[0].get());
        if (!scene) { fprintf(stderr, "not a scene\n"); return 1; }
        Integ *integ = dynamic_cast<Integ *>(scene->integrator());
        if (!integ) { fprintf(stderr, "not a SamplingIntegrator\n"); return 1; }
        Sensor *sensor = scene->sensors()[0].get();
        Film *film = sensor->film();

        // render() prologue, integrator.cpp:115-134,227-245
        mi::Vector<uint32_t, 2> film_size = film->crop_size();
        auto crop_offset = film->crop_offset();
        if (spp_override) sensor->sampler()->set_sample_count(spp_override);
        uint32_t spp = sensor->sampler()->sample_count();
        uint32_t spp_per_pass = spp, n_passes = 1;
        size_t wavefront = (size_t) film_size.x() * film_size.y() * spp_per_pass, limit = 0xffffffffu;
        if (wavefront > limit) {
            spp_per_pass /= (uint32_t) ((wavefront + limit - 1) / limit);
            n_passes = spp / spp_per_pass;
        }
        if (spp % spp_per_pass != 0) { fprintf(stderr, "sample_count %% spp_per_pass != 0\n"); return 1; }

        mi::ref<ReplaySampler> sampler = new ReplaySampler(0 /* base_seed */, spp, tcn, pcn);

        bool correlate_pixel = integ->m_path_correlation_depth > 0;
        bool box = film->rfilter()->is_box_filter();
        Vector2f scale = 1.f / Vector2f(film->crop_size()),
                 offset = -Vector2f(film->crop_offset()) * scale;

        printf("# scene=%s seed=%u spp=%u spp_per_pass=%u n_passes=%u tcn=%u pcn=%u W=%u H=%u box=%d\n",
               scene_path.c_str(), seed, spp, spp_per_pass, n_passes, tcn, pcn, film_size.x(), film_size.y(), (int) box);

        std::ifstream lf(lanes_path);
        uint64_t idx64;
        float aovs[8];
        while (lf >> idx64) {
            uint32_t idx = (uint32_t) idx64;
            sampler->seed_lane(seed, idx, spp_per_pass);
            uint32_t pixel = idx / spp_per_pass;
            uint32_t py = pixel / film_size.x(), px = pixel - py * film_size.x();
            Vector2f pos((float) (px + crop_offset.x()), (float) (py + crop_offset.y()));
            for (uint32_t pass = 0; pass < n_passes; ++pass) {
                // render_sample() Doppler branch, integrator.cpp:476-509
                // ... or the stock branch (:409-472) for integrators that are not Doppler integrators (path, velocity)
                const bool doppler = integ->m_is_doppler_integrator;
                Vector2f sample_pos = pos + (doppler ? Vector2f(sampler->next_2d_correlate(true, correlate_pixel))
                                                     : Vector2f(sampler->next_2d(true))),
                         adjusted = dr::fmadd(sample_pos, scale, offset);
                float time = sensor->shutter_open();
                if (sensor->shutter_open_time() > 0.f)
                    time += (doppler ? sampler->next_1d_time(true, integ->m_time_sampling_method, integ->m_antithetic_shift,
                                                             integ->m_use_stratified_sampling_for_each_interval)
                                     : sampler->next_1d(true)) *
                            sensor->shutter_open_time();
                auto [ray, ray_weight] = sensor->sample_ray_differential(time, 0.f, Point2f(adjusted), Point2f(.5f));
                if (getenv("DTOF_DEBUG")) {
                    // debug aid: primary surface interaction + emitter sample as the reference computes them
                    mi::Ray<mi::Point<float, 3>, S> r2(ray);
                    r2.time = time < 0.0015f ? time : time - 0.0015f;
                    auto si = scene->ray_intersect(r2, +mi::RayFlags::All, true);
                    fprintf(stderr, "  si.t=%.9g p=(%.9g %.9g %.9g) n=(%.9g %.9g %.9g) shn=(%.9g %.9g %.9g) s=(%.9g %.9g %.9g) wi=(%.9g %.9g %.9g)\n",
                            si.t, si.p.x(), si.p.y(), si.p.z(), si.n.x(), si.n.y(), si.n.z(), si.sh_frame.n.x(),
                            si.sh_frame.n.y(), si.sh_frame.n.z(), si.sh_frame.s.x(), si.sh_frame.s.y(), si.sh_frame.s.z(),
                            si.wi.x(), si.wi.y(), si.wi.z());
                    auto [ds, w] = scene->sample_emitter_direction(si, Point2f(0.3f, 0.6f), true, true);
                    fprintf(stderr, "  ds.p=(%.9g %.9g %.9g) d=(%.9g %.9g %.9g) dist=%.9g pdf=%.9g w=(%.9g %.9g %.9g)\n", ds.p.x(),
                            ds.p.y(), ds.p.z(), ds.d.x(), ds.d.y(), ds.d.z(), ds.dist, ds.pdf, w.x(), w.y(), w.z());
                }
                sampler->begin_path();
                g_closest_queries = 0;
                auto [spec, valid] = integ->sample(scene, sampler.get(), ray, nullptr, aovs, true);
                // the velocity integrator has no path loop (two queries, no draws)
                const bool has_loop = integ->class_()->name() != "VelocityIntegrator";
                sampler->end_path(has_loop ? g_closest_queries : -1, doppler);
                S rgb = ray_weight * spec;
                printf("%u %u %u %u %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", idx, pass, px, py,
                       sample_pos.x(), sample_pos.y(), time, ray.o.x(), ray.o.y(), ray.o.z(), ray.d.x(), ray.d.y(),
                       ray.d.z(), ray.maxt, rgb.x(), rgb.y(), rgb.z());
                sampler->advance();
            }
        }
    }
    fflush(stdout);
    _Exit(0); // skip static shutdown ordering issues; fixtures are already flushed
}
