// Mitsuba-side binding of libdtof_b200.so: the plugin a maintainer of juhyeonkim95/Mitsuba3DopplerToF adds as
// src/integrators/dopplertofpath_b200.cpp (plugin file "dopplertofpath_b200.so"; the SAME source installed as
// plugins/dopplertofpath.so answers <integrator type="dopplertofpath"> of an untouched scene file: the plugin manager
// goes by the file name, src/core/plugin.cpp:93-127 -- oracle/ref_harness/Makefile builds that drop-in view too).
// It keeps the reference's property surface, walks the loaded mitsuba::Scene, hands plain arrays to the C ABI
// (include/dtof.h) and puts the returned RGBW tensor into the film, i.e. it replaces
//   SamplingIntegrator::render            src/render/integrator.cpp:104-347
// for this integrator. Built for scalar_rgb (Float = float, host memory), with the reference's exact compile flags
// (INTEGRATION.md). It compiles against the UNMODIFIED reference headers and loads into the unmodified reference
// runtime (`mitsuba -m scalar_rgb scene.xml`): oracle/ref_harness/Makefile builds it next to a reference runtime in
// oracle/_ref/plugins/, tests/test_mitsuba_plugin.py renders through it on the GPU box.
//
// Two pieces of state are private in the reference (declared inside .cpp files, no accessor, not exposed through
// traverse()): Instance::m_transform (src/shapes/instance.cpp:338-340) and CorrelatedSampler's correlate numbers
// (src/samplers/correlated.cpp:179-183). They are read through layout mirrors of those two classes (below), checked
// against the class name; upstream would rather add the accessors listed in INTEGRATION.md section 2.
#include <cmath>
#include <cstdlib>

#include <mitsuba/core/properties.h>
#include <mitsuba/core/transform.h>
#include <mitsuba/render/bsdf.h>
#include <mitsuba/render/emitter.h>
#include <mitsuba/render/film.h>
#include <mitsuba/render/imageblock.h>
#include <mitsuba/render/integrator.h>
#include <mitsuba/render/mesh.h>
#include <mitsuba/render/sampler.h>
#include <mitsuba/render/scene.h>
#include <mitsuba/render/sensor.h>

#include "dtof.h"

NAMESPACE_BEGIN(mitsuba)

namespace {
// Collects the children / parameters an object exposes through traverse() (the only public window onto
// Rectangle::m_to_world, ShapeGroup::m_shapes, TwoSidedBRDF::m_brdf, PointLight::m_position ...).
struct Collector : TraversalCallback {
    std::vector<std::pair<std::string, Object *>> objects;
    std::vector<std::pair<std::string, void *>> params;
    void put_parameter_impl(const std::string &name, void *ptr, uint32_t, const std::type_info &) override {
        params.emplace_back(name, ptr);
    }
    void put_object(const std::string &name, Object *obj, uint32_t) override { objects.emplace_back(name, obj); }
    template <typename T> T *param(const char *name) {
        for (auto &p : params)
            if (p.first == name)
                return (T *) p.second;
        return nullptr;
    }
};
} // namespace

// Layout mirrors (same base class, same member types in the same order => same offsets under the Itanium ABI).
// Never instantiated; only used to read two members of objects the reference's own plugins constructed.
template <typename Float, typename Spectrum>
struct InstanceLayout : public Shape<Float, Spectrum> {                  // src/shapes/instance.cpp:52-341
    ref<Object> m_shapegroup;
    ref<AnimatedTransform> m_transform;
};
template <typename Float, typename Spectrum>
struct CorrelatedSamplerLayout : public PCG32Sampler<Float, Spectrum> {  // src/samplers/correlated.cpp:13-187
    using PCG32 = mitsuba::PCG32<dr::uint32_array_t<Float>>;
    PCG32 m_rng_time;
    int m_time_correlate_number;
    PCG32 m_rng_path;
    int m_path_correlate_number;
    dr::uint32_array_t<Float> m_permutation_seed;
};

// Protected members declared in the reference's headers, read through a pointer to member named from a derived class
// (Sampler::m_base_seed include/mitsuba/render/sampler.h:184, Mesh::m_flip_normals include/mitsuba/render/mesh.h:456).
template <typename Float, typename Spectrum>
struct SamplerPeek : public Sampler<Float, Spectrum> {
    static uint32_t base_seed(const Sampler<Float, Spectrum> *s) { return s->*(&SamplerPeek::m_base_seed); }
};
template <typename Float, typename Spectrum>
struct MeshPeek : public Mesh<Float, Spectrum> {
    static bool flip_normals(const Mesh<Float, Spectrum> *m) { return m->*(&MeshPeek::m_flip_normals); }
};

template <typename Float, typename Spectrum>
class DopplerToFPathB200 final : public MonteCarloIntegrator<Float, Spectrum> {
public:
    MI_IMPORT_BASE(MonteCarloIntegrator, m_max_depth, m_rr_depth, m_hide_emitters, m_time_sampling_method, m_antithetic_shift,
                   m_use_stratified_sampling_for_each_interval, m_path_correlation_depth, m_is_doppler_integrator, m_stop,
                   m_render_timer, should_stop)
    MI_IMPORT_TYPES(Scene, Sensor, Film, ImageBlock, Sampler, Shape, Mesh, BSDF, Emitter, Medium)

    DopplerToFPathB200(const Properties &props) : Base(props) {
        if constexpr (dr::is_jit_v<Float> || !is_rgb_v<Spectrum>)
            Throw("dopplertofpath_b200 is a scalar_rgb host plugin: the GPU work happens behind the C ABI");
        m_is_doppler_integrator = true;
        // same reads, defaults and syntactic sugar as DopplerToFPathIntegrator (src/integrators/dopplertofpath.cpp:19-57)
        m_p = dtof_params{};
        m_p.time = props.get<ScalarFloat>("time", 0.0015f);
        m_p.w_g = props.get<ScalarFloat>("w_g", 30.f);
        m_p.g_1 = props.get<ScalarFloat>("g_1", 0.5f);
        m_p.g_0 = props.get<ScalarFloat>("g_0", 0.5f);
        ScalarFloat w_s = props.get<ScalarFloat>("w_s", 30.f);
        m_p.sensor_phase_offset = props.get<ScalarFloat>("sensor_phase_offset", 0.f);
        if (props.has_property("hetero_offset"))
            m_p.sensor_phase_offset = props.get<ScalarFloat>("hetero_offset", 0.0) * 2 * 3.14159265358979323846;
        if (props.has_property("hetero_frequency"))
            m_p.hetero_frequency = props.get<ScalarFloat>("hetero_frequency", 1.0);
        else
            m_p.hetero_frequency = (w_s - m_p.w_g) * 1e6 * m_p.time;
        std::string wave = props.get<std::string>("wave_function_type", "sinusoidal");
        if (wave == "sinusoidal") m_p.wave_function_type = DTOF_WAVE_SINUSOIDAL;
        else if (wave == "rectangular") m_p.wave_function_type = DTOF_WAVE_RECTANGULAR;
        else if (wave == "triangular") m_p.wave_function_type = DTOF_WAVE_TRIANGULAR;
        else if (wave == "trapezoidal") m_p.wave_function_type = DTOF_WAVE_TRAPEZOIDAL;
        else Throw("unknown wave_function_type \"%s\"", wave);
        m_p.low_frequency_component_only = props.get<bool>("low_frequency_component_only", true);
        // `device` (one GPU) or `devices` = "0,1,2,3" / "all": one context over several GPUs of the node; one render is
        // sharded over them behind the C ABI and summed on the first (dtof_create_multi). The environment variable
        // DTOF_DEVICES supplies the list for a scene file that must stay untouched.
        m_device = props.get<int>("device", 0);
        std::string devices = props.get<std::string>("devices", "");
        if (devices.empty() && getenv("DTOF_DEVICES"))
            devices = getenv("DTOF_DEVICES");
        std::vector<int> list;
        if (devices == "all") {
            list.push_back(-1);
        } else {
            for (size_t at = 0; at < devices.size();) {
                size_t end = devices.find(',', at);
                if (end == std::string::npos)
                    end = devices.size();
                list.push_back(std::atoi(devices.substr(at, end - at).c_str()));
                at = end + 1;
            }
        }
        if (list.size() == 1 && list[0] == -1) {   // "all": as many as dtof_create accepts
            list.clear();
            for (int d = 0; d < 16; ++d) {
                dtof_ctx *probe = nullptr;
                if (dtof_create(&probe, d) != DTOF_OK)
                    break;
                dtof_destroy(probe);
                list.push_back(d);
            }
        }
        Log(Info, "dopplertofpath on B200: CUDA library libdtof_b200 (ABI %u) on %u device(s)", dtof_abi_version(),
            (uint32_t) std::max<size_t>(list.size(), 1));
        if (list.size() > 1) {
            if (dtof_create_multi(&m_ctx, list.data(), (uint32_t) list.size()) != DTOF_OK)
                Throw("dtof_create_multi(devices=%s) failed: no usable CUDA devices (there is no CPU fallback)", devices);
        } else {
            if (list.size() == 1)
                m_device = list[0];
            if (dtof_create(&m_ctx, m_device) != DTOF_OK)
                Throw("dtof_create(device=%i) failed: no usable CUDA device (there is no CPU fallback)", m_device);
        }
    }
    ~DopplerToFPathB200() { dtof_destroy(m_ctx); }

    TensorXf render(Scene *scene, Sensor *sensor, uint32_t seed, uint32_t spp, bool develop, bool /*evaluate*/) override {
        m_stop = false;
        m_render_timer.reset();
        Film *film = sensor->film();
        // flatten + BVH build + H2D once per (scene, sensor, film geometry). parameters_changed() on this integrator (what
        // mi.traverse(...).update() ends with) drops the cache: an edited scene is flattened again.
        const ScalarVector2u csize = film->crop_size();
        const ScalarPoint2u coff = film->crop_offset();
        if (m_uploaded != scene || m_uploaded_sensor != sensor || m_uploaded_size != csize || m_uploaded_offset != coff) {
            upload(scene, sensor);
            m_uploaded = scene, m_uploaded_sensor = sensor, m_uploaded_size = csize, m_uploaded_offset = coff;
        }
        dtof_params p = m_p;
        p.max_depth = (int32_t) m_max_depth, p.rr_depth = (int32_t) m_rr_depth, p.hide_emitters = m_hide_emitters;
        p.time_sampling_method = (uint32_t) m_time_sampling_method;      // ETimeSampling order == dtof_time_sampling
        p.antithetic_shift = m_antithetic_shift;
        p.use_stratified_sampling_for_each_interval = m_use_stratified_sampling_for_each_interval;
        p.path_correlation_depth = m_path_correlation_depth;
        const Sampler *sampler = sensor->sampler();
        p.sample_count = spp ? spp : sampler->sample_count();            // integrator.cpp:121-124
        p.base_seed = SamplerPeek<Float, Spectrum>::base_seed(sampler);   // sampler.cpp:13-14
        if (sampler->class_()->name() == "IndependentSampler") {
            // base-class next_1d_time / next_*_correlate (sampler.h:131-144): one independent stream, i.e. the correlated
            // sampler's `rng` stream under uniform time sampling without path correlation (correlated.cpp:92-97, 156-161)
            p.time_sampling_method = DTOF_TIME_UNIFORM;
            p.use_stratified_sampling_for_each_interval = 0;
            p.path_correlation_depth = 0;
            p.time_correlate_number = p.path_correlate_number = 1;
        } else {
            if (sampler->class_()->name() != "CorrelatedSampler")
                Throw("dopplertofpath_b200 needs the 'correlated' or 'independent' sampler (got %s)", sampler->class_()->name());
            auto *cs = reinterpret_cast<const CorrelatedSamplerLayout<Float, Spectrum> *>(sampler);
            p.time_correlate_number = (uint32_t) cs->m_time_correlate_number;   // correlated.cpp:18-19
            p.path_correlate_number = (uint32_t) cs->m_path_correlate_number;
            if (p.time_correlate_number - 1u > 4095u || p.path_correlate_number - 1u > 4095u)
                Throw("implausible correlate numbers (%u, %u): the CorrelatedSampler layout differs from the mirrored one",
                      p.time_correlate_number, p.path_correlate_number);
        }
        p.seed = seed;

        ScalarVector2u size = film->crop_size();
        film->prepare({});                                               // channels R,G,B,W (hdrfilm.cpp:235-279)
        size_t n = (size_t) size.x() * size.y() * 4;
        std::unique_ptr<float[]> rgbw(new float[n]);
        // The wavefront is submitted in chunks of whole pixels (~2^28 lanes, a fraction of a second of GPU time) so that
        // cancel() and the `timeout` property are honoured between them (should_stop(), include/mitsuba/render/
        // integrator.h:106-108); a stopped render keeps what was accumulated so far, like the reference's block loop.
        dtof_pass_info pi;
        if (dtof_pass_info_for(m_ctx, &p, &pi) != DTOF_OK)
            Throw("dtof_pass_info_for: %s", dtof_last_error(m_ctx));
        const uint64_t chunk = std::max<uint64_t>(1, (uint64_t(1) << 28) / pi.spp_per_pass) * pi.spp_per_pass;
        bool first = true;
        // (the first chunk is always submitted: scene upload counts towards the timer, and the reference's block loop has
        // rendered its first blocks too by the time should_stop() turns true)
        for (uint64_t begin = 0; begin < pi.wavefront_size && (first || !should_stop()); begin += chunk, first = false) {
            dtof_params q = p;
            q.lane_begin = begin, q.lane_end = std::min<uint64_t>(pi.wavefront_size, begin + chunk);
            if (dtof_render_accumulate(m_ctx, &q, first) != DTOF_OK)     // H2D params, kernels
                Throw("dtof_render_accumulate: %s", dtof_last_error(m_ctx));
        }
        if (first)                                                       // stopped before the first chunk
            std::fill(rgbw.get(), rgbw.get() + n, 0.f);
        else if (dtof_read_film(m_ctx, rgbw.get(), nullptr) != DTOF_OK)  // D2H film
            Throw("dtof_read_film: %s", dtof_last_error(m_ctx));
        if (should_stop())
            Log(Warn, "Rendering stopped early (cancel / timeout): the film holds the chunks rendered so far.");
        // hand the accumulation tensor to the film exactly as the JIT branch does (integrator.cpp:266,310-323)
        size_t shape[3] = { size.y(), size.x(), 4 };
        TensorXf tensor(dr::load<DynamicBuffer<Float>>(rgbw.get(), n), 3, shape);
        ref<ImageBlock> block = new ImageBlock(tensor, film->crop_offset(), film->rfilter(), /*border*/ false);
        film->put_block(block);
        Log(Info, "Rendering finished. (took %s)", util::time_string((float) m_render_timer.value(), true));
        return develop ? film->develop() : TensorXf();
    }

    // The scalar entry point stays available for callers that sample single rays (integrator.h:200-205):
    // it is the reference's own CPU code and is not on the accelerated path.
    std::pair<Spectrum, Mask> sample(const Scene *, Sampler *, const RayDifferential3f &, const Medium *, Float *,
                                     Mask) const override {
        Throw("dopplertofpath_b200 renders whole images through render(); use 'dopplertofpath' for per-ray sample()");
    }

    void parameters_changed(const std::vector<std::string> & /*keys*/ = {}) override { m_uploaded = nullptr; }

    MI_DECLARE_CLASS()

private:
    // A constant RGB value is all the accelerated path knows: a bitmap / checkerboard / mesh-attribute texture must not
    // be silently evaluated at one point.
    template <typename Tex> static const Tex *uniform(const Tex *t, const char *what) {
        if (t && t->is_spatially_varying())
            Throw("a spatially varying texture (%s of %s) is outside the accelerated path", what, t->class_()->name());
        return t;
    }
    static void m34(const ScalarTransform4f &t, float *out) {
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c)
                out[4 * r + c] = t.matrix(r, c);
    }
    struct MeshBuf {
        std::vector<float> pos, nrm, uv;
        std::vector<uint32_t> idx;
    };
    uint32_t bsdf_index(const BSDF *bsdf, std::vector<dtof_bsdf> &out) {
        dtof_bsdf b{};
        b.kind = DTOF_BSDF_DIFFUSE;
        const BSDF *inner = bsdf;
        if (bsdf->class_()->name() == "TwoSidedBRDF") {
            Collector c;
            const_cast<BSDF *>(bsdf)->traverse(&c);
            if (c.objects.size() != 2 || c.objects[0].second != c.objects[1].second)
                Throw("twosided with two different BRDFs is outside the accelerated path");
            inner = (const BSDF *) c.objects[0].second;
            b.twosided = 1;
        }
        SurfaceInteraction3f si = dr::zeros<SurfaceInteraction3f>();
        if (inner->class_()->name() == "SmoothConductor") {               // eta, k, specular_reflectance (conductor.cpp:232-236)
            b.kind = DTOF_BSDF_CONDUCTOR;
            Collector c;
            const_cast<BSDF *>(inner)->traverse(&c);
            auto tex = [&](const char *name) -> Spectrum {
                for (auto &o : c.objects)
                    if (o.first == name)
                        return uniform((const Texture<Float, Spectrum> *) o.second, name)->eval(si);
                Throw("conductor without '%s'", name);
            };
            Spectrum e = tex("eta"), k = tex("k"), r = tex("specular_reflectance");
            for (int i = 0; i < 3; ++i)
                b.eta[i] = e[i], b.k[i] = k[i], b.reflectance[i] = r[i];
        } else if (inner->class_()->name() == "SmoothDielectric" || inner->class_()->name() == "ThinDielectric") {
            // eta + optional tints (dielectric.cpp:230-237, thindielectric.cpp:128-138)
            b.kind = inner->class_()->name() == "SmoothDielectric" ? DTOF_BSDF_DIELECTRIC : DTOF_BSDF_THINDIELECTRIC;
            Collector c;
            const_cast<BSDF *>(inner)->traverse(&c);
            b.eta[0] = (float) *c.param<ScalarFloat>("eta");
            for (int i = 0; i < 3; ++i)
                b.reflectance[i] = b.k[i] = 1.f;
            for (auto &o : c.objects) {
                Spectrum v = uniform((const Texture<Float, Spectrum> *) o.second, o.first.c_str())->eval(si);
                float *dst = o.first == "specular_reflectance" ? b.reflectance : o.first == "specular_transmittance" ? b.k : nullptr;
                if (dst)
                    dst[0] = v[0], dst[1] = v[1], dst[2] = v[2];
            }
        } else if (inner->class_()->name() == "RoughConductor") {         // roughconductor.cpp:160-224
            b.kind = DTOF_BSDF_ROUGHCONDUCTOR;
            Collector c;
            const_cast<BSDF *>(inner)->traverse(&c);
            const std::string descr = inner->to_string();              // distribution / sample_visible are not traversed
            b.distribution = descr.find("distribution = ggx") != std::string::npos ? 1u : 0u;
            if (descr.find("sample_visible = 0") != std::string::npos)
                Throw("roughconductor with sample_visible=false is outside the accelerated path");
            for (int i = 0; i < 3; ++i)
                b.reflectance[i] = 1.f;
            for (auto &o : c.objects) {
                const Texture<Float, Spectrum> *t = uniform((const Texture<Float, Spectrum> *) o.second, o.first.c_str());
                if (o.first == "alpha")
                    b.alpha[0] = b.alpha[1] = t->eval_1(si);
                else if (o.first == "alpha_u")
                    b.alpha[0] = t->eval_1(si);
                else if (o.first == "alpha_v")
                    b.alpha[1] = t->eval_1(si);
                else {
                    Spectrum v = t->eval(si);
                    float *dst = o.first == "eta" ? b.eta : o.first == "k" ? b.k : o.first == "specular_reflectance" ? b.reflectance : nullptr;
                    if (dst)
                        dst[0] = v[0], dst[1] = v[1], dst[2] = v[2];
                }
            }
        } else if (inner->class_()->name() == "RoughDielectric") {        // roughdielectric.cpp:161-238
            b.kind = DTOF_BSDF_ROUGHDIELECTRIC;
            Collector c;
            const_cast<BSDF *>(inner)->traverse(&c);
            const std::string descr = inner->to_string();              // distribution / sample_visible are not traversed
            b.distribution = descr.find("distribution = ggx") != std::string::npos ? 1u : 0u;
            if (descr.find("sample_visible = 0") != std::string::npos)
                Throw("roughdielectric with sample_visible=false is outside the accelerated path");
            b.eta[0] = (float) *c.param<ScalarFloat>("eta");
            for (int i = 0; i < 3; ++i)
                b.reflectance[i] = b.k[i] = 1.f;
            for (auto &o : c.objects) {
                const Texture<Float, Spectrum> *t = uniform((const Texture<Float, Spectrum> *) o.second, o.first.c_str());
                if (o.first == "alpha")
                    b.alpha[0] = b.alpha[1] = t->eval_1(si);
                else if (o.first == "alpha_u")
                    b.alpha[0] = t->eval_1(si);
                else if (o.first == "alpha_v")
                    b.alpha[1] = t->eval_1(si);
                else {
                    Spectrum v = t->eval(si);
                    float *dst = o.first == "specular_reflectance" ? b.reflectance : o.first == "specular_transmittance" ? b.k : nullptr;
                    if (dst)
                        dst[0] = v[0], dst[1] = v[1], dst[2] = v[2];
                }
            }
        } else if (inner->class_()->name() == "SmoothPlastic") {          // plastic.cpp:157-208
            b.kind = DTOF_BSDF_PLASTIC;
            Collector c;
            const_cast<BSDF *>(inner)->traverse(&c);
            b.eta[0] = (float) *c.param<ScalarFloat>("eta");
            // `nonlinear` is not a traversed parameter; the plugin prints it (to_string, plastic.cpp:368-382)
            b.eta[1] = inner->to_string().find("nonlinear = 1") != std::string::npos ? 1.f : 0.f;
            for (int i = 0; i < 3; ++i)
                b.reflectance[i] = 0.5f, b.k[i] = 1.f;
            for (auto &o : c.objects) {
                Spectrum v = uniform((const Texture<Float, Spectrum> *) o.second, o.first.c_str())->eval(si);
                float *dst = o.first == "diffuse_reflectance" ? b.reflectance : o.first == "specular_reflectance" ? b.k : nullptr;
                if (dst)
                    dst[0] = v[0], dst[1] = v[1], dst[2] = v[2];
            }
        } else {
            if (inner->class_()->name() != "SmoothDiffuse")
                Throw("BSDF \"%s\" is outside the accelerated path (diffuse | conductor | roughconductor | dielectric | thindielectric | roughdielectric | plastic | twosided(...))", inner->class_()->name());
            Collector c;
            const_cast<BSDF *>(inner)->traverse(&c);
            for (auto &o : c.objects)
                uniform((const Texture<Float, Spectrum> *) o.second, o.first.c_str());
            Spectrum r = inner->eval_diffuse_reflectance(si);            // constant RGB reflectance
            b.reflectance[0] = r[0], b.reflectance[1] = r[1], b.reflectance[2] = r[2];
        }
        for (size_t i = 0; i < out.size(); ++i)
            if (!memcmp(&out[i], &b, sizeof(b)))
                return (uint32_t) i;
        out.push_back(b);
        return (uint32_t) out.size() - 1;
    }
    void add_shape(const Shape *shape, std::vector<dtof_mesh> &meshes, std::vector<MeshBuf> &bufs, std::vector<dtof_bsdf> &bsdfs) {
        dtof_mesh m{};
        MeshBuf buf;
        m.emitter = -1;
        m.bsdf = bsdf_index(shape->bsdf(), bsdfs);
        if (shape->is_mesh()) {                                          // Mesh, Cube, PLY, OBJ: world / group space buffers
            const Mesh *mesh = (const Mesh *) shape;
            buf.pos.assign(mesh->vertex_positions_buffer().data(), mesh->vertex_positions_buffer().data() + 3 * mesh->vertex_count());
            buf.idx.assign(mesh->faces_buffer().data(), mesh->faces_buffer().data() + 3 * mesh->face_count());
            if (mesh->has_vertex_normals())
                buf.nrm.assign(mesh->vertex_normals_buffer().data(), mesh->vertex_normals_buffer().data() + 3 * mesh->vertex_count());
            if (mesh->has_vertex_texcoords())
                buf.uv.assign(mesh->vertex_texcoords_buffer().data(), mesh->vertex_texcoords_buffer().data() + 2 * mesh->vertex_count());
            m.kind = DTOF_SHAPE_MESH;
            m.flip_normals = MeshPeek<Float, Spectrum>::flip_normals(mesh);
        } else if (shape->class_()->name() == "Rectangle") {             // analytic quad -> 2 triangles + parametrisation
            Collector c;
            const_cast<Shape *>(shape)->traverse(&c);
            const ScalarTransform4f &tw = *c.param<ScalarTransform4f>("to_world");
            const float corners[4][2] = { { -1, -1 }, { 1, -1 }, { 1, 1 }, { -1, 1 } };
            for (auto &cn : corners) {
                ScalarPoint3f p = tw.transform_affine(ScalarPoint3f(cn[0], cn[1], 0.f));
                buf.pos.insert(buf.pos.end(), { p.x(), p.y(), p.z() });
            }
            buf.uv = { 0, 0, 1, 0, 1, 1, 0, 1 };
            bool ccw = dr::det(ScalarMatrix3f(tw.matrix)) > 0;
            buf.idx = ccw ? std::vector<uint32_t>{ 0, 1, 2, 0, 2, 3 } : std::vector<uint32_t>{ 0, 2, 1, 0, 3, 2 };
            m.kind = DTOF_SHAPE_RECTANGLE;
            m34(tw, m.rect_to_world);
        } else {
            Throw("shape \"%s\" is outside the accelerated path (meshes and rectangles)", shape->class_()->name());
        }
        meshes.push_back(m);
        bufs.push_back(std::move(buf));
    }
    void upload(const Scene *scene, const Sensor *sensor) {
        std::vector<dtof_mesh> meshes;
        std::vector<MeshBuf> bufs;
        std::vector<dtof_bsdf> bsdfs;
        std::vector<dtof_instance> instances;
        std::vector<dtof_emitter> emitters;
        std::vector<std::pair<const Shape *, uint32_t>> mesh_of;
        // static group first, then one instance per animated shape (the XML rewrite already produced
        // shapegroup + instance pairs, src/core/xml.cpp:1166-1192)
        for (auto &s : scene->shapes())
            if (!s->is_instance()) {
                mesh_of.emplace_back(s.get(), (uint32_t) meshes.size());
                add_shape(s.get(), meshes, bufs, bsdfs);
            }
        dtof_instance st{};
        st.n_meshes = (uint32_t) meshes.size();
        const float ident[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
        memcpy(st.m0, ident, sizeof(ident));
        memcpy(st.m1, ident, sizeof(ident));
        if (st.n_meshes)
            instances.push_back(st);
        for (auto &s : scene->shapes())
            if (s->is_instance()) {
                const AnimatedTransform *at =                            // Instance::m_transform, instance.cpp:63
                    reinterpret_cast<const InstanceLayout<Float, Spectrum> *>(s.get())->m_transform.get();
                if (!at || at->size() == 0)
                    Throw("instance without keyframes: the Instance layout differs from the mirrored one");
                dtof_instance in{};
                in.first_mesh = (uint32_t) meshes.size();
                in.animated = 1;
                in.t0 = at->get_min_time(), in.t1 = at->get_max_time();
                m34(at->eval(in.t0), in.m0);                             // Instance::embree_geometry, instance.cpp:295-310
                m34(at->eval(in.t1), in.m1);
                Collector c;                                             // ShapeGroup::traverse lists the members
                ((Shape *) const_cast<Shape *>(s.get())->get_shapegroup())->traverse(&c);
                for (auto &o : c.objects)
                    add_shape((const Shape *) o.second, meshes, bufs, bsdfs);
                in.n_meshes = (uint32_t) meshes.size() - in.first_mesh;
                instances.push_back(in);
            }
        for (auto &e : scene->emitters()) {                              // order = Scene::m_emitters (scene.cpp:40-64)
            dtof_emitter d{};
            SurfaceInteraction3f si = dr::zeros<SurfaceInteraction3f>();
            si.wi = ScalarVector3f(0, 0, 1);
            Collector c;
            const_cast<Emitter *>(e.get())->traverse(&c);
            if (e->class_()->name() == "PointLight") {
                d.kind = DTOF_EMITTER_POINT;
                const ScalarPoint3f &p = *c.param<ScalarPoint3f>("position");
                d.position[0] = p.x(), d.position[1] = p.y(), d.position[2] = p.z();
                Spectrum I = uniform((const Texture<Float, Spectrum> *) c.objects[0].second, "intensity")->eval(si);
                d.value[0] = I[0], d.value[1] = I[1], d.value[2] = I[2];
            } else if (e->class_()->name() == "AreaLight") {
                d.kind = DTOF_EMITTER_AREA;
                for (auto &mo : mesh_of)
                    if (mo.first->emitter() == e.get()) {
                        d.mesh = mo.second;
                        meshes[mo.second].emitter = (int32_t) emitters.size();
                    }
                for (auto &o : c.objects)
                    if (o.first == "radiance")
                        uniform((const Texture<Float, Spectrum> *) o.second, "radiance");
                Spectrum L = e->eval(si);
                d.value[0] = L[0], d.value[1] = L[1], d.value[2] = L[2];
            } else if (e->class_()->name() == "SpotLight") {               // spot.cpp:89-114
                d.kind = DTOF_EMITTER_SPOT;
                const ScalarTransform4f tw = e->world_transform();   // scalar_rgb: Transform4f is the scalar transform
                const ScalarTransform4f inv = tw.inverse();
                for (int r = 0; r < 3; ++r) {
                    d.position[r] = tw.matrix(r, 3);
                    for (int cc = 0; cc < 3; ++cc)
                        d.to_local[3 * r + cc] = inv.matrix(r, cc);
                }
                for (auto &o : c.objects)
                    if (o.first == "intensity") {
                        Spectrum I = uniform((const Texture<Float, Spectrum> *) o.second, o.first.c_str())->eval(si);
                        d.value[0] = I[0], d.value[1] = I[1], d.value[2] = I[2];
                    } else if (o.first == "texture" && ((const Texture<Float, Spectrum> *) o.second)->is_spatially_varying()) {
                        Throw("spot emitter with a projection texture is outside the accelerated path");
                    }
                // the two angles are not traversed; the plugin prints them in radians (to_string, spot.cpp:265-276)
                const std::string descr = e->to_string();
                auto angle = [&](const char *key) -> float {
                    size_t at = descr.find(key);
                    if (at == std::string::npos)
                        Throw("spot emitter: cannot read %s", key);
                    return std::strtof(descr.c_str() + at + strlen(key), nullptr);
                };
                d.cutoff_angle = angle("cutoff_angle = ");
                d.beam_width = angle("beam_width = ");
            } else if (e->class_()->name() == "DirectionalEmitter") {      // directional.cpp:65-91, 149-176
                d.kind = DTOF_EMITTER_DIRECTIONAL;                       // the library derives the bounding sphere itself
                const ScalarTransform4f tw = e->world_transform();
                for (int r = 0; r < 3; ++r)
                    d.position[r] = tw.matrix(r, 2);                     // to_world * (0, 0, 1)
                for (auto &o : c.objects)
                    if (o.first == "irradiance") {
                        Spectrum I = uniform((const Texture<Float, Spectrum> *) o.second, o.first.c_str())->eval(si);
                        d.value[0] = I[0], d.value[1] = I[1], d.value[2] = I[2];
                    }
            } else if (e->class_()->name() == "ConstantBackgroundEmitter") {
                d.kind = DTOF_EMITTER_CONSTANT;                          // the library derives the bounding sphere itself
                Spectrum L = e->eval(si);
                d.value[0] = L[0], d.value[1] = L[1], d.value[2] = L[2];
            } else {
                Throw("emitter \"%s\" is outside the accelerated path (point | spot | directional | area | constant)", e->class_()->name());
            }
            emitters.push_back(d);
        }
        for (size_t i = 0; i < meshes.size(); ++i) {
            meshes[i].n_vertices = (uint32_t) bufs[i].pos.size() / 3;
            meshes[i].n_faces = (uint32_t) bufs[i].idx.size() / 3;
            meshes[i].positions = bufs[i].pos.data();
            meshes[i].normals = bufs[i].nrm.empty() ? nullptr : bufs[i].nrm.data();
            meshes[i].texcoords = bufs[i].uv.empty() ? nullptr : bufs[i].uv.data();
            meshes[i].faces = bufs[i].idx.data();
        }
        dtof_scene_desc d{};
        d.n_meshes = (uint32_t) meshes.size(), d.meshes = meshes.data();
        d.n_instances = (uint32_t) instances.size(), d.instances = instances.data();
        d.n_bsdfs = (uint32_t) bsdfs.size(), d.bsdfs = bsdfs.data();
        d.n_emitters = (uint32_t) emitters.size(), d.emitters = emitters.data();
        // camera: PerspectiveCamera (perspective.cpp:172-198); sample_to_camera is exposed through traverse()
        Collector c;
        const_cast<Sensor *>(sensor)->traverse(&c);
        m34(*c.param<ScalarTransform4f>("to_world"), d.camera.to_world);
        const Film *film = sensor->film();
        ScalarFloat x_fov = *c.param<ScalarFloat>("x_fov"), nearc = *c.param<ScalarFloat>("near_clip"), farc = *c.param<ScalarFloat>("far_clip");
        ScalarTransform4f s2c = perspective_projection(film->size(), film->crop_size(), film->crop_offset(), x_fov, nearc, farc).inverse();
        for (int r = 0; r < 4; ++r)
            for (int k = 0; k < 4; ++k)
                d.camera.sample_to_camera[4 * r + k] = s2c.matrix(r, k);
        d.camera.near_clip = nearc, d.camera.far_clip = farc;
        d.camera.shutter_open = sensor->shutter_open(), d.camera.shutter_open_time = sensor->shutter_open_time();
        d.film.width = film->crop_size().x(), d.film.height = film->crop_size().y();
        d.film.crop_offset_x = film->crop_offset().x(), d.film.crop_offset_y = film->crop_offset().y();
        if (film->sample_border())   // integrator.cpp:176-178 samples a larger region then; not built, do not ignore it
            Throw("a film with sample_border=true is outside the accelerated path");
        const std::string rf = film->rfilter()->class_()->name();
        static const std::pair<const char *, uint32_t> filters[] = {
            { "BoxFilter", DTOF_RFILTER_BOX }, { "TentFilter", DTOF_RFILTER_TENT }, { "GaussianFilter", DTOF_RFILTER_GAUSSIAN },
            { "MitchellNetravaliFilter", DTOF_RFILTER_MITCHELL }, { "CatmullRomFilter", DTOF_RFILTER_CATMULLROM },
            { "LanczosSincFilter", DTOF_RFILTER_LANCZOS } };
        bool known = false;
        for (auto &f : filters)
            if (rf == f.first)
                d.film.rfilter = f.second, known = true;
        if (!known)
            Throw("reconstruction filter \"%s\" is not one the accelerated path knows", rf);
        d.film.rfilter_radius = film->rfilter()->radius();
        d.film.gaussian_stddev = d.film.rfilter_radius / 4.f;            // gaussian.cpp:50-53: radius = 4 * stddev
        d.film.mitchell_b = d.film.mitchell_c = 1.f / 3.f;
        if (d.film.rfilter == DTOF_RFILTER_MITCHELL) {
            // B and C are not traversed; to_string prints them with six decimals (mitchell.cpp:85-87). Thirds, the usual
            // choice, are not representable in six decimals and are snapped back.
            const std::string descr = film->rfilter()->to_string();
            auto value = [&](const char *key) -> float {
                size_t at = descr.find(key);
                if (at == std::string::npos)
                    Throw("mitchell filter: cannot read %s", key);
                double v = std::strtod(descr.c_str() + at + strlen(key), nullptr);
                double thirds = std::round(v * 3.0) / 3.0;
                return (float) (std::abs(v - thirds) < 1e-6 ? thirds : v);
            };
            d.film.mitchell_b = value("B=");
            d.film.mitchell_c = value("C=");
        }
        if (dtof_upload_scene(m_ctx, &d) != DTOF_OK)
            Throw("dtof_upload_scene: %s", dtof_last_error(m_ctx));
    }

    dtof_params m_p;
    dtof_ctx *m_ctx = nullptr;
    const Scene *m_uploaded = nullptr;
    const Sensor *m_uploaded_sensor = nullptr;
    ScalarVector2u m_uploaded_size = ScalarVector2u(0, 0);
    ScalarPoint2u m_uploaded_offset = ScalarPoint2u(0, 0);
    int m_device = 0;
};

MI_IMPLEMENT_CLASS_VARIANT(DopplerToFPathB200, MonteCarloIntegrator)
MI_EXPORT_PLUGIN(DopplerToFPathB200, "Doppler ToF path tracer (B200 CUDA library behind the dtof C ABI)")
NAMESPACE_END(mitsuba)
