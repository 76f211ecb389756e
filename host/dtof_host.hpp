// C++ host side of the B200 Doppler-ToF path: everything above the C ABI of libdtof_b200.so (include/dtof.h).
//
// Mirrors, for the hot path only, what the reference's host does before `Integrator::render`:
//   xml::load_file            src/core/xml.cpp (defaults / -D parameters :300-330, transforms :820-1007,
//                             <animation> keyframes :520-524,882-900,996-1007, animated shape -> shapegroup +
//                             instance rewrite :1166-1192, "unreferenced property" errors :1204-1223)
//   Properties / plugins      src/integrators/dopplertofpath.cpp:19-57, src/render/integrator.cpp:24-27,54-100,568-585,
//                             src/render/sampler.cpp:13-14, src/samplers/correlated.cpp:17-23
//   Transform / Animated      include/mitsuba/core/transform.h:24-551, src/core/transform.cpp:22-36
//   PerspectiveCamera         src/sensors/perspective.cpp:172-198, include/mitsuba/render/sensor.h:227-262,
//                             src/render/sensor.cpp:149-203
//   shapes                    src/shapes/rectangle.cpp, src/shapes/cube.cpp:109-165, src/shapes/ply.cpp, obj.cpp
// The Python package mitsuba3dopplertof_b200/ is the same interface for tests and torch.distributed; the two
// flatten a scene to identical dtof_scene_desc contents (tests/test_cpp_host.py).
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/dtof.h"

namespace dtof_host {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------------------------------------- XML DOM
struct XmlNode {
    std::string tag;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlNode>> children;
    const std::string *attr(const std::string &name) const;
};
std::unique_ptr<XmlNode> parse_xml(const std::string &text);

// ---------------------------------------------------------------------------------------------- transforms
// 4x4 transform with its tracked inverse transpose. XML transforms are composed in double
// (Properties::Float = double) and narrowed to float32 when a plugin reads them.
struct Transform4 {
    double m[16];    // row-major
    double it[16];   // inverse transpose
    static Transform4 identity();
    static Transform4 from_matrix(const double *m16);
    static Transform4 translate(double x, double y, double z);
    static Transform4 scale(double x, double y, double z);
    static Transform4 rotate(double ax, double ay, double az, double angle_deg);
    static Transform4 look_at(const double *origin, const double *target, const double *up);
    Transform4 operator*(const Transform4 &o) const;   // this @ o
    Transform4 narrowed() const;                       // every entry rounded to float32
    bool has_scale() const;
    void m34(float *out12) const;
    // float32 application with Dr.Jit's fma chains (transform.h:96-150)
    void affine_point(const float *p, float *out) const;
    void normal(const float *n, float *out) const;
};

struct AnimatedTransform {
    std::vector<float> times;
    std::vector<Transform4> transforms;   // narrowed to float32 on append
    void append(double time, const Transform4 &t);
    size_t size() const { return times.size(); }
    float min_time() const;
    float max_time() const;
    void eval(float time, float *out16) const;   // linear matrix interpolation between keyframes 0 and 1
};

// ---------------------------------------------------------------------------------------------- scene objects
// `diffuse`, or `conductor` (reflectance = specular_reflectance, eta + i k = complex index of refraction)
struct Bsdf {
    float reflectance[3] = { 0.5f, 0.5f, 0.5f };
    bool twosided = false;
    uint32_t kind = DTOF_BSDF_DIFFUSE;
    float eta[3] = { 0.f, 0.f, 0.f }, k[3] = { 1.f, 1.f, 1.f };
    float alpha[2] = { 0.f, 0.f };   // roughconductor: alpha_u, alpha_v
    uint32_t distribution = 0;       // roughconductor: 0 beckmann, 1 ggx
};

struct Shape {
    enum Kind { Rectangle, Cube, Mesh } kind = Mesh;
    std::string id;
    bool has_static = false, has_anim = false;
    Transform4 to_world = Transform4::identity();
    AnimatedTransform anim;
    bool has_bsdf = false;
    Bsdf bsdf;
    bool emitter = false;
    float radiance[3] = { 1, 1, 1 };
    bool flip_normals = false;
    std::vector<float> positions, normals, texcoords;   // Mesh payload (object space)
    std::vector<uint32_t> faces;
    bool smooth_normals = false;   // the file had no normals: computed at flatten time, after to_world (mesh.cpp:283-345)
    bool animated() const { return has_anim && anim.size() > 1; }
};

// `point` (src/emitters/point.cpp), or with constant_env the constant environment emitter (src/emitters/constant.cpp),
// whose radiance is kept in `intensity`
struct PointLight {
    float position[3] = { 0, 0, 0 };
    float intensity[3] = { 1, 1, 1 };
    bool constant_env = false;
    // `spot` (src/emitters/spot.cpp, no projection texture): linear part of to_world^-1 and the two angles in radians
    bool spot = false;
    float to_local[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    float cutoff_angle = 0.f, beam_width = 0.f;
    // `directional` (src/emitters/directional.cpp): `position` holds to_world * (0, 0, 1), `intensity` the irradiance
    bool directional = false;
};

struct Film {
    uint32_t width = 768, height = 576;
    bool has_crop = false;
    uint32_t crop_w = 0, crop_h = 0, crop_x = 0, crop_y = 0;
    std::string rfilter = "gaussian";   // src/render/film.cpp:49-54
    bool has_radius = false;
    double radius = 1.0, stddev = 0.5;
    double mitchell_b = 1.0 / 3.0, mitchell_c = 1.0 / 3.0;   // src/rfilters/mitchell.cpp:52-60
    int lanczos_lobes = 3;                                   // src/rfilters/lanczos.cpp:38-42
    dtof_film abi() const;
};

struct CorrelatedSampler {
    uint32_t sample_count = 4, seed = 0, time_correlate_number = 2, path_correlate_number = 2;
    bool correlated = true;   // false: `independent` (PCG32Sampler only), usable by the non-Doppler integrators
};

struct PerspectiveSensor {
    Transform4 to_world = Transform4::identity();
    double fov = 45.0;
    std::string fov_axis = "x";
    double near_clip = 1e-2, far_clip = 1e4, shutter_open = 0.0, shutter_close = 0.0;
    Film film;
    CorrelatedSampler sampler;
    dtof_camera abi() const;
};

// `dopplertofpath` property surface; defaults and derived values exactly as the reference constructors.
struct DopplerToFPathIntegrator {
    float time = 0.0015f, w_g = 30.f, g_1 = 0.5f, g_0 = 0.5f, w_s = 30.f, sensor_phase_offset = 0.f, hetero_frequency = 0.f;
    uint32_t wave_function_type = DTOF_WAVE_SINUSOIDAL;
    bool low_frequency_component_only = true;
    uint32_t time_sampling_method = DTOF_TIME_ANTITHETIC;
    float antithetic_shift = 0.5f;
    bool use_stratified_sampling_for_each_interval = true;
    uint32_t path_correlation_depth = 0;
    int32_t max_depth = -1, rr_depth = 5;
    bool hide_emitters = false;
    double timeout = -1.0;
    uint32_t kind = DTOF_INTEGRATOR_DOPPLERTOFPATH;   // or _VELOCITY (src/integrators/velocity.cpp), _PATH (src/integrators/path.cpp)
    // props: name -> textual value (already $-substituted); throws on unknown names / bad enum strings
    explicit DopplerToFPathIntegrator(const std::map<std::string, std::string> &props = {},
                                      uint32_t kind = DTOF_INTEGRATOR_DOPPLERTOFPATH);
    dtof_params params(const CorrelatedSampler &s, uint32_t seed = 0, uint32_t spp = 0) const;
};

// Owns every buffer the dtof_scene_desc points into.
struct FlatScene {
    struct MeshBuf {
        std::vector<float> positions, normals, texcoords;
        std::vector<uint32_t> faces;
    };
    std::vector<MeshBuf> bufs;
    std::vector<dtof_mesh> meshes;
    std::vector<dtof_instance> instances;
    std::vector<dtof_bsdf> bsdfs;
    std::vector<dtof_emitter> emitters;
    dtof_scene_desc desc{};
    uint64_t n_triangles = 0;
    void finalize();                                    // wire the pointers of `desc`
    void serialize(const std::string &path) const;      // canonical dump (host parity test)
};

struct Scene {
    std::vector<Shape> shapes;
    std::vector<PointLight> emitters;
    std::vector<std::pair<char, uint32_t>> order;   // ('s', shape index) | ('e', emitter index) in file order
    PerspectiveSensor sensor;
    DopplerToFPathIntegrator integrator;
    std::unique_ptr<FlatScene> flatten() const;
};

// `mi.load_file(path, **params)` / `mitsuba -Dkey=value scene.xml`
Scene load_file(const std::string &path, const std::map<std::string, std::string> &params = {});
Scene load_string(const std::string &xml, const std::string &base_dir, const std::map<std::string, std::string> &params = {});

// mesh files
// `compute_missing = false` leaves `normals` empty when the file has none (the caller computes them after to_world)
void load_mesh_file(const std::string &path, bool face_normals, std::vector<float> &pos, std::vector<uint32_t> &faces,
                    std::vector<float> &normals, std::vector<float> &uvs, bool compute_missing = true,
                    bool flip_tex_coords = true);   // obj only (obj.cpp:151)
void load_serialized_file(const std::string &path, int shape_index, bool face_normals, std::vector<float> &pos,
                          std::vector<uint32_t> &faces, std::vector<float> &normals, std::vector<float> &uvs,
                          bool compute_missing = true);
void vertex_normals(const std::vector<float> &pos, const std::vector<uint32_t> &faces, std::vector<float> &out);

// ---------------------------------------------------------------------------------------------- renderer
// RAII wrapper of one dtof_ctx; errors become exceptions carrying dtof_last_error(). No CPU fallback.
class Renderer {
public:
    explicit Renderer(int device = 0);
    ~Renderer();
    Renderer(const Renderer &) = delete;
    void upload(const FlatScene &flat);
    // Integrator::render(scene, sensor, seed, spp, develop): returns H*W*3 (develop) or H*W*4 RGBW floats
    std::vector<float> render(const dtof_params &p, bool develop, uint32_t width, uint32_t height);
    float last_kernel_ms();
    dtof_ctx *ctx = nullptr;

private:
    void check(dtof_status s, const char *what);
};

void write_pfm(const std::string &path, const float *rgb, uint32_t w, uint32_t h, uint32_t channels);
void write_npy(const std::string &path, const float *data, uint32_t h, uint32_t w, uint32_t c);

} // namespace dtof_host
