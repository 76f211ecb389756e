// Host-side scene model: transforms, XML loader, plugin property surfaces, flattening into dtof_scene_desc.
// See dtof_host.hpp for the reference files each part mirrors. Compile with -ffp-contract=off: fused
// multiply-adds are explicit (std::fmaf) exactly where Dr.Jit's helpers fuse.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>

#include "dtof_host.hpp"

namespace dtof_host {

namespace {

constexpr double kPi = 3.14159265358979323846;

void mat_identity(double *m) {
    for (int i = 0; i < 16; ++i)
        m[i] = (i % 5 == 0) ? 1.0 : 0.0;
}
void mat_mul(const double *a, const double *b, double *o) {
    double r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k)
                s += a[4 * i + k] * b[4 * k + j];
            r[4 * i + j] = s;
        }
    memcpy(o, r, sizeof(r));
}
// general 4x4 inverse (Gauss-Jordan with partial pivoting), double precision
bool mat_inverse(const double *m, double *out) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            a[i][j] = m[4 * i + j];
            a[i][4 + j] = i == j ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r)
            if (std::fabs(a[r][c]) > std::fabs(a[p][c]))
                p = r;
        if (a[p][c] == 0.0)
            return false;
        if (p != c)
            for (int j = 0; j < 8; ++j)
                std::swap(a[p][j], a[c][j]);
        double inv = 1.0 / a[c][c];
        for (int j = 0; j < 8; ++j)
            a[c][j] *= inv;
        for (int r = 0; r < 4; ++r) {
            if (r == c)
                continue;
            double f = a[r][c];
            if (f != 0.0)
                for (int j = 0; j < 8; ++j)
                    a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            out[4 * i + j] = a[i][4 + j];
    return true;
}
void transpose(const double *m, double *o) {
    double r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r[4 * j + i] = m[4 * i + j];
    memcpy(o, r, sizeof(r));
}
// Dr.Jit Matrix4f product in float32 (ext/drjit/include/drjit/matrix.h): sum = a.col(0) * b(0,j); fmadd chain
void mat_mul_f32(const double *a, const double *b, double *o) {
    double r[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            float s = (float) a[4 * i] * (float) b[j];
            for (int k = 1; k < 4; ++k)
                s = std::fmaf((float) a[4 * i + k], (float) b[4 * k + j], s);
            r[4 * i + j] = s;
        }
    memcpy(o, r, sizeof(r));
}

std::vector<double> parse_floats(const std::string &s) {
    std::vector<double> v;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && (std::isspace((unsigned char) s[i]) || s[i] == ','))
            ++i;
        if (i >= s.size())
            break;
        size_t b = i;
        while (i < s.size() && !std::isspace((unsigned char) s[i]) && s[i] != ',')
            ++i;
        std::string tok = s.substr(b, i - b);
        char *end = nullptr;
        double d = std::strtod(tok.c_str(), &end);
        if (end == tok.c_str() || *end)
            throw Error("could not parse floating point value \"" + tok + "\"");
        v.push_back(d);
    }
    return v;
}
double parse_float(const std::string &s) {
    auto v = parse_floats(s);
    if (v.size() != 1)
        throw Error("could not parse floating point value \"" + s + "\"");
    return v[0];
}
long long parse_int(const std::string &s) {
    char *end = nullptr;
    std::string t = s;
    while (!t.empty() && std::isspace((unsigned char) t.back()))
        t.pop_back();
    long long v = std::strtoll(t.c_str(), &end, 10);
    if (end == t.c_str() || *end)
        throw Error("could not parse integer value \"" + s + "\"");
    return v;
}
bool parse_bool(const std::string &s) {
    std::string t;
    for (char c : s)
        if (!std::isspace((unsigned char) c))
            t += (char) std::tolower((unsigned char) c);
    if (t == "true")
        return true;
    if (t == "false")
        return false;
    throw Error("could not parse boolean value \"" + s + "\" -- must be \"true\" or \"false\"");
}

} // namespace

// ================================================================================================ Transform4
Transform4 Transform4::identity() {
    Transform4 t;
    mat_identity(t.m);
    mat_identity(t.it);
    return t;
}
Transform4 Transform4::from_matrix(const double *m16) {
    Transform4 t;
    memcpy(t.m, m16, sizeof(t.m));
    double inv[16];
    if (!mat_inverse(m16, inv))
        throw Error("singular matrix in <matrix>");
    transpose(inv, t.it);
    return t;
}
Transform4 Transform4::translate(double x, double y, double z) {
    Transform4 t = identity();
    t.m[3] = x, t.m[7] = y, t.m[11] = z;
    t.it[12] = -x, t.it[13] = -y, t.it[14] = -z;
    return t;
}
Transform4 Transform4::scale(double x, double y, double z) {
    Transform4 t = identity();
    t.m[0] = x, t.m[5] = y, t.m[10] = z;
    t.it[0] = 1.0 / x, t.it[5] = 1.0 / y, t.it[10] = 1.0 / z;
    return t;
}
Transform4 Transform4::rotate(double ax, double ay, double az, double angle_deg) {
    double len = std::sqrt(ax * ax + ay * ay + az * az);
    double x = ax / len, y = ay / len, z = az / len;
    double ang = angle_deg * (kPi / 180.0), s = std::sin(ang), c = std::cos(ang);
    Transform4 t = identity();
    double r[9] = { c + x * x * (1 - c),     x * y * (1 - c) - z * s, x * z * (1 - c) + y * s,
                    y * x * (1 - c) + z * s, c + y * y * (1 - c),     y * z * (1 - c) - x * s,
                    z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c) };
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            t.m[4 * i + j] = t.it[4 * i + j] = r[3 * i + j];
    return t;
}
Transform4 Transform4::look_at(const double *o, const double *tg, const double *up) {
    double d[3] = { tg[0] - o[0], tg[1] - o[1], tg[2] - o[2] };
    double dl = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (double &v : d)
        v /= dl;
    double l[3] = { up[1] * d[2] - up[2] * d[1], up[2] * d[0] - up[0] * d[2], up[0] * d[1] - up[1] * d[0] };
    double ll = std::sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    for (double &v : l)
        v /= ll;
    double nu[3] = { d[1] * l[2] - d[2] * l[1], d[2] * l[0] - d[0] * l[2], d[0] * l[1] - d[1] * l[0] };
    double m[16];
    mat_identity(m);
    for (int i = 0; i < 3; ++i) {
        m[4 * i + 0] = l[i];
        m[4 * i + 1] = nu[i];
        m[4 * i + 2] = d[i];
        m[4 * i + 3] = o[i];
    }
    return from_matrix(m);
}
Transform4 Transform4::operator*(const Transform4 &o) const {
    Transform4 r;
    mat_mul(m, o.m, r.m);
    mat_mul(it, o.it, r.it);
    return r;
}
Transform4 Transform4::narrowed() const {
    Transform4 r;
    for (int i = 0; i < 16; ++i) {
        r.m[i] = (float) m[i];
        r.it[i] = (float) it[i];
    }
    return r;
}
bool Transform4::has_scale() const {
    for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int i = 0; i < 3; ++i)
            s += m[4 * i + j] * m[4 * i + j];
        if (std::fabs(s - 1.0) > 1e-3)
            return true;
    }
    return false;
}
void Transform4::m34(float *out) const {
    for (int i = 0; i < 12; ++i)
        out[i] = (float) m[i];
}
void Transform4::affine_point(const float *p, float *out) const {
    for (int r = 0; r < 3; ++r) {
        float a = (float) m[4 * r + 3];
        for (int i = 0; i < 3; ++i)
            a = std::fmaf((float) m[4 * r + i], p[i], a);
        out[r] = a;
    }
}
void Transform4::normal(const float *n, float *out) const {
    for (int r = 0; r < 3; ++r) {
        float a = (float) it[4 * r] * n[0];
        for (int i = 1; i < 3; ++i)
            a = std::fmaf((float) it[4 * r + i], n[i], a);
        out[r] = a;
    }
}

// ================================================================================================ AnimatedTransform
void AnimatedTransform::append(double time, const Transform4 &t) {
    float tf = (float) time;
    if (!times.empty() && (double) tf <= (double) times.back())
        throw Error("AnimatedTransform::append(): time values must be strictly monotonically increasing!");
    times.push_back(tf);
    transforms.push_back(t.narrowed());
}
float AnimatedTransform::min_time() const { return times.empty() ? 100.f : *std::min_element(times.begin(), times.end()); }
float AnimatedTransform::max_time() const { return times.empty() ? -100.f : *std::max_element(times.begin(), times.end()); }
void AnimatedTransform::eval(float time, float *out) const {
    if (size() <= 1) {
        for (int i = 0; i < 16; ++i)
            out[i] = transforms.empty() ? ((i % 5 == 0) ? 1.f : 0.f) : (float) transforms[0].m[i];
        return;
    }
    float t0 = times[0], t1 = times[1];
    float t = std::min(std::max((time - t0) / (t1 - t0), 0.f), 1.f);
    float s = 1.f - t;
    for (int i = 0; i < 16; ++i) {
        float a = (float) transforms[0].m[i] * s, b = (float) transforms[1].m[i] * t;
        out[i] = a + b;
    }
}

// ================================================================================================ sensor / film
dtof_film Film::abi() const {
    dtof_film f{};
    f.width = has_crop ? crop_w : width;
    f.height = has_crop ? crop_h : height;
    f.crop_offset_x = crop_x;
    f.crop_offset_y = crop_y;
    if (rfilter == "box") {
        f.rfilter = DTOF_RFILTER_BOX;
        f.rfilter_radius = 0.5f;
    } else if (rfilter == "tent") {
        f.rfilter = DTOF_RFILTER_TENT;
        f.rfilter_radius = has_radius ? (float) radius : 1.f;
    } else if (rfilter == "gaussian") {
        f.rfilter = DTOF_RFILTER_GAUSSIAN;
        f.rfilter_radius = (float) (4.0 * stddev);   // src/rfilters/gaussian.cpp:50-53
    } else if (rfilter == "mitchell" || rfilter == "catmullrom") {
        f.rfilter = rfilter == "mitchell" ? DTOF_RFILTER_MITCHELL : DTOF_RFILTER_CATMULLROM;
        f.rfilter_radius = 2.f;
    } else if (rfilter == "lanczos") {
        f.rfilter = DTOF_RFILTER_LANCZOS;
        f.rfilter_radius = (float) lanczos_lobes;   // src/rfilters/lanczos.cpp:38-42
    } else {
        throw Error("rfilter '" + rfilter + "' is not a reconstruction filter (box|tent|gaussian|mitchell|catmullrom|lanczos)");
    }
    f.gaussian_stddev = (float) stddev;
    f.mitchell_b = (float) mitchell_b, f.mitchell_c = (float) mitchell_c;
    return f;
}

namespace {
// src/render/sensor.cpp:149-203 ('fov' branch, double precision)
double parse_fov(double fov, std::string axis, double aspect) {
    for (char &c : axis)
        c = (char) std::tolower((unsigned char) c);
    if (axis == "smaller")
        axis = aspect > 1 ? "y" : "x";
    else if (axis == "larger")
        axis = aspect > 1 ? "x" : "y";
    double result;
    auto rad = [](double d) { return d * (kPi / 180.0); };
    auto deg = [](double r) { return r * (180.0 / kPi); };
    if (axis == "x")
        result = fov;
    else if (axis == "y")
        result = deg(2.0 * std::atan(std::tan(0.5 * rad(fov)) * aspect));
    else if (axis == "diagonal") {
        double diagonal = 2.0 * std::tan(0.5 * rad(fov));
        double width = diagonal / std::sqrt(1.0 + 1.0 / (aspect * aspect));
        result = deg(2.0 * std::atan(width * 0.5));
    } else
        throw Error("The 'fov_axis' parameter must be set to one of 'smaller', 'larger', 'diagonal', 'x', or 'y'!");
    if (result <= 0.0 || result >= 180.0)
        throw Error("The horizontal field of view must be in the range [0, 180]!");
    return result;
}

struct TF32 {   // float32 transform pair used by the sensor-internal composition
    double m[16], it[16];
};
TF32 tf_mul(const TF32 &a, const TF32 &b) {
    TF32 r;
    mat_mul_f32(a.m, b.m, r.m);
    mat_mul_f32(a.it, b.it, r.it);
    return r;
}
TF32 tf_scale(float x, float y, float z) {
    TF32 t;
    mat_identity(t.m);
    mat_identity(t.it);
    t.m[0] = x, t.m[5] = y, t.m[10] = z;
    t.it[0] = 1.f / x, t.it[5] = 1.f / y, t.it[10] = 1.f / z;
    return t;
}
TF32 tf_translate(float x, float y, float z) {
    TF32 t;
    mat_identity(t.m);
    mat_identity(t.it);
    t.m[3] = x, t.m[7] = y, t.m[11] = z;
    t.it[12] = -x, t.it[13] = -y, t.it[14] = -z;
    return t;
}
} // namespace

dtof_camera PerspectiveSensor::abi() const {
    if (shutter_close < shutter_open)   // src/render/sensor.cpp:18-20
        throw Error("Shutter opening time must be less than or equal to the shutter closing time!");
    if (near_clip <= 0 || near_clip >= far_clip)
        throw Error("The 'near_clip' parameter must be greater than zero and smaller than 'far_clip'.");
    Transform4 tw = to_world.narrowed();
    if (tw.has_scale())
        throw Error("Scale factors in the camera-to-world transformation are not allowed!");
    const float fw = (float) film.width, fh = (float) film.height;
    const uint32_t cw = film.has_crop ? film.crop_w : film.width, ch = film.has_crop ? film.crop_h : film.height;
    const float x_fov = (float) parse_fov(fov, fov_axis, (double) film.width / (double) film.height);
    const float nearf = (float) near_clip, farf = (float) far_clip;
    // perspective_projection<float> (sensor.h:227-262) + Transform::perspective (transform.h:216-233), float32
    const float rel_sx = (float) cw / fw, rel_sy = (float) ch / fh;
    const float rel_ox = (float) film.crop_x / fw, rel_oy = (float) film.crop_y / fh;
    const float aspect = fw / fh;
    const float recip = 1.f / (farf - nearf);
    const float half = x_fov * 0.5f;
    const float radf = half * (float) (kPi / 180.0);
    const float tanv = (float) std::tan((double) radf);
    const float cot = 1.f / tanv;
    TF32 persp;
    for (int i = 0; i < 16; ++i)
        persp.m[i] = persp.it[i] = 0.0;
    persp.m[0] = cot, persp.m[5] = cot, persp.m[10] = farf * recip, persp.m[11] = -nearf * farf * recip, persp.m[14] = 1.f;
    double inv[16] = { 0 };
    inv[0] = tanv, inv[5] = tanv, inv[15] = 1.f / nearf, inv[11] = 1.f, inv[14] = (nearf - farf) / (farf * nearf);
    transpose(inv, persp.it);
    TF32 c2s = tf_mul(tf_mul(tf_mul(tf_mul(tf_scale(1.f / rel_sx, 1.f / rel_sy, 1.f), tf_translate(-rel_ox, -rel_oy, 0.f)),
                                    tf_scale(-0.5f, -0.5f * aspect, 1.f)),
                             tf_translate(-1.f, -1.f / aspect, 0.f)),
                      persp);
    dtof_camera cam{};
    tw.m34(cam.to_world);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            cam.sample_to_camera[4 * i + j] = (float) c2s.it[4 * j + i];   // inverse().matrix = inverse_transpose^T
    cam.near_clip = nearf;
    cam.far_clip = farf;
    float so = (float) shutter_open, sc = (float) shutter_close;
    cam.shutter_open = so;
    cam.shutter_open_time = sc - so;
    return cam;
}

// ================================================================================================ integrator
DopplerToFPathIntegrator::DopplerToFPathIntegrator(const std::map<std::string, std::string> &props, uint32_t kind_) : kind(kind_) {
    static const std::set<std::string> velocity_known = {
        "time", "max_depth", "rr_depth", "hide_emitters", "timeout", "block_size", "samples_per_pass", "time_sampling_method",
        "antithetic_shift", "use_stratified_sampling_for_each_interval", "path_correlation_depth", "is_doppler_integrator" };
    if (kind == DTOF_INTEGRATOR_VELOCITY)
        for (auto &kv : props)
            if (!velocity_known.count(kv.first))
                throw Error("velocity: unreferenced property \"" + kv.first + "\"");
    if (kind == DTOF_INTEGRATOR_PATH)   // PathIntegrator (src/integrators/path.cpp:93) reads nothing of its own
        for (auto &kv : props)
            if (kv.first == "time" || !velocity_known.count(kv.first))
                throw Error("path: unreferenced property \"" + kv.first + "\"");
    static const std::set<std::string> known = {
        "time", "w_g", "g_1", "g_0", "w_s", "sensor_phase_offset", "hetero_offset", "hetero_frequency", "wave_function_type",
        "low_frequency_component_only", "max_depth", "rr_depth", "hide_emitters", "timeout", "is_doppler_integrator",
        "time_sampling_method", "antithetic_shift", "use_stratified_sampling_for_each_interval", "path_correlation_depth",
        "block_size", "samples_per_pass" };
    for (auto &kv : props)
        if (!known.count(kv.first))
            throw Error("dopplertofpath: unreferenced property \"" + kv.first + "\"");   // xml.cpp:1204-1223
    auto has = [&](const char *k) { return props.count(k) != 0; };
    auto getf = [&](const char *k, float d) { return has(k) ? (float) parse_float(props.at(k)) : d; };
    auto getb = [&](const char *k, bool d) { return has(k) ? parse_bool(props.at(k)) : d; };
    // DopplerToFPathIntegrator ctor, dopplertofpath.cpp:19-57 (ScalarFloat = float)
    time = getf("time", 0.0015f);
    w_g = getf("w_g", 30.f);
    g_1 = getf("g_1", 0.5f);
    g_0 = getf("g_0", 0.5f);
    w_s = getf("w_s", 30.f);
    sensor_phase_offset = getf("sensor_phase_offset", 0.f);
    if (has("hetero_offset"))   // float * int * double -> double -> float
        sensor_phase_offset = (float) ((double) getf("hetero_offset", 0.f) * 2 * kPi);
    if (has("hetero_frequency")) {
        hetero_frequency = getf("hetero_frequency", 1.f);
        w_s = (float) ((double) w_g + (double) (hetero_frequency / time) * 1e-6);
    } else {
        hetero_frequency = (float) ((double) (w_s - w_g) * 1e6 * (double) time);
    }
    std::string wave = has("wave_function_type") ? props.at("wave_function_type") : "sinusoidal";
    if (wave == "sinusoidal") wave_function_type = DTOF_WAVE_SINUSOIDAL;
    else if (wave == "rectangular") wave_function_type = DTOF_WAVE_RECTANGULAR;
    else if (wave == "triangular") wave_function_type = DTOF_WAVE_TRIANGULAR;
    else if (wave == "trapezoidal") wave_function_type = DTOF_WAVE_TRAPEZOIDAL;
    else throw Error("dopplertofpath: unknown wave_function_type '" + wave + "'");   // reference: enum left uninitialised
    low_frequency_component_only = getb("low_frequency_component_only", true);
    // SamplingIntegrator ctor, integrator.cpp:54-100
    std::string method = has("time_sampling_method") ? props.at("time_sampling_method") : "antithetic";
    if (method == "uniform") time_sampling_method = DTOF_TIME_UNIFORM;
    else if (method == "stratified") time_sampling_method = DTOF_TIME_STRATIFIED;
    else if (method == "antithetic") time_sampling_method = DTOF_TIME_ANTITHETIC;
    else if (method == "antithetic_mirror") time_sampling_method = DTOF_TIME_ANTITHETIC_MIRROR;
    else throw Error("dopplertofpath: unknown time_sampling_method '" + method + "'");
    antithetic_shift = getf("antithetic_shift", time_sampling_method == DTOF_TIME_ANTITHETIC ? 0.5f : 0.f);
    use_stratified_sampling_for_each_interval = getb("use_stratified_sampling_for_each_interval", true);
    path_correlation_depth = has("path_correlation_depth") ? (uint32_t) parse_int(props.at("path_correlation_depth")) : 0u;
    if (has("samples_per_pass"))
        throw Error("'samples_per_pass' is deprecated in the reference and unsupported here");
    // MonteCarloIntegrator ctor, integrator.cpp:568-585
    long long md = has("max_depth") ? parse_int(props.at("max_depth")) : -1;
    if (md < 0 && md != -1)
        throw Error("\"max_depth\" must be set to -1 (infinite) or a value >= 0");
    max_depth = (int32_t) md;
    long long rr = has("rr_depth") ? parse_int(props.at("rr_depth")) : 5;
    if (rr <= 0)
        throw Error("\"rr_depth\" must be set to a value greater than zero!");
    rr_depth = (int32_t) rr;
    hide_emitters = getb("hide_emitters", false);
    timeout = has("timeout") ? parse_float(props.at("timeout")) : -1.0;
}

dtof_params DopplerToFPathIntegrator::params(const CorrelatedSampler &s, uint32_t seed, uint32_t spp) const {
    dtof_params p{};
    p.time = time, p.w_g = w_g, p.g_1 = g_1, p.g_0 = g_0;
    p.sensor_phase_offset = sensor_phase_offset;
    p.hetero_frequency = hetero_frequency;
    p.wave_function_type = wave_function_type;
    p.low_frequency_component_only = low_frequency_component_only;
    p.max_depth = max_depth, p.rr_depth = rr_depth, p.hide_emitters = hide_emitters;
    p.time_sampling_method = time_sampling_method;
    p.antithetic_shift = antithetic_shift;
    p.use_stratified_sampling_for_each_interval = use_stratified_sampling_for_each_interval;
    p.path_correlation_depth = path_correlation_depth;
    p.sample_count = spp ? spp : s.sample_count;   // integrator.cpp:121-124
    p.base_seed = s.seed;
    p.time_correlate_number = s.time_correlate_number;
    p.path_correlate_number = s.path_correlate_number;
    p.seed = seed;
    p.integrator = kind;
    if (kind == DTOF_INTEGRATOR_DOPPLERTOFPATH && !s.correlated) {
        // the base-class next_1d_time / next_*_correlate (sampler.h:131-144) draw from the one independent stream: the
        // correlated sampler's `rng` stream under uniform time sampling without path correlation (correlated.cpp:92-97)
        p.time_sampling_method = DTOF_TIME_UNIFORM;
        p.use_stratified_sampling_for_each_interval = 0;
        p.path_correlation_depth = 0;
        p.time_correlate_number = p.path_correlate_number = 1;
        return p;
    }
    if (time_sampling_method == DTOF_TIME_ANTITHETIC_MIRROR && s.time_correlate_number != 2)
        throw Error("antithetic_mirror requires time_correlate_number == 2");   // correlated.cpp:141-142
    if (s.time_correlate_number < 1 || s.path_correlate_number < 1)
        throw Error("correlate numbers must be >= 1");
    return p;
}

// ================================================================================================ flatten
namespace {

const float kCubeV[24][3] = {
    { 1, -1, -1 }, { 1, -1, 1 }, { -1, -1, 1 }, { -1, -1, -1 }, { 1, 1, -1 }, { -1, 1, -1 }, { -1, 1, 1 }, { 1, 1, 1 },
    { 1, -1, -1 }, { 1, 1, -1 }, { 1, 1, 1 }, { 1, -1, 1 }, { 1, -1, 1 }, { 1, 1, 1 }, { -1, 1, 1 }, { -1, -1, 1 },
    { -1, -1, 1 }, { -1, 1, 1 }, { -1, 1, -1 }, { -1, -1, -1 }, { 1, 1, -1 }, { 1, -1, -1 }, { -1, -1, -1 }, { -1, 1, -1 } };
const float kCubeN[6][3] = { { 0, -1, 0 }, { 0, 1, 0 }, { 1, 0, 0 }, { 0, 0, 1 }, { -1, 0, 0 }, { 0, 0, -1 } };
const float kCubeUV[4][2] = { { 0, 1 }, { 1, 1 }, { 1, 0 }, { 0, 0 } };
const uint32_t kCubeF[12][3] = { { 0, 1, 2 }, { 3, 0, 2 }, { 4, 5, 6 }, { 7, 4, 6 }, { 8, 9, 10 }, { 11, 8, 10 },
                                 { 12, 13, 14 }, { 15, 12, 14 }, { 16, 17, 18 }, { 19, 16, 18 }, { 20, 21, 22 }, { 23, 20, 22 } };

// dr::normalize in float32: v * rsqrt(squared_norm(v)), squared_norm as an fma chain
void normalize_f32(float *v) {
    float sq = v[0] * v[0];
    sq = std::fmaf(v[1], v[1], sq);
    sq = std::fmaf(v[2], v[2], sq);
    float inv = 1.f / std::sqrt(sq);
    v[0] *= inv, v[1] *= inv, v[2] *= inv;
}

struct FlatMesh {
    FlatScene::MeshBuf buf;
    uint32_t flip = 0, kind = DTOF_SHAPE_MESH;
    float rect_to_world[12] = { 0 };
    bool has_normals = false, has_uv = false;
};

FlatMesh flatten_shape(const Shape &sh, const Transform4 &trafo) {
    FlatMesh fm;
    Transform4 t32 = trafo.narrowed();
    fm.flip = sh.flip_normals ? 1u : 0u;
    if (sh.kind == Shape::Rectangle) {
        if (sh.flip_normals) {   // rectangle.cpp:91-94: baked into to_world
            t32 = (trafo * Transform4::scale(1.0, 1.0, -1.0)).narrowed();
            fm.flip = 0;
        }
        const float corners[4][3] = { { -1, -1, 0 }, { 1, -1, 0 }, { 1, 1, 0 }, { -1, 1, 0 } };
        fm.buf.positions.resize(12);
        for (int i = 0; i < 4; ++i)
            t32.affine_point(corners[i], &fm.buf.positions[3 * i]);
        fm.buf.texcoords = { 0, 0, 1, 0, 1, 1, 0, 1 };
        fm.has_uv = true;
        const double *m = t32.m;
        double det = m[0] * (m[5] * m[10] - m[6] * m[9]) - m[1] * (m[4] * m[10] - m[6] * m[8]) + m[2] * (m[4] * m[9] - m[5] * m[8]);
        // winding such that normalize(cross(p1-p0, p2-p0)) == normalize(to_world * Normal(0,0,1))
        if (det > 0)
            fm.buf.faces = { 0, 1, 2, 0, 2, 3 };
        else
            fm.buf.faces = { 0, 2, 1, 0, 3, 2 };
        fm.kind = DTOF_SHAPE_RECTANGLE;
        t32.m34(fm.rect_to_world);
        return fm;
    }
    if (sh.kind == Shape::Cube) {
        fm.buf.positions.resize(72);
        fm.buf.normals.resize(72);
        fm.buf.texcoords.resize(48);
        for (int i = 0; i < 24; ++i) {
            t32.affine_point(kCubeV[i], &fm.buf.positions[3 * i]);
            t32.normal(kCubeN[i / 4], &fm.buf.normals[3 * i]);
            normalize_f32(&fm.buf.normals[3 * i]);
            fm.buf.texcoords[2 * i] = kCubeUV[i % 4][0];
            fm.buf.texcoords[2 * i + 1] = kCubeUV[i % 4][1];
        }
        fm.buf.faces.assign(&kCubeF[0][0], &kCubeF[0][0] + 36);
        fm.has_normals = fm.has_uv = true;
        return fm;
    }
    bool ident = true;
    for (int i = 0; i < 16; ++i)
        ident = ident && t32.m[i] == ((i % 5 == 0) ? 1.0 : 0.0);
    size_t nv = sh.positions.size() / 3;
    fm.buf.positions = sh.positions;
    fm.buf.normals = sh.normals;
    fm.buf.texcoords = sh.texcoords;
    fm.buf.faces = sh.faces;
    fm.has_normals = !sh.normals.empty();
    fm.has_uv = !sh.texcoords.empty();
    if (!ident) {
        for (size_t i = 0; i < nv; ++i) {
            t32.affine_point(&sh.positions[3 * i], &fm.buf.positions[3 * i]);
            if (fm.has_normals) {
                t32.normal(&sh.normals[3 * i], &fm.buf.normals[3 * i]);
                normalize_f32(&fm.buf.normals[3 * i]);
            }
        }
    }
    if (sh.smooth_normals && !fm.has_normals) {
        // Mesh::recompute_vertex_normals (src/render/mesh.cpp:283-345) runs after the loader applied to_world
        vertex_normals(fm.buf.positions, fm.buf.faces, fm.buf.normals);
        fm.has_normals = true;
    }
    return fm;
}

} // namespace

void FlatScene::finalize() {
    n_triangles = 0;
    for (size_t i = 0; i < meshes.size(); ++i) {
        dtof_mesh &m = meshes[i];
        MeshBuf &b = bufs[i];
        m.n_vertices = (uint32_t) (b.positions.size() / 3);
        m.n_faces = (uint32_t) (b.faces.size() / 3);
        m.positions = b.positions.data();
        m.normals = b.normals.empty() ? nullptr : b.normals.data();
        m.texcoords = b.texcoords.empty() ? nullptr : b.texcoords.data();
        m.faces = b.faces.data();
        n_triangles += m.n_faces;
    }
    desc.n_meshes = (uint32_t) meshes.size();
    desc.meshes = meshes.data();
    desc.n_instances = (uint32_t) instances.size();
    desc.instances = instances.data();
    desc.n_bsdfs = (uint32_t) bsdfs.size();
    desc.bsdfs = bsdfs.data();
    desc.n_emitters = (uint32_t) emitters.size();
    desc.emitters = emitters.data();
}

void FlatScene::serialize(const std::string &path) const {
    std::ofstream f(path, std::ios::binary);
    if (!f)
        throw Error("cannot write " + path);
    auto w32 = [&](uint32_t v) { f.write((const char *) &v, 4); };
    auto wf = [&](const float *p, size_t n) { f.write((const char *) p, (std::streamsize) (4 * n)); };
    f.write("DTOFDESC1\n", 10);
    w32(desc.n_meshes);
    for (uint32_t i = 0; i < desc.n_meshes; ++i) {
        const dtof_mesh &m = meshes[i];
        w32(m.n_vertices), w32(m.n_faces), w32(m.normals ? 1 : 0), w32(m.texcoords ? 1 : 0);
        w32(m.bsdf), w32((uint32_t) m.emitter), w32(m.flip_normals), w32(m.kind);
        wf(m.rect_to_world, 12);
        wf(m.positions, 3 * (size_t) m.n_vertices);
        if (m.normals) wf(m.normals, 3 * (size_t) m.n_vertices);
        if (m.texcoords) wf(m.texcoords, 2 * (size_t) m.n_vertices);
        f.write((const char *) m.faces, (std::streamsize) (12 * (size_t) m.n_faces));
    }
    w32(desc.n_instances);
    f.write((const char *) instances.data(), (std::streamsize) (instances.size() * sizeof(dtof_instance)));
    w32(desc.n_bsdfs);
    f.write((const char *) bsdfs.data(), (std::streamsize) (bsdfs.size() * sizeof(dtof_bsdf)));
    w32(desc.n_emitters);
    f.write((const char *) emitters.data(), (std::streamsize) (emitters.size() * sizeof(dtof_emitter)));
    f.write((const char *) &desc.camera, sizeof(dtof_camera));
    f.write((const char *) &desc.film, sizeof(dtof_film));
}

std::unique_ptr<FlatScene> Scene::flatten() const {
    auto fs = std::make_unique<FlatScene>();
    auto bsdf_id = [&](const Shape &s) -> uint32_t {
        Bsdf b = s.has_bsdf ? s.bsdf : Bsdf();   // Shape default BSDF: diffuse (src/render/shape.cpp:60-65)
        for (size_t i = 0; i < fs->bsdfs.size(); ++i) {
            const dtof_bsdf &o = fs->bsdfs[i];
            const bool c = b.kind != DTOF_BSDF_DIFFUSE;   // every other kind uses eta / k
            if (o.kind == b.kind && (o.twosided != 0) == b.twosided && o.reflectance[0] == b.reflectance[0] &&
                o.reflectance[1] == b.reflectance[1] && o.reflectance[2] == b.reflectance[2] &&
                (!c || (!memcmp(o.eta, b.eta, sizeof(o.eta)) && !memcmp(o.k, b.k, sizeof(o.k)))) &&
                o.alpha[0] == b.alpha[0] && o.alpha[1] == b.alpha[1] && o.distribution == b.distribution)
                return (uint32_t) i;
        }
        dtof_bsdf nb{};
        nb.kind = b.kind;
        nb.twosided = b.twosided ? 1u : 0u;
        memcpy(nb.reflectance, b.reflectance, sizeof(nb.reflectance));
        nb.alpha[0] = b.alpha[0], nb.alpha[1] = b.alpha[1], nb.distribution = b.distribution;
        if (b.kind != DTOF_BSDF_DIFFUSE) {
            memcpy(nb.eta, b.eta, sizeof(nb.eta));
            memcpy(nb.k, b.k, sizeof(nb.k));
        }
        fs->bsdfs.push_back(nb);
        return (uint32_t) fs->bsdfs.size() - 1;
    };
    auto push_mesh = [&](FlatMesh &&fm, uint32_t bsdf) {
        dtof_mesh m{};
        m.bsdf = bsdf;
        m.emitter = -1;
        m.flip_normals = fm.flip;
        m.kind = fm.kind;
        memcpy(m.rect_to_world, fm.rect_to_world, sizeof(m.rect_to_world));
        fs->meshes.push_back(m);
        fs->bufs.push_back(std::move(fm.buf));
    };
    const float ident12[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
    std::vector<int> mesh_of_shape(shapes.size(), -1);
    uint32_t n_static = 0;
    for (size_t i = 0; i < shapes.size(); ++i) {
        const Shape &s = shapes[i];
        if (s.animated())
            continue;
        Transform4 tw = Transform4::identity();
        if (s.has_static)
            tw = s.to_world;
        else if (s.has_anim && s.anim.size() == 1)
            tw = s.anim.transforms[0];
        uint32_t b = bsdf_id(s);
        mesh_of_shape[i] = (int) fs->meshes.size();
        push_mesh(flatten_shape(s, tw), b);
        ++n_static;
    }
    if (n_static) {
        dtof_instance in{};
        in.first_mesh = 0, in.n_meshes = n_static, in.animated = 0;
        memcpy(in.m0, ident12, sizeof(ident12));
        memcpy(in.m1, ident12, sizeof(ident12));
        fs->instances.push_back(in);
    }
    for (size_t i = 0; i < shapes.size(); ++i) {
        const Shape &s = shapes[i];
        if (!s.animated())
            continue;
        if (s.emitter)   // shapegroup.cpp:27-30
            throw Error("Instancing of emitters is not supported");
        uint32_t b = bsdf_id(s);
        // Instance::embree_geometry (instance.cpp:295-310): matrices at get_min_time / get_max_time
        float t0 = s.anim.min_time(), t1 = s.anim.max_time();
        float m0[16], m1[16];
        s.anim.eval(t0, m0);
        s.anim.eval(t1, m1);
        dtof_instance in{};
        in.first_mesh = (uint32_t) fs->meshes.size(), in.n_meshes = 1, in.animated = 1;
        in.t0 = t0, in.t1 = t1;
        memcpy(in.m0, m0, sizeof(in.m0));
        memcpy(in.m1, m1, sizeof(in.m1));
        fs->instances.push_back(in);
        mesh_of_shape[i] = (int) fs->meshes.size();
        push_mesh(flatten_shape(s, Transform4::identity()), b);
    }
    // Emitter order = order of appearance among the scene's children (scene.cpp:40-64)
    for (auto &o : order) {
        if (o.first == 's') {
            const Shape &s = shapes[o.second];
            if (!s.emitter)
                continue;
            int mi = mesh_of_shape[o.second];
            fs->meshes[mi].emitter = (int32_t) fs->emitters.size();
            dtof_emitter e{};
            e.kind = DTOF_EMITTER_AREA;
            e.mesh = (uint32_t) mi;
            memcpy(e.value, s.radiance, sizeof(e.value));
            fs->emitters.push_back(e);
        } else {
            const PointLight &pl = emitters[o.second];
            dtof_emitter e{};
            e.kind = pl.constant_env ? DTOF_EMITTER_CONSTANT : pl.spot ? DTOF_EMITTER_SPOT : pl.directional ? DTOF_EMITTER_DIRECTIONAL : DTOF_EMITTER_POINT;
            if (pl.spot) {
                memcpy(e.to_local, pl.to_local, sizeof(e.to_local));
                e.cutoff_angle = pl.cutoff_angle, e.beam_width = pl.beam_width;
            }
            if (pl.constant_env)
                for (const dtof_emitter &prev : fs->emitters)
                    if (prev.kind == DTOF_EMITTER_CONSTANT)
                        throw Error("Only one environment emitter can be specified per scene.");   // scene.cpp:53-55
            memcpy(e.position, pl.position, sizeof(e.position));
            memcpy(e.value, pl.intensity, sizeof(e.value));
            fs->emitters.push_back(e);
        }
    }
    fs->desc.camera = sensor.abi();
    fs->desc.film = sensor.film.abi();
    fs->finalize();
    return fs;
}

// ================================================================================================ XML loader
namespace {

struct Loader {
    std::map<std::string, std::string> defaults;
    std::set<std::string> cli, used;
    std::string base_dir;
    std::map<std::string, Bsdf> bsdfs;

    std::string sub(const std::string &s) {
        std::string o;
        for (size_t i = 0; i < s.size();) {
            if (s[i] == '$' && i + 1 < s.size() && (std::isalnum((unsigned char) s[i + 1]) || s[i + 1] == '_')) {
                size_t b = ++i;
                while (i < s.size() && (std::isalnum((unsigned char) s[i]) || s[i] == '_'))
                    ++i;
                std::string k = s.substr(b, i - b);
                auto it = defaults.find(k);
                if (it == defaults.end())
                    throw Error("undefined parameter $" + k);
                used.insert(k);
                o += it->second;
            } else {
                o += s[i++];
            }
        }
        return o;
    }
    bool has(const XmlNode &n, const char *name) { return n.attr(name) != nullptr; }
    std::string attr(const XmlNode &n, const char *name) {
        const std::string *v = n.attr(name);
        if (!v)
            throw Error("<" + n.tag + ">: missing attribute '" + name + "'");
        return sub(*v);
    }
    std::string attr(const XmlNode &n, const char *name, const std::string &dflt) {
        const std::string *v = n.attr(name);
        return v ? sub(*v) : dflt;
    }
    // textual property values keyed by name (value semantics are applied by the consumer)
    struct Prop {
        std::string tag, value;
        std::vector<double> vec;
    };
    std::map<std::string, Prop> props(const XmlNode &node) {
        std::map<std::string, Prop> out;
        for (auto &ch : node.children) {
            const std::string &t = ch->tag;
            if (t == "float" || t == "integer" || t == "boolean" || t == "string") {
                Prop p{ t, attr(*ch, "value"), {} };
                if (t == "float") parse_float(p.value);
                if (t == "integer") parse_int(p.value);
                if (t == "boolean") parse_bool(p.value);
                out[attr(*ch, "name")] = p;
            } else if (t == "rgb" || t == "spectrum") {
                if (!has(*ch, "value"))   // <spectrum filename=...>: a tabulated spectrum, not a constant
                    throw Error("<" + t + "> '" + attr(*ch, "name", "") + "' without a constant value is outside the supported subset");
                Prop p{ t, attr(*ch, "value"), {} };
                p.vec = parse_floats(p.value);
                if (p.vec.size() != 1 && p.vec.size() != 3)
                    throw Error("<" + t + "> '" + attr(*ch, "name") + "': only constant / RGB values are in scope");
                if (p.vec.size() == 1)
                    p.vec = { p.vec[0], p.vec[0], p.vec[0] };
                out[attr(*ch, "name")] = p;
            } else if (t == "point" || t == "vector") {
                Prop p{ t, "", vec(*ch, 0.0) };
                out[attr(*ch, "name")] = p;
            } else {
                // child objects are handled (or refused) by the code that walks the parent; a <texture>, a NAMED <ref> (a
                // texture / spectrum bound to a property) or an unknown tag would silently fall back to the plugin's
                // default value (0.5 reflectance, unit radiance ...): refuse it
                static const char *structural[] = { "transform", "animation", "bsdf", "emitter", "ref", "sampler", "film",
                                                    "rfilter", "shape", "sensor", "integrator", "default", "include", "alias" };
                bool ok = false;
                for (const char *k : structural)
                    ok = ok || t == k;
                if (t == "ref" && has(*ch, "name"))
                    ok = false;
                if (!ok)
                    throw Error("<" + t + (has(*ch, "name") ? " name='" + attr(*ch, "name") + "'" : std::string()) + "> inside <" +
                                node.tag + "> is outside the supported subset: only constant float / integer / boolean / string / "
                                "rgb / spectrum / point / vector properties are read (textures and spatially varying values are "
                                "not in scope)");
            }
        }
        return out;
    }
    std::vector<double> vec(const XmlNode &n, double dflt) {
        if (has(n, "value")) {
            auto v = parse_floats(attr(n, "value"));
            if (v.size() == 1)
                v = { v[0], v[0], v[0] };
            if (v.size() != 3)
                throw Error("<" + n.tag + ">: expected 1 or 3 values");
            return v;
        }
        std::vector<double> v(3, dflt);
        const char *names[3] = { "x", "y", "z" };
        for (int i = 0; i < 3; ++i)
            if (has(n, names[i]))
                v[i] = parse_float(attr(n, names[i]));
        return v;
    }
    Transform4 transform(const XmlNode &node) {
        Transform4 t = Transform4::identity();
        for (auto &op : node.children) {
            Transform4 o = Transform4::identity();
            if (op->tag == "matrix") {
                auto v = parse_floats(attr(*op, "value"));
                if (v.size() == 9) {
                    std::vector<double> m(16, 0.0);
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j)
                            m[4 * i + j] = v[3 * i + j];
                    m[15] = 1.0;
                    v = m;
                }
                if (v.size() != 16)
                    throw Error("matrix: expected 16 or 9 values");
                o = Transform4::from_matrix(v.data());
            } else if (op->tag == "translate") {
                auto v = vec(*op, 0.0);
                o = Transform4::translate(v[0], v[1], v[2]);
            } else if (op->tag == "scale") {
                auto v = vec(*op, 1.0);
                o = Transform4::scale(v[0], v[1], v[2]);
            } else if (op->tag == "rotate") {
                auto v = vec(*op, 0.0);
                o = Transform4::rotate(v[0], v[1], v[2], parse_float(attr(*op, "angle")));
            } else if (op->tag == "lookat" || op->tag == "look_at") {
                auto og = parse_floats(attr(*op, "origin")), tg = parse_floats(attr(*op, "target"));
                auto up = parse_floats(attr(*op, "up", "0,0,0"));
                if (og.size() != 3 || tg.size() != 3 || up.size() != 3 || (up[0] == 0 && up[1] == 0 && up[2] == 0))
                    throw Error("lookat without 'up' is outside the supported subset");
                o = Transform4::look_at(og.data(), tg.data(), up.data());
            } else {
                throw Error("unsupported transform op <" + op->tag + ">");
            }
            t = o * t;   // left-multiplication, xml.cpp:820-1007
        }
        return t;
    }
    AnimatedTransform animation(const XmlNode &node) {
        AnimatedTransform at;
        for (auto &kf : node.children) {
            if (kf->tag != "transform")
                throw Error("<animation> may only contain <transform time=..> keyframes");
            at.append(parse_float(attr(*kf, "time")), transform(*kf));
        }
        return at;
    }
    static void reject_unknown(const std::map<std::string, Prop> &p, const std::set<std::string> &known, const std::string &who) {
        for (auto &kv : p)
            if (!known.count(kv.first))
                throw Error(who + ": unreferenced property \"" + kv.first + "\"");
    }
    // RGB eta / k of the reference's named conductor materials (generated table, host/dtof_conductor_ior.inc)
    struct ConductorIor {
        const char *name;
        float eta[3], k[3];
    };
    static const ConductorIor *conductor_ior(const std::string &material) {
        static const ConductorIor table[] = {
#include "dtof_conductor_ior.inc"
        };
        for (const ConductorIor &e : table)
            if (material == e.name)
                return &e;
        throw Error("conductor material \"" + material + "\" is not in the table of named materials");
    }
    // lookup_ior (include/mitsuba/render/ior.h:23-98): a number, or a material of the table
    static float lookup_ior(const std::string &value) {
        static const std::pair<const char *, float> table[] = {
            { "vacuum", 1.0f }, { "helium", 1.000036f }, { "hydrogen", 1.000132f }, { "air", 1.000277f },
            { "carbon dioxide", 1.00045f }, { "water", 1.3330f }, { "acetone", 1.36f }, { "ethanol", 1.361f },
            { "carbon tetrachloride", 1.461f }, { "glycerol", 1.4729f }, { "benzene", 1.501f }, { "silicone oil", 1.52045f },
            { "bromine", 1.661f }, { "water ice", 1.31f }, { "fused quartz", 1.458f }, { "pyrex", 1.470f },
            { "acrylic glass", 1.49f }, { "polypropylene", 1.49f }, { "bk7", 1.5046f }, { "sodium chloride", 1.544f },
            { "amber", 1.55f }, { "pet", 1.5750f }, { "diamond", 2.419f } };
        std::string name;
        for (char c : value)
            name += (char) std::tolower((unsigned char) c);
        char *end = nullptr;
        float v = std::strtof(name.c_str(), &end);
        if (end != name.c_str() && *end == 0)
            return v;
        for (auto &e : table)
            if (name == e.first)
                return e.second;
        throw Error("Unable to find an IOR value for \"" + name + "\"!");
    }
    Bsdf bsdf(const XmlNode &node) {
        std::string typ = attr(node, "type");
        if (typ == "twosided") {
            const XmlNode *inner = nullptr;
            int n = 0;
            for (auto &c : node.children)
                if (c->tag == "bsdf" || c->tag == "ref") {
                    inner = c.get();
                    ++n;
                }
            if (n != 1)
                throw Error("twosided with two different BRDFs is outside the hot-path scope");
            Bsdf b = bsdf_or_ref(*inner);
            if (b.kind == DTOF_BSDF_DIELECTRIC || b.kind == DTOF_BSDF_THINDIELECTRIC || b.kind == DTOF_BSDF_ROUGHDIELECTRIC)   // twosided.cpp:102-103
                throw Error("Only materials without a transmission component can be nested!");
            b.twosided = true;
            return b;
        }
        if (typ == "diffuse") {
            auto p = props(node);
            reject_unknown(p, { "reflectance" }, "diffuse");
            Bsdf b;
            if (p.count("reflectance"))
                for (int i = 0; i < 3; ++i)
                    b.reflectance[i] = (float) p["reflectance"].vec[i];
            return b;
        }
        if (typ == "conductor") {   // SmoothConductor ctor, src/bsdfs/conductor.cpp:213-230
            auto p = props(node);
            reject_unknown(p, { "specular_reflectance", "material", "eta", "k" }, "conductor");
            const std::string material = p.count("material") ? p["material"].value : "none";
            const ConductorIor *named = nullptr;
            if (material != "none") {   // conductor.cpp:221-229: a named material replaces eta / k
                if (p.count("eta"))
                    throw Error("Should specify either (eta, k) or material, not both.");
                named = conductor_ior(material);
            }
            Bsdf b;
            b.kind = DTOF_BSDF_CONDUCTOR;
            for (int i = 0; i < 3; ++i) {
                b.reflectance[i] = p.count("specular_reflectance") ? (float) p["specular_reflectance"].vec[i] : 1.f;
                b.eta[i] = named ? named->eta[i] : p.count("eta") ? (float) p["eta"].vec[i] : 0.f;
                b.k[i] = named ? named->k[i] : p.count("k") ? (float) p["k"].vec[i] : 1.f;
            }
            return b;
        }
        if (typ == "roughconductor") {   // RoughConductor ctor, src/bsdfs/roughconductor.cpp:160-211
            auto p = props(node);
            reject_unknown(p, { "specular_reflectance", "material", "eta", "k", "distribution", "alpha", "alpha_u", "alpha_v", "sample_visible" },
                           "roughconductor");
            const std::string material = p.count("material") ? p["material"].value : "none";
            const ConductorIor *named = nullptr;
            if (material != "none") {   // conductor.cpp:221-229: a named material replaces eta / k
                if (p.count("eta"))
                    throw Error("Should specify either (eta, k) or material, not both.");
                named = conductor_ior(material);
            }
            const std::string distr = p.count("distribution") ? p["distribution"].value : "beckmann";
            if (distr != "beckmann" && distr != "ggx")
                throw Error("Specified an invalid distribution \"" + distr + "\", must be \"beckmann\" or \"ggx\"!");
            if (p.count("sample_visible") && !parse_bool(p["sample_visible"].value))
                throw Error("roughconductor with sample_visible=false is outside the hot-path scope");
            Bsdf b;
            b.kind = DTOF_BSDF_ROUGHCONDUCTOR;
            b.distribution = distr == "ggx" ? 1u : 0u;
            if (p.count("alpha_u") || p.count("alpha_v")) {
                if (!p.count("alpha_u") || !p.count("alpha_v"))
                    throw Error("Microfacet model: both 'alpha_u' and 'alpha_v' must be specified.");
                if (p.count("alpha"))
                    throw Error("Microfacet model: please specifyeither 'alpha' or 'alpha_u'/'alpha_v'.");
                b.alpha[0] = (float) parse_float(p["alpha_u"].value), b.alpha[1] = (float) parse_float(p["alpha_v"].value);
            } else {
                b.alpha[0] = b.alpha[1] = p.count("alpha") ? (float) parse_float(p["alpha"].value) : 0.1f;
            }
            for (int i = 0; i < 3; ++i) {
                b.reflectance[i] = p.count("specular_reflectance") ? (float) p["specular_reflectance"].vec[i] : 1.f;
                b.eta[i] = named ? named->eta[i] : p.count("eta") ? (float) p["eta"].vec[i] : 0.f;
                b.k[i] = named ? named->k[i] : p.count("k") ? (float) p["k"].vec[i] : 1.f;
            }
            return b;
        }
        if (typ == "roughdielectric") {   // RoughDielectric ctor, src/bsdfs/roughdielectric.cpp:161-213
            auto p = props(node);
            reject_unknown(p, { "int_ior", "ext_ior", "specular_reflectance", "specular_transmittance", "distribution", "alpha", "alpha_u",
                                "alpha_v", "sample_visible" }, "roughdielectric");
            const float int_ior = lookup_ior(p.count("int_ior") ? p["int_ior"].value : "bk7");
            const float ext_ior = lookup_ior(p.count("ext_ior") ? p["ext_ior"].value : "air");
            if (int_ior < 0.f || ext_ior < 0.f || int_ior == ext_ior)
                throw Error("The interior and exterior indices of refraction must be positive and differ!");
            std::string distr = p.count("distribution") ? p["distribution"].value : "beckmann";
            for (char &ch : distr)
                ch = (char) std::tolower((unsigned char) ch);
            if (distr != "beckmann" && distr != "ggx")
                throw Error("Specified an invalid distribution \"" + distr + "\", must be \"beckmann\" or \"ggx\"!");
            if (p.count("sample_visible") && !parse_bool(p["sample_visible"].value))
                throw Error("roughdielectric with sample_visible=false is outside the hot-path scope");
            Bsdf b;
            b.kind = DTOF_BSDF_ROUGHDIELECTRIC;
            b.distribution = distr == "ggx" ? 1u : 0u;
            if (p.count("alpha_u") || p.count("alpha_v")) {
                if (!p.count("alpha_u") || !p.count("alpha_v"))
                    throw Error("Microfacet model: both 'alpha_u' and 'alpha_v' must be specified.");
                if (p.count("alpha"))
                    throw Error("Microfacet model: please specifyeither 'alpha' or 'alpha_u'/'alpha_v'.");
                b.alpha[0] = (float) parse_float(p["alpha_u"].value), b.alpha[1] = (float) parse_float(p["alpha_v"].value);
            } else {
                b.alpha[0] = b.alpha[1] = p.count("alpha") ? (float) parse_float(p["alpha"].value) : 0.1f;
            }
            b.eta[0] = int_ior / ext_ior, b.eta[1] = b.eta[2] = 0.f;
            for (int i = 0; i < 3; ++i) {
                b.reflectance[i] = p.count("specular_reflectance") ? (float) p["specular_reflectance"].vec[i] : 1.f;
                b.k[i] = p.count("specular_transmittance") ? (float) p["specular_transmittance"].vec[i] : 1.f;
            }
            return b;
        }
        if (typ == "plastic") {   // SmoothPlastic ctor, src/bsdfs/plastic.cpp:157-183
            auto p = props(node);
            reject_unknown(p, { "int_ior", "ext_ior", "diffuse_reflectance", "specular_reflectance", "nonlinear" }, "plastic");
            const float int_ior = lookup_ior(p.count("int_ior") ? p["int_ior"].value : "polypropylene");
            const float ext_ior = lookup_ior(p.count("ext_ior") ? p["ext_ior"].value : "air");
            if (int_ior < 0.f || ext_ior < 0.f)
                throw Error("The interior and exterior indices of refraction must be positive!");
            Bsdf b;
            b.kind = DTOF_BSDF_PLASTIC;
            b.eta[0] = int_ior / ext_ior;
            b.eta[1] = p.count("nonlinear") && parse_bool(p["nonlinear"].value) ? 1.f : 0.f;
            b.eta[2] = 0.f;
            for (int i = 0; i < 3; ++i) {
                b.reflectance[i] = p.count("diffuse_reflectance") ? (float) p["diffuse_reflectance"].vec[i] : 0.5f;
                b.k[i] = p.count("specular_reflectance") ? (float) p["specular_reflectance"].vec[i] : 1.f;
            }
            return b;
        }
        if (typ == "dielectric" || typ == "thindielectric") {   // dielectric.cpp:199-228, thindielectric.cpp:104-126
            auto p = props(node);
            reject_unknown(p, { "int_ior", "ext_ior", "specular_reflectance", "specular_transmittance" }, typ.c_str());
            const float int_ior = lookup_ior(p.count("int_ior") ? p["int_ior"].value : "bk7");
            const float ext_ior = lookup_ior(p.count("ext_ior") ? p["ext_ior"].value : "air");
            if (int_ior < 0.f || ext_ior < 0.f)
                throw Error("The interior and exterior indices of refraction must be positive!");
            Bsdf b;
            b.kind = typ == "dielectric" ? DTOF_BSDF_DIELECTRIC : DTOF_BSDF_THINDIELECTRIC;
            b.eta[0] = int_ior / ext_ior, b.eta[1] = b.eta[2] = 0.f;
            for (int i = 0; i < 3; ++i) {
                b.reflectance[i] = p.count("specular_reflectance") ? (float) p["specular_reflectance"].vec[i] : 1.f;
                b.k[i] = p.count("specular_transmittance") ? (float) p["specular_transmittance"].vec[i] : 1.f;
            }
            return b;
        }
        throw Error("bsdf type '" + typ + "' is outside the hot-path scope (diffuse|conductor|roughconductor|dielectric|thindielectric|roughdielectric|plastic|twosided)");
    }
    Bsdf bsdf_or_ref(const XmlNode &node) {
        if (node.tag == "ref") {
            std::string id = attr(node, "id");
            auto it = bsdfs.find(id);
            if (it == bsdfs.end())
                throw Error("reference to unknown id '" + id + "'");
            return it->second;
        }
        return bsdf(node);
    }
    Shape shape(const XmlNode &node) {
        std::string typ = attr(node, "type");
        auto p = props(node);
        Shape sh;
        if (typ == "rectangle") sh.kind = Shape::Rectangle;
        else if (typ == "cube") sh.kind = Shape::Cube;
        else if (typ == "obj" || typ == "ply" || typ == "serialized") sh.kind = Shape::Mesh;
        else throw Error("shape type '" + typ + "' is outside the hot-path scope (rectangle|cube|obj|ply|serialized)");
        sh.id = attr(node, "id", "");
        if (p.count("flip_normals")) {
            sh.flip_normals = parse_bool(p["flip_normals"].value);
            p.erase("flip_normals");
        }
        for (auto &ch : node.children) {
            const std::string *nm = ch->attr("name");
            if (ch->tag == "transform" && nm && *nm == "to_world") {
                sh.to_world = transform(*ch);
                sh.has_static = true;
                sh.has_anim = false;
            } else if (ch->tag == "animation" && nm && *nm == "to_world") {
                sh.anim = animation(*ch);
                sh.has_anim = true;
                sh.has_static = false;
            } else if (ch->tag == "bsdf" || ch->tag == "ref") {
                sh.bsdf = bsdf_or_ref(*ch);
                sh.has_bsdf = true;
            } else if (ch->tag == "emitter") {
                if (attr(*ch, "type") != "area")
                    throw Error("only 'area' emitters can be attached to shapes");
                auto ep = props(*ch);
                sh.emitter = true;
                if (ep.count("radiance"))
                    for (int i = 0; i < 3; ++i)
                        sh.radiance[i] = (float) ep["radiance"].vec[i];
            }
        }
        if (typ == "obj" || typ == "ply" || typ == "serialized") {
            if (!p.count("filename"))
                throw Error("shape '" + typ + "': missing 'filename'");
            std::string fn = p["filename"].value;
            p.erase("filename");
            bool face_normals = false;
            if (p.count("face_normals")) {
                face_normals = parse_bool(p["face_normals"].value);
                p.erase("face_normals");
            }
            std::string path = (!fn.empty() && fn[0] == '/') ? fn : base_dir + "/" + fn;
            if (typ == "serialized") {
                int shape_index = 0;
                if (p.count("shape_index")) {
                    shape_index = (int) parse_int(p["shape_index"].value);
                    p.erase("shape_index");
                }
                load_serialized_file(path, shape_index, face_normals, sh.positions, sh.faces, sh.normals, sh.texcoords, false);
            } else {
                bool flip_uv = true;   // obj.cpp:151 (ply.cpp has no such property)
                if (typ == "obj" && p.count("flip_tex_coords")) {
                    flip_uv = parse_bool(p["flip_tex_coords"].value);
                    p.erase("flip_tex_coords");
                }
                load_mesh_file(path, face_normals, sh.positions, sh.faces, sh.normals, sh.texcoords, false, flip_uv);
            }
            if (face_normals)
                sh.normals.clear();
            sh.smooth_normals = sh.normals.empty() && !face_normals;   // computed at flatten time, after to_world
        }
        if (!p.empty())
            throw Error("shape '" + typ + "': unreferenced property \"" + p.begin()->first + "\"");
        return sh;
    }
    PerspectiveSensor sensor(const XmlNode &node) {
        if (attr(node, "type") != "perspective")
            throw Error("only the 'perspective' sensor is in the hot-path scope");
        auto p = props(node);
        PerspectiveSensor s;
        // a sensor without a <sampler> gets the reference's default: `independent`, 4 spp (src/render/sensor.cpp:47-48)
        s.sampler.correlated = false;
        s.sampler.time_correlate_number = s.sampler.path_correlate_number = 1;
        for (auto &ch : node.children) {
            const std::string *nm = ch->attr("name");
            if (ch->tag == "transform" && nm && *nm == "to_world") {
                s.to_world = transform(*ch);
            } else if (ch->tag == "sampler") {
                const std::string skind = attr(*ch, "type");
                if (skind != "correlated" && skind != "independent")
                    throw Error("sampler '" + skind + "' is outside the hot-path scope (correlated; independent for path/velocity)");
                auto sp = props(*ch);
                CorrelatedSampler cs;
                if (skind == "independent") {   // PCG32Sampler only: the stream `path` / `velocity` draw from (sampler.cpp:115-134)
                    reject_unknown(sp, { "sample_count", "seed" }, "independent sampler");
                    cs.correlated = false;
                    cs.time_correlate_number = 1;
                }
                // use_stratified_sampling_for_each_interval is an INTEGRATOR property (SURVEY.md finding 7)
                reject_unknown(sp, { "sample_count", "seed", "time_correlate_number", "path_correlate_number" }, "correlated sampler");
                if (sp.count("sample_count")) cs.sample_count = (uint32_t) parse_int(sp["sample_count"].value);
                if (sp.count("seed")) cs.seed = (uint32_t) parse_int(sp["seed"].value);
                if (sp.count("time_correlate_number")) cs.time_correlate_number = (uint32_t) parse_int(sp["time_correlate_number"].value);
                cs.path_correlate_number = sp.count("path_correlate_number") ? (uint32_t) parse_int(sp["path_correlate_number"].value)
                                                                             : cs.time_correlate_number;
                s.sampler = cs;
            } else if (ch->tag == "film") {
                auto fp = props(*ch);
                reject_unknown(fp, { "width", "height", "crop_width", "crop_height", "crop_offset_x", "crop_offset_y", "file_format",
                                     "pixel_format", "component_format", "sample_border", "compensate" }, "hdrfilm");
                // sample_border (film.cpp:35, integrator.cpp:176-178) would change the border pixels if ignored: refuse
                if (fp.count("sample_border") && parse_bool(fp["sample_border"].value))
                    throw Error("hdrfilm: sample_border=true is outside the hot-path scope");
                if (fp.count("pixel_format") && fp["pixel_format"].value != "rgb")
                    throw Error("hdrfilm: only pixel_format=\"rgb\" is inside the hot-path scope");
                Film f;
                if (fp.count("width")) f.width = (uint32_t) parse_int(fp["width"].value);
                if (fp.count("height")) f.height = (uint32_t) parse_int(fp["height"].value);
                if (fp.count("crop_width") || fp.count("crop_height")) {
                    f.has_crop = true;
                    f.crop_w = fp.count("crop_width") ? (uint32_t) parse_int(fp["crop_width"].value) : f.width;
                    f.crop_h = fp.count("crop_height") ? (uint32_t) parse_int(fp["crop_height"].value) : f.height;
                    f.crop_x = fp.count("crop_offset_x") ? (uint32_t) parse_int(fp["crop_offset_x"].value) : 0;
                    f.crop_y = fp.count("crop_offset_y") ? (uint32_t) parse_int(fp["crop_offset_y"].value) : 0;
                }
                for (auto &rf : ch->children)
                    if (rf->tag == "rfilter") {
                        f.rfilter = attr(*rf, "type");
                        auto rp = props(*rf);
                        if (rp.count("radius")) {
                            f.has_radius = true;
                            f.radius = parse_float(rp["radius"].value);
                        }
                        if (rp.count("stddev"))
                            f.stddev = parse_float(rp["stddev"].value);
                        if (rp.count("B"))
                            f.mitchell_b = parse_float(rp["B"].value);
                        if (rp.count("C"))
                            f.mitchell_c = parse_float(rp["C"].value);
                        if (rp.count("lobes"))
                            f.lanczos_lobes = (int) parse_int(rp["lobes"].value);
                        for (auto &kv : rp) {
                            const std::string &k = kv.first;
                            const bool ok = (f.rfilter == "tent" && k == "radius") || (f.rfilter == "gaussian" && k == "stddev") ||
                                            (f.rfilter == "mitchell" && (k == "B" || k == "C")) || (f.rfilter == "lanczos" && k == "lobes");
                            if (!ok)
                                throw Error("rfilter '" + f.rfilter + "': unreferenced property \"" + k + "\"");
                        }
                    }
                s.film = f;
            }
        }
        reject_unknown(p, { "fov", "fov_axis", "near_clip", "far_clip", "shutter_open", "shutter_close" }, "perspective");
        if (p.count("fov")) s.fov = parse_float(p["fov"].value);
        if (p.count("fov_axis")) s.fov_axis = p["fov_axis"].value;
        if (p.count("near_clip")) s.near_clip = parse_float(p["near_clip"].value);
        if (p.count("far_clip")) s.far_clip = parse_float(p["far_clip"].value);
        if (p.count("shutter_open")) s.shutter_open = parse_float(p["shutter_open"].value);
        if (p.count("shutter_close")) s.shutter_close = parse_float(p["shutter_close"].value);
        return s;
    }
    Scene load(const XmlNode &root) {
        if (root.tag != "scene")
            throw Error("root element must be <scene>");
        Scene sc;
        bool have_integrator = false;
        for (auto &node : root.children) {
            if (node->tag == "default") {
                std::string k = attr(*node, "name");
                if (!cli.count(k))
                    defaults[k] = sub(*node->attr("value"));
            } else if (node->tag == "integrator") {
                std::string typ = attr(*node, "type");
                if (typ != "dopplertofpath" && typ != "velocity" && typ != "path")
                    throw Error("integrator '" + typ + "' is outside the hot-path scope (dopplertofpath|velocity|path)");
                std::map<std::string, std::string> ip;
                for (auto &kv : props(*node))
                    ip[kv.first] = kv.second.value;
                sc.integrator = DopplerToFPathIntegrator(ip, typ == "velocity" ? DTOF_INTEGRATOR_VELOCITY
                                                             : typ == "path" ? DTOF_INTEGRATOR_PATH : DTOF_INTEGRATOR_DOPPLERTOFPATH);
                have_integrator = true;
            } else if (node->tag == "sensor") {
                sc.sensor = sensor(*node);
            } else if (node->tag == "bsdf") {
                Bsdf b = bsdf(*node);
                if (has(*node, "id"))
                    bsdfs[attr(*node, "id")] = b;
            } else if (node->tag == "shape") {
                sc.order.emplace_back('s', (uint32_t) sc.shapes.size());
                sc.shapes.push_back(shape(*node));
            } else if (node->tag == "emitter") {
                std::string typ = attr(*node, "type");
                if (typ != "point" && typ != "constant" && typ != "spot" && typ != "directional")
                    throw Error("emitter '" + typ + "' is outside the hot-path scope (point|spot|directional|area|constant)");
                auto p = props(*node);
                PointLight pl;
                if (typ == "directional") {   // DirectionalEmitter ctor, src/emitters/directional.cpp:65-91
                    pl.directional = true;
                    for (auto &kv : p)
                        if (kv.first != "irradiance" && kv.first != "direction")
                            throw Error("emitter 'directional': unreferenced property \"" + kv.first + "\"");
                    bool have_tw = false;
                    Transform4 tw = Transform4::identity();
                    for (auto &ch : node->children) {
                        const std::string *nm = ch->attr("name");
                        if (ch->tag == "transform" && nm && *nm == "to_world")
                            tw = transform(*ch), have_tw = true;
                    }
                    if (p.count("direction")) {
                        if (have_tw)
                            throw Error("Only one of the parameters 'direction' and 'to_world' can be specified at the same time!'");
                        float d[3] = { (float) p["direction"].vec[0], (float) p["direction"].vec[1], (float) p["direction"].vec[2] };
                        for (int rep = 0; rep < 2; ++rep) {   // normalize(direction); look_at(0, direction, up) normalises again
                            const float inv = 1.f / std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                            for (int i = 0; i < 3; ++i)
                                d[i] *= inv;
                        }
                        memcpy(pl.position, d, sizeof(d));
                    } else {
                        for (int i = 0; i < 3; ++i)
                            pl.position[i] = (float) tw.m[4 * i + 2];   // to_world.transform_affine((0, 0, 1))
                    }
                    if (p.count("irradiance"))
                        for (int i = 0; i < 3; ++i)
                            pl.intensity[i] = (float) p["irradiance"].vec[i];
                    sc.order.emplace_back('e', (uint32_t) sc.emitters.size());
                    sc.emitters.push_back(pl);
                    continue;
                }
                if (typ == "spot") {   // SpotLight ctor, src/emitters/spot.cpp:89-114
                    pl.spot = true;
                    for (auto &kv : p)
                        if (kv.first != "intensity" && kv.first != "cutoff_angle" && kv.first != "beam_width")
                            throw Error("emitter 'spot': unreferenced property \"" + kv.first + "\" (projection textures are out of scope)");
                    Transform4 tw = Transform4::identity();
                    for (auto &ch : node->children) {
                        const std::string *nm = ch->attr("name");
                        if (ch->tag == "transform" && nm && *nm == "to_world")
                            tw = transform(*ch);
                    }
                    for (int i = 0; i < 3; ++i) {
                        pl.position[i] = (float) tw.m[4 * i + 3];
                        for (int j = 0; j < 3; ++j)
                            pl.to_local[3 * i + j] = (float) tw.it[4 * j + i];   // inverse = (inverse transpose)^T
                        if (p.count("intensity"))
                            pl.intensity[i] = (float) p["intensity"].vec[i];
                    }
                    const float cutoff = p.count("cutoff_angle") ? (float) parse_float(p["cutoff_angle"].value) : 20.f;
                    const float beam = p.count("beam_width") ? (float) parse_float(p["beam_width"].value) : cutoff * 3.f / 4.f;
                    const float deg = (float) (M_PI / 180.0);
                    pl.cutoff_angle = cutoff * deg, pl.beam_width = beam * deg;
                    sc.order.emplace_back('e', (uint32_t) sc.emitters.size());
                    sc.emitters.push_back(pl);
                    continue;
                }
                if (typ == "constant") {
                    pl.constant_env = true;
                    for (auto &kv : p)
                        if (kv.first != "radiance")
                            throw Error("emitter 'constant': unreferenced property \"" + kv.first + "\"");
                    if (p.count("radiance"))
                        for (int i = 0; i < 3; ++i)
                            pl.intensity[i] = (float) p["radiance"].vec[i];
                    sc.order.emplace_back('e', (uint32_t) sc.emitters.size());
                    sc.emitters.push_back(pl);
                    continue;
                }
                bool have_pos = p.count("position") != 0;
                if (have_pos)
                    for (int i = 0; i < 3; ++i)
                        pl.position[i] = (float) p["position"].vec[i];
                for (auto &ch : node->children) {
                    const std::string *nm = ch->attr("name");
                    if (ch->tag == "transform" && nm && *nm == "to_world") {
                        if (have_pos)
                            throw Error("Only one of the parameters 'position' and 'to_world' can be specified");
                        Transform4 t = transform(*ch).narrowed();
                        pl.position[0] = (float) t.m[3], pl.position[1] = (float) t.m[7], pl.position[2] = (float) t.m[11];
                    }
                }
                if (p.count("intensity"))
                    for (int i = 0; i < 3; ++i)
                        pl.intensity[i] = (float) p["intensity"].vec[i];
                sc.order.emplace_back('e', (uint32_t) sc.emitters.size());
                sc.emitters.push_back(pl);
            } else {
                throw Error("unsupported top-level element <" + node->tag + ">");
            }
        }
        for (auto &k : cli)
            if (!used.count(k))
                throw Error("Unused parameter \"" + k + "\"!");   // xml.cpp:1069
        if (!have_integrator)
            throw Error("scene has no integrator");
        return sc;
    }
};

std::string dirname_of(const std::string &path) {
    size_t p = path.find_last_of('/');
    return p == std::string::npos ? "." : path.substr(0, p);
}

} // namespace

Scene load_string(const std::string &xml, const std::string &base_dir, const std::map<std::string, std::string> &params) {
    Loader L;
    L.defaults = params;
    for (auto &kv : params)
        L.cli.insert(kv.first);
    L.base_dir = base_dir;
    auto root = parse_xml(xml);
    return L.load(*root);
}

Scene load_file(const std::string &path, const std::map<std::string, std::string> &params) {
    std::ifstream f(path, std::ios::binary);
    if (!f)
        throw Error("\"" + path + "\": file does not exist!");
    std::stringstream ss;
    ss << f.rdbuf();
    return load_string(ss.str(), dirname_of(path), params);
}

// ================================================================================================ renderer
Renderer::Renderer(int device) {
    dtof_status s = dtof_create(&ctx, device);
    if (s != DTOF_OK)
        throw Error("dtof_create(device=" + std::to_string(device) + ") failed with status " + std::to_string((int) s) +
                    ": no usable CUDA device (the product path has no CPU fallback)");
}
Renderer::~Renderer() { dtof_destroy(ctx); }
void Renderer::check(dtof_status s, const char *what) {
    if (s != DTOF_OK)
        throw Error(std::string(what) + ": [status " + std::to_string((int) s) + "] " + dtof_last_error(ctx));
}
void Renderer::upload(const FlatScene &flat) { check(dtof_upload_scene(ctx, &flat.desc), "dtof_upload_scene"); }
std::vector<float> Renderer::render(const dtof_params &p, bool develop, uint32_t width, uint32_t height) {
    std::vector<float> out((size_t) width * height * (develop ? 3 : 4));
    check(dtof_render(ctx, &p, develop ? nullptr : out.data(), develop ? out.data() : nullptr), "dtof_render");
    return out;
}
float Renderer::last_kernel_ms() {
    float ms = 0.f;
    check(dtof_last_kernel_ms(ctx, &ms), "dtof_last_kernel_ms");
    return ms;
}

void write_pfm(const std::string &path, const float *data, uint32_t w, uint32_t h, uint32_t channels) {
    std::ofstream f(path, std::ios::binary);
    if (!f)
        throw Error("cannot write " + path);
    f << "PF\n" << w << " " << h << "\n-1.0\n";   // little endian, bottom-to-top scanlines, RGB
    std::vector<float> row(3 * (size_t) w);
    for (uint32_t y = 0; y < h; ++y) {
        const float *src = data + (size_t) (h - 1 - y) * w * channels;
        for (uint32_t x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c)
                row[3 * x + c] = src[(size_t) x * channels + c];
        f.write((const char *) row.data(), (std::streamsize) (row.size() * 4));
    }
}

void write_npy(const std::string &path, const float *data, uint32_t h, uint32_t w, uint32_t c) {
    std::ofstream f(path, std::ios::binary);
    if (!f)
        throw Error("cannot write " + path);
    std::string hdr = "{'descr': '<f4', 'fortran_order': False, 'shape': (" + std::to_string(h) + ", " + std::to_string(w) + ", " +
                      std::to_string(c) + "), }";
    size_t total = 10 + hdr.size() + 1;
    hdr += std::string((64 - total % 64) % 64, ' ');
    hdr += '\n';
    const unsigned char magic[8] = { 0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0 };
    f.write((const char *) magic, 8);
    uint16_t len = (uint16_t) hdr.size();
    f.write((const char *) &len, 2);
    f << hdr;
    f.write((const char *) data, (std::streamsize) ((size_t) h * w * c * 4));
}

} // namespace dtof_host
