// TEST INFRASTRUCTURE (runs only in the build container): evaluates the REFERENCE's own reconstruction-filter plugins
// (src/rfilters/{box,tent,gaussian,mitchell,catmullrom,lanczos}.cpp, scalar_rgb) at a fixed set of offsets and prints
// radius + values as JSON; tests/golden/rfilter_vectors.json pins the oracle's FilmSplat::eval with them.
// usage: rfilter_vectors  -> JSON on stdout
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/filesystem.h>
#include <mitsuba/core/fresolver.h>
#include <mitsuba/core/jit.h>
#include <mitsuba/core/logger.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/profiler.h>
#include <mitsuba/core/properties.h>
#include <mitsuba/core/rfilter.h>
#include <mitsuba/core/spectrum.h>
#include <mitsuba/core/thread.h>

#include <cstdio>
#include <string>
#include <vector>

namespace mi = mitsuba;
using F = float;
using S = mi::Color<float, 3>;

int main(int, char **argv) {
    mi::Jit::static_initialization();
    mi::Class::static_initialization();
    mi::Thread::static_initialization();
    mi::Logger::static_initialization();
    mi::Bitmap::static_initialization();
    mi::Profiler::static_initialization();
    mi::Thread::thread()->logger()->set_log_level(mi::Error);
    mi::ref<mi::FileResolver> fr = mi::Thread::thread()->file_resolver();
    const char *ref_dir = getenv("DTOF_REF_DIR");
    fr->append(ref_dir ? mi::fs::path(ref_dir) : mi::fs::path(argv[0]).parent_path());

    struct Case { const char *name, *type; const char *fkey[2]; float fval[2]; const char *ikey; int ival; };
    const Case cases[] = {
        { "tent", "tent", { nullptr, nullptr }, { 0, 0 }, nullptr, 0 },
        { "tent_r2.5", "tent", { "radius", nullptr }, { 2.5f, 0 }, nullptr, 0 },
        { "gaussian", "gaussian", { nullptr, nullptr }, { 0, 0 }, nullptr, 0 },
        { "gaussian_s0.8", "gaussian", { "stddev", nullptr }, { 0.8f, 0 }, nullptr, 0 },
        { "mitchell", "mitchell", { nullptr, nullptr }, { 0, 0 }, nullptr, 0 },
        { "mitchell_b0.2_c0.6", "mitchell", { "B", "C" }, { 0.2f, 0.6f }, nullptr, 0 },
        { "catmullrom", "catmullrom", { nullptr, nullptr }, { 0, 0 }, nullptr, 0 },
        { "lanczos", "lanczos", { nullptr, nullptr }, { 0, 0 }, nullptr, 0 },
        { "lanczos_l2", "lanczos", { nullptr, nullptr }, { 0, 0 }, "lobes", 2 },
        { "lanczos_l5", "lanczos", { nullptr, nullptr }, { 0, 0 }, "lobes", 5 },
    };
    std::vector<float> xs = { 0.f, 1e-8f, -1e-8f, 1e-6f };
    for (int i = -240; i <= 240; ++i)
        xs.push_back((float) i * (1.f / 43.f) + 0.00137f);
    for (float x : { 1.f, -1.f, 2.f, -2.f, 3.f, 5.f, 0.5f, -0.5f, 1.9999999f, 2.0000002f, 2.9999998f })
        xs.push_back(x);
    std::string out = "{\n  \"x\": [";
    char buf[64];
    for (size_t i = 0; i < xs.size(); ++i) {
        snprintf(buf, sizeof(buf), "%s%.9g", i ? ", " : "", xs[i]);
        out += buf;
    }
    out += "],\n  \"filters\": {\n";
    bool first = true;
    for (const Case &c : cases) {
        mi::Properties props(c.type);
        for (int k = 0; k < 2; ++k)
            if (c.fkey[k])
                props.set_float(c.fkey[k], c.fval[k]);
        if (c.ikey)
            props.set_int(c.ikey, c.ival);
        mi::ref<mi::ReconstructionFilter<F, S>> rf = mi::PluginManager::instance()->create_object<mi::ReconstructionFilter<F, S>>(props);
        snprintf(buf, sizeof(buf), "%.9g", rf->radius());
        out += std::string(first ? "" : ",\n") + "    \"" + c.name + "\": {\"type\": \"" + c.type + "\", \"radius\": " + buf +
               ", \"describe\": \"" + rf->to_string() + "\", \"props\": {";
        bool pf = true;
        for (int k = 0; k < 2; ++k)
            if (c.fkey[k]) {
                snprintf(buf, sizeof(buf), "%s\"%s\": %.9g", pf ? "" : ", ", c.fkey[k], c.fval[k]);
                out += buf;
                pf = false;
            }
        if (c.ikey) {
            snprintf(buf, sizeof(buf), "%s\"%s\": %d", pf ? "" : ", ", c.ikey, c.ival);
            out += buf;
        }
        out += "}, \"y\": [";
        for (size_t i = 0; i < xs.size(); ++i) {
            snprintf(buf, sizeof(buf), "%s%.9g", i ? ", " : "", rf->eval(xs[i]));
            out += buf;
        }
        out += "]}";
        first = false;
    }
    out += "\n  }\n}\n";
    fputs(out.c_str(), stdout);
    return 0;
}
