// TEST / DATA INFRASTRUCTURE (runs only in the build container): prints the RGB complex indices of refraction the
// REFERENCE derives for its named conductor materials (src/bsdfs/conductor.cpp:213-230 -> complex_ior_from_file,
// include/mitsuba/render/ior.h:100-143: measured spectra in resources/data/ior/*.spd converted by
// spectrum_list_to_srgb). The product's hosts look material names up in the resulting table
// (mitsuba3dopplertof_b200/conductor_ior.json, host/dtof_conductor_ior.inc) instead of shipping the spectra.
// usage: conductor_table <reference resources dir> name [name ...]   -> JSON on stdout
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/filesystem.h>
#include <mitsuba/core/fresolver.h>
#include <mitsuba/core/jit.h>
#include <mitsuba/core/logger.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/profiler.h>
#include <mitsuba/core/properties.h>
#include <mitsuba/core/spectrum.h>
#include <mitsuba/core/thread.h>
#include <mitsuba/render/bsdf.h>
#include <mitsuba/render/interaction.h>
#include <mitsuba/render/texture.h>

#include <cstdio>
#include <string>
#include <vector>

namespace mi = mitsuba;
namespace dr = drjit;
using F = float;
using S = mi::Color<float, 3>;

struct Collector : mi::TraversalCallback {
    std::vector<std::pair<std::string, mi::Object *>> objects;
    void put_parameter_impl(const std::string &, void *, uint32_t, const std::type_info &) override {}
    void put_object(const std::string &name, mi::Object *obj, uint32_t) override { objects.emplace_back(name, obj); }
};

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: conductor_table <resources dir> name [name ...]\n");
        return 2;
    }
    mi::Jit::static_initialization();
    mi::Class::static_initialization();
    mi::Thread::static_initialization();
    mi::Logger::static_initialization();
    mi::Bitmap::static_initialization();
    mi::Profiler::static_initialization();
    mi::Thread::thread()->logger()->set_log_level(mi::Error);
    mi::color_management_static_initialization(false, false);
    mi::ref<mi::FileResolver> fr = mi::Thread::thread()->file_resolver();
    const char *ref_dir = getenv("DTOF_REF_DIR");
    fr->append(ref_dir ? mi::fs::path(ref_dir) : mi::fs::path(argv[0]).parent_path());
    fr->append(mi::fs::path(argv[1]));
    std::string out = "{\n";
    bool first = true;
    for (int i = 2; i < argc; ++i) {
        try {
            mi::Properties props("conductor");
            props.set_string("material", argv[i]);
            mi::ref<mi::BSDF<F, S>> bsdf = mi::PluginManager::instance()->create_object<mi::BSDF<F, S>>(props);
            Collector c;
            bsdf->traverse(&c);
            mi::SurfaceInteraction<F, S> si = dr::zeros<mi::SurfaceInteraction<F, S>>();
            S eta(0.f), k(0.f);
            for (auto &o : c.objects) {
                auto *t = (const mi::Texture<F, S> *) o.second;
                if (o.first == "eta")
                    eta = t->eval(si);
                else if (o.first == "k")
                    k = t->eval(si);
            }
            char buf[512];
            snprintf(buf, sizeof(buf), "%s  \"%s\": {\"eta\": [%.9g, %.9g, %.9g], \"k\": [%.9g, %.9g, %.9g]}", first ? "" : ",\n", argv[i],
                     eta[0], eta[1], eta[2], k[0], k[1], k[2]);
            out += buf;
            first = false;
        } catch (const std::exception &e) {   // a few spectra are all zero: the reference itself cannot load them
            fprintf(stderr, "%s: the reference throws: %s\n", argv[i], e.what());
        }
    }
    out += "\n}\n";
    fputs(out.c_str(), stdout);
    return 0;
}
