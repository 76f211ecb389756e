// dtof_render -- `mitsuba`-like command line front end of the B200 Doppler-ToF path
// (the reference's src/mitsuba/mitsuba.cpp:150-360 for this one integrator):
//
//   dtof_render [-D key=value]... [-o out.pfm|out.npy] [-s seed] [--spp n] [--device i] [--raw]
//               [--dump-desc file] [--info] scene.xml
//
// loads the scene XML unchanged (defaults overridable with -D like the reference CLI), flattens it, uploads it
// through the C ABI (include/dtof.h) and writes the developed image (or, with --raw, the RGBW accumulation tensor).
// --dump-desc / --info need no GPU: they stop after the host half (flattening / validation + BVH build).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dtof_host.hpp"

using namespace dtof_host;

static void usage() {
    fprintf(stderr, "usage: dtof_render [-D key=value]... [-o out.pfm|out.npy] [-s seed] [--spp n] [--device i] [--raw]\n"
                    "                   [--dump-desc file] [--info] scene.xml\n");
}

int main(int argc, char **argv) {
    std::map<std::string, std::string> params;
    std::string out = "", scene_path, dump;
    uint32_t seed = 0, spp = 0;
    int device = 0;
    bool raw = false, info = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) {
                usage();
                exit(2);
            }
            return argv[++i];
        };
        if (a == "-D" || (a.rfind("-D", 0) == 0 && a.size() > 2)) {
            std::string kv = a == "-D" ? next() : a.substr(2);
            size_t eq = kv.find('=');
            if (eq == std::string::npos) {
                fprintf(stderr, "-D expects key=value\n");
                return 2;
            }
            params[kv.substr(0, eq)] = kv.substr(eq + 1);
        } else if (a == "-o")
            out = next();
        else if (a == "-s")
            seed = (uint32_t) std::stoul(next());
        else if (a == "--spp")
            spp = (uint32_t) std::stoul(next());
        else if (a == "--device")
            device = std::stoi(next());
        else if (a == "--raw")
            raw = true;
        else if (a == "--info")
            info = true;
        else if (a == "--dump-desc")
            dump = next();
        else if (a == "-h" || a == "--help") {
            usage();
            return 0;
        } else if (!a.empty() && a[0] == '-') {
            fprintf(stderr, "unknown option %s\n", a.c_str());
            return 2;
        } else
            scene_path = a;
    }
    if (scene_path.empty()) {
        usage();
        return 2;
    }
    try {
        Scene scene = load_file(scene_path, params);
        auto flat = scene.flatten();
        if (!dump.empty())
            flat->serialize(dump);
        if (info) {
            dtof_scene_info si;
            char err[512] = "";
            if (dtof_scene_info_for(&flat->desc, &si, err, sizeof(err)) != DTOF_OK)
                throw Error(err);
            printf("triangles %u  bvh nodes %u  instances %u  depth %u  traversal data %.3f MB  shading data %.3f MB  build %.1f ms\n",
                   si.n_triangles, si.n_nodes, si.n_instances, si.bvh_depth, si.traversal_bytes / 1e6, si.shading_bytes / 1e6,
                   si.build_ms);
        }
        if ((info || !dump.empty()) && out.empty())
            return 0;
        if (out.empty()) {
            size_t dot = scene_path.find_last_of('.');
            out = scene_path.substr(0, dot) + ".pfm";
        }
        Renderer r(device);
        r.upload(*flat);
        dtof_params p = scene.integrator.params(scene.sensor.sampler, seed, spp);
        const uint32_t W = flat->desc.film.width, H = flat->desc.film.height;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<float> img = r.render(p, !raw, W, H);
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        // same wording as the reference's log line (src/render/integrator.cpp:342-344)
        printf("Rendering finished. (took %.3fs; kernel %.3f ms; %.1f Msamples/s)\n", secs, r.last_kernel_ms(),
               (double) W * H * p.sample_count / secs / 1e6);
        uint32_t ch = raw ? 4 : 3;
        if (out.size() > 4 && out.compare(out.size() - 4, 4, ".npy") == 0)
            write_npy(out, img.data(), H, W, ch);
        else
            write_pfm(out, img.data(), W, H, ch);
    } catch (const std::exception &e) {
        fprintf(stderr, "Caught a critical exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
