// TEST INFRASTRUCTURE (fixture generator): known-answer vectors computed by the REFERENCE's own
// header templates with scalar float types. Writes JSON to stdout -> tests/golden/header_vectors.json.
// Functions exercised (reference file:line):
//   sample_tea_32                 include/mitsuba/core/random.h:77-90
//   PCG32                         ext/drjit/include/drjit/random.h:55-137
//   permute_kensler               include/mitsuba/core/random.h:235-292
//   dr::sincos / dr::fmod         ext/drjit/include/drjit/math.h:76-176, array_router.h:484-486
//   eval_modulation_function_*    include/mitsuba/render/waveform_utils.h:24-62
//   square_to_cosine_hemisphere   include/mitsuba/core/warp.h:54-89,320-330
//   square_to_uniform_triangle    include/mitsuba/core/warp.h:153-156
//   coordinate_system             include/mitsuba/core/vector.h:116-136
//   perspective_projection        include/mitsuba/render/sensor.h:227-262 (+ Transform::inverse)
//   TentFilter weights            restated inline from src/rfilters/tent.cpp:53-55 (plugin, not a header)
#include <mitsuba/core/fwd.h>
#include <mitsuba/core/random.h>
#include <mitsuba/core/transform.h>
#include <mitsuba/core/vector.h>
#include <mitsuba/core/warp.h>
#include <mitsuba/render/sensor.h>
#include <mitsuba/render/waveform_utils.h>

#include <cstdio>
#include <vector>

namespace mi = mitsuba;
namespace dr = drjit;

static void pf(float f) { printf("%.9g", f); }

int main() {
    printf("{\n");
    // ---- TEA
    printf("\"tea32\": [");
    {
        uint32_t in[][2] = { { 0, 0 }, { 1, 1 }, { 0, 33685504u }, { 1, 16842752u }, { 2, 16842752u }, { 7, 0xffffffffu },
                             { 0xdeadbeefu, 12345u }, { 3, 4294967295u } };
        bool first = true;
        for (auto &p : in) {
            auto [a, b] = mi::sample_tea_32<uint32_t>(p[0], p[1]);
            printf("%s[%u,%u,%u,%u]", first ? "" : ",", p[0], p[1], a, b);
            first = false;
        }
    }
    printf("],\n\"tea_float32_1_1_4\": ");
    pf(mi::sample_tea_float32<uint32_t>(1, 1, 4));
    // ---- PCG32
    printf(",\n\"pcg32\": [");
    {
        uint64_t seeds[][2] = { { 42, 54 }, { 0x5df5f2bfull, 0x54ce08baull }, { 0x5ac7cb65ull, 0xac39bb45ull },
                                { PCG32_DEFAULT_STATE, PCG32_DEFAULT_STREAM }, { 0xffffffffull, 0xffffffffull } };
        bool first = true;
        for (auto &s : seeds) {
            mi::PCG32<uint32_t> r(1, s[0], s[1]), r2(1, s[0], s[1]);
            printf("%s{\"initstate\":%llu,\"initseq\":%llu,\"state\":%llu,\"inc\":%llu,\"u32\":[", first ? "" : ",",
                   (unsigned long long) s[0], (unsigned long long) s[1], (unsigned long long) r.state,
                   (unsigned long long) r.inc);
            for (int i = 0; i < 8; ++i)
                printf("%s%u", i ? "," : "", r.next_uint32());
            printf("],\"f32\":[");
            for (int i = 0; i < 8; ++i) {
                printf("%s", i ? "," : "");
                pf(r2.next_float32());
            }
            printf("]}");
            first = false;
        }
    }
    // ---- Kensler
    printf("],\n\"kensler\": [");
    {
        bool first = true;
        uint32_t cases[][3] = { { 3, 512, 12345 }, { 0, 512, 0 }, { 511, 512, 0xdeadbeef }, { 7, 100, 99 }, { 99, 100, 1 },
                                { 5, 1, 77 }, { 1, 2, 3 }, { 255, 300, 4242 }, { 12, 2048, 0x9e3779b9 } };
        for (auto &c : cases) {
            printf("%s[%u,%u,%u,%u]", first ? "" : ",", c[0], c[1], c[2], mi::permute_kensler<uint32_t>(c[0], c[1], c[2]));
            first = false;
        }
        // full permutations (bijection check material)
        for (uint32_t n : { 7u, 16u, 100u })
            for (uint32_t i = 0; i < n; ++i)
                printf(",[%u,%u,%u,%u]", i, n, 1000u + n, mi::permute_kensler<uint32_t>(i, n, 1000u + n));
    }
    // ---- sincos / waveforms over a grid
    printf("],\n\"sincos\": [");
    {
        bool first = true;
        for (int i = -40; i <= 400; ++i) {
            float x = 0.1234f * (float) i + 0.001f * (float) (i * i % 7);
            auto [s, c] = dr::sincos(x);
            printf("%s[", first ? "" : ",");
            pf(x), printf(","), pf(s), printf(","), pf(c), printf("]");
            first = false;
        }
    }
    printf("],\n\"waveform_lowpass\": [");
    {
        bool first = true;
        for (int type = 0; type < 4; ++type)
            for (int i = -30; i <= 300; ++i) {
                float t = 0.0731f * (float) i + 0.0003f * (float) type;
                float v = mi::eval_modulation_function_value_low_pass<float>(t, (mi::EWaveformType) type);
                printf("%s[%d,", first ? "" : ",", type);
                pf(t), printf(","), pf(v), printf("]");
                first = false;
            }
    }
    printf("],\n\"waveform_full\": [");
    {
        bool first = true;
        for (int type = 0; type < 4; ++type)
            for (int i = -30; i <= 300; ++i) {
                float t = 0.0731f * (float) i + 0.0003f * (float) type;
                float v = mi::eval_modulation_function_value<float>(t, (mi::EWaveformType) type);
                printf("%s[%d,", first ? "" : ",", type);
                pf(t), printf(","), pf(v), printf("]");
                first = false;
            }
    }
    // ---- warps
    printf("],\n\"cosine_hemisphere\": [");
    {
        bool first = true;
        mi::PCG32<uint32_t> r(1, 7, 9);
        for (int i = 0; i < 64; ++i) {
            float u = r.next_float32(), v = r.next_float32();
            if (i == 0) u = v = 0.5f;
            if (i == 1) u = 0.f, v = 0.f;
            if (i == 2) u = 0.75f, v = 0.25f;
            auto w = mi::warp::square_to_cosine_hemisphere(mi::Point<float, 2>(u, v));
            printf("%s[", first ? "" : ",");
            pf(u), printf(","), pf(v), printf(","), pf(w.x()), printf(","), pf(w.y()), printf(","), pf(w.z()), printf("]");
            first = false;
        }
    }
    printf("],\n\"uniform_triangle\": [");
    {
        bool first = true;
        mi::PCG32<uint32_t> r(1, 11, 13);
        for (int i = 0; i < 16; ++i) {
            float u = r.next_float32(), v = r.next_float32();
            auto w = mi::warp::square_to_uniform_triangle(mi::Point<float, 2>(u, v));
            printf("%s[", first ? "" : ",");
            pf(u), printf(","), pf(v), printf(","), pf(w.x()), printf(","), pf(w.y()), printf("]");
            first = false;
        }
    }
    printf("],\n\"coordinate_system\": [");
    {
        bool first = true;
        mi::PCG32<uint32_t> r(1, 3, 5);
        for (int i = 0; i < 32; ++i) {
            mi::Vector<float, 3> n(r.next_float32() * 2 - 1, r.next_float32() * 2 - 1, r.next_float32() * 2 - 1);
            n = dr::normalize(n);
            if (i == 0) n = mi::Vector<float, 3>(0, 0, 1);
            if (i == 1) n = mi::Vector<float, 3>(0, 0, -1);
            if (i == 2) n = mi::Vector<float, 3>(0, 1, 0);
            auto [s, t] = mi::coordinate_system(n);
            printf("%s[", first ? "" : ",");
            pf(n.x()), printf(","), pf(n.y()), printf(","), pf(n.z()), printf(",");
            pf(s.x()), printf(","), pf(s.y()), printf(","), pf(s.z()), printf(",");
            pf(t.x()), printf(","), pf(t.y()), printf(","), pf(t.z()), printf("]");
            first = false;
        }
    }
    // ---- perspective projection: sample_to_camera matrices (row-major)
    printf("],\n\"perspective\": [");
    {
        struct C {
            int fw, fh, cw, ch, ox, oy;
            float fov, nearc, farc;
        } cases[] = { { 256, 256, 256, 256, 0, 0, 19.5f, 1e-2f, 1e4f },   { 512, 512, 512, 512, 0, 0, 39.3077f, 1e-2f, 1e4f },
                      { 640, 480, 640, 480, 0, 0, 45.f, 1e-2f, 1e4f },    { 1024, 1024, 256, 128, 64, 32, 30.f, 0.1f, 100.f },
                      { 2048, 2048, 2048, 2048, 0, 0, 19.5f, 1e-2f, 1e4f } };
        bool first = true;
        for (auto &c : cases) {
            auto c2s = mi::perspective_projection<float>(mi::Vector<int, 2>(c.fw, c.fh), mi::Vector<int, 2>(c.cw, c.ch),
                                                         mi::Vector<int, 2>(c.ox, c.oy), c.fov, c.nearc, c.farc);
            auto s2c = c2s.inverse();
            printf("%s{\"film\":[%d,%d],\"crop\":[%d,%d],\"offset\":[%d,%d],\"fov\":", first ? "" : ",", c.fw, c.fh, c.cw,
                   c.ch, c.ox, c.oy);
            pf(c.fov), printf(",\"near\":"), pf(c.nearc), printf(",\"far\":"), pf(c.farc);
            printf(",\"sample_to_camera\":[");
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) {
                    printf("%s", (i || j) ? "," : "");
                    pf(s2c.matrix(i, j));
                }
            printf("],\"near_p\":[");
            // a few transformed points with projective divide (what sample_ray does)
            float uv[][2] = { { 0.5f, 0.5f }, { 0.f, 0.f }, { 1.f, 1.f }, { 0.251f, 0.77f } };
            for (int k = 0; k < 4; ++k) {
                auto p = s2c * mi::Point<float, 3>(uv[k][0], uv[k][1], 0.f);
                printf("%s[", k ? "," : "");
                pf(uv[k][0]), printf(","), pf(uv[k][1]), printf(","), pf(p.x()), printf(","), pf(p.y()), printf(","), pf(p.z());
                printf("]");
            }
            printf("]}");
            first = false;
        }
    }
    printf("]\n}\n");
    return 0;
}
