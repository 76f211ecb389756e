// PLY / OBJ readers of the C++ host (the parts of src/shapes/ply.cpp, src/shapes/obj.cpp and
// Mesh::recompute_vertex_normals, src/render/mesh.cpp:283-345, that the hot path observes).
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <tuple>

#include <zlib.h>

#include "dtof_host.hpp"

namespace dtof_host {

namespace {

struct PlyProp {
    bool is_list = false;
    std::string type, count_type, name;
};
struct PlyElement {
    std::string name;
    size_t count = 0;
    std::vector<PlyProp> props;
};

size_t ply_size(const std::string &t) {
    if (t == "char" || t == "int8" || t == "uchar" || t == "uint8") return 1;
    if (t == "short" || t == "int16" || t == "ushort" || t == "uint16") return 2;
    if (t == "int" || t == "int32" || t == "uint" || t == "uint32" || t == "float" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    throw Error("PLY: unknown property type '" + t + "'");
}

double ply_read(std::istream &f, const std::string &t, bool big) {
    unsigned char b[8];
    size_t n = ply_size(t);
    f.read((char *) b, (std::streamsize) n);
    if (!f)
        throw Error("PLY: truncated file");
    if (big)
        for (size_t i = 0; i < n / 2; ++i)
            std::swap(b[i], b[n - 1 - i]);
    if (t == "char" || t == "int8") { int8_t v; memcpy(&v, b, 1); return v; }
    if (t == "uchar" || t == "uint8") { uint8_t v; memcpy(&v, b, 1); return v; }
    if (t == "short" || t == "int16") { int16_t v; memcpy(&v, b, 2); return v; }
    if (t == "ushort" || t == "uint16") { uint16_t v; memcpy(&v, b, 2); return v; }
    if (t == "int" || t == "int32") { int32_t v; memcpy(&v, b, 4); return v; }
    if (t == "uint" || t == "uint32") { uint32_t v; memcpy(&v, b, 4); return v; }
    if (t == "float" || t == "float32") { float v; memcpy(&v, b, 4); return v; }
    double v;
    memcpy(&v, b, 8);
    return v;
}

void fan(const std::vector<long long> &idx, std::vector<uint32_t> &faces) {
    for (size_t i = 1; i + 1 < idx.size(); ++i) {
        faces.push_back((uint32_t) idx[0]);
        faces.push_back((uint32_t) idx[i]);
        faces.push_back((uint32_t) idx[i + 1]);
    }
}

void load_ply(const std::string &path, std::vector<float> &pos, std::vector<uint32_t> &faces, std::vector<float> &nrm,
              std::vector<float> &uv) {
    std::ifstream f(path, std::ios::binary);
    if (!f)
        throw Error("\"" + path + "\": file does not exist!");
    std::string line;
    std::getline(f, line);
    while (!line.empty() && (line.back() == '\r' || line.back() == ' '))
        line.pop_back();
    if (line != "ply")
        throw Error(path + ": not a PLY file");
    std::string fmt;
    std::vector<PlyElement> elements;
    for (;;) {
        if (!std::getline(f, line))
            throw Error(path + ": truncated PLY header");
        std::istringstream ss(line);
        std::vector<std::string> tok;
        for (std::string t; ss >> t;)
            tok.push_back(t);
        if (tok.empty() || tok[0] == "comment")
            continue;
        if (tok[0] == "format")
            fmt = tok.at(1);
        else if (tok[0] == "element")
            elements.push_back(PlyElement{ tok.at(1), (size_t) std::stoull(tok.at(2)), {} });
        else if (tok[0] == "property") {
            PlyProp p;
            if (tok.at(1) == "list") {
                p.is_list = true;
                p.count_type = tok.at(2);
                p.type = tok.at(3);
                p.name = tok.at(4);
            } else {
                p.type = tok.at(1);
                p.name = tok.at(2);
            }
            if (elements.empty())
                throw Error(path + ": property before element");
            elements.back().props.push_back(p);
        } else if (tok[0] == "end_header")
            break;
    }
    const bool ascii = fmt == "ascii", big = fmt == "binary_big_endian";
    if (!ascii && !big && fmt != "binary_little_endian")
        throw Error(path + ": unknown PLY format '" + fmt + "'");
    std::map<std::string, std::vector<double>> verts;
    for (auto &el : elements) {
        for (size_t r = 0; r < el.count; ++r) {
            std::istringstream ls;
            if (ascii) {
                if (!std::getline(f, line))
                    throw Error(path + ": truncated PLY body");
                ls.str(line);
            }
            for (auto &p : el.props) {
                if (p.is_list) {
                    long long k;
                    if (ascii) {
                        double kd;
                        ls >> kd;
                        k = (long long) kd;
                    } else {
                        k = (long long) ply_read(f, p.count_type, big);
                    }
                    std::vector<long long> idx((size_t) k);
                    for (auto &v : idx) {
                        if (ascii) {
                            double d;
                            ls >> d;
                            v = (long long) d;
                        } else {
                            v = (long long) ply_read(f, p.type, big);
                        }
                    }
                    if (el.name == "face" && (p.name == "vertex_indices" || p.name == "vertex_index" || ascii))
                        fan(idx, faces);
                } else {
                    double d;
                    if (ascii)
                        ls >> d;
                    else
                        d = ply_read(f, p.type, big);
                    if (el.name == "vertex")
                        verts[p.name].push_back(d);
                }
            }
        }
    }
    if (!verts.count("x") || !verts.count("y") || !verts.count("z"))
        throw Error(path + ": PLY has no vertex positions");
    size_t nv = verts["x"].size();
    pos.resize(3 * nv);
    for (size_t i = 0; i < nv; ++i) {
        pos[3 * i] = (float) verts["x"][i];
        pos[3 * i + 1] = (float) verts["y"][i];
        pos[3 * i + 2] = (float) verts["z"][i];
    }
    nrm.clear();
    if (verts.count("nx")) {
        nrm.resize(3 * nv);
        for (size_t i = 0; i < nv; ++i) {
            nrm[3 * i] = (float) verts["nx"][i];
            nrm[3 * i + 1] = (float) verts["ny"][i];
            nrm[3 * i + 2] = (float) verts["nz"][i];
        }
    }
    uv.clear();
    const char *pairs[3][2] = { { "u", "v" }, { "s", "t" }, { "texture_u", "texture_v" } };
    for (auto &pr : pairs)
        if (verts.count(pr[0])) {
            uv.resize(2 * nv);
            for (size_t i = 0; i < nv; ++i) {
                uv[2 * i] = (float) verts[pr[0]][i];
                uv[2 * i + 1] = (float) verts[pr[1]][i];
            }
            break;
        }
}

void load_obj(const std::string &path, std::vector<float> &pos, std::vector<uint32_t> &faces, std::vector<float> &nrm,
              std::vector<float> &uv, bool flip_tex_coords) {
    std::ifstream f(path);
    if (!f)
        throw Error("\"" + path + "\": file does not exist!");
    std::vector<std::array<double, 3>> v, vn;
    std::vector<std::array<double, 2>> vt;
    std::map<std::tuple<long long, long long, long long>, uint32_t> keys;
    bool all_uv = true, all_n = true;
    std::string line;
    pos.clear(), faces.clear(), nrm.clear(), uv.clear();
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::string cmd;
        if (!(ss >> cmd))
            continue;
        if (cmd == "v") {
            std::array<double, 3> a{};
            ss >> a[0] >> a[1] >> a[2];
            v.push_back(a);
        } else if (cmd == "vt") {
            std::array<double, 2> a{};
            ss >> a[0] >> a[1];
            vt.push_back(a);
        } else if (cmd == "vn") {
            std::array<double, 3> a{};
            ss >> a[0] >> a[1] >> a[2];
            vn.push_back(a);
        } else if (cmd == "f") {
            std::vector<long long> idx;
            for (std::string t; ss >> t;) {
                long long k[3] = { 0, 0, 0 };
                size_t b = 0;
                for (int c = 0; c < 3 && b <= t.size(); ++c) {
                    size_t e = t.find('/', b);
                    std::string part = t.substr(b, e == std::string::npos ? std::string::npos : e - b);
                    if (!part.empty())
                        k[c] = std::stoll(part);
                    if (e == std::string::npos)
                        break;
                    b = e + 1;
                }
                if (k[0] < 0) k[0] += (long long) v.size() + 1;
                if (k[1] < 0) k[1] += (long long) vt.size() + 1;
                if (k[2] < 0) k[2] += (long long) vn.size() + 1;
                auto key = std::make_tuple(k[0], k[1], k[2]);
                auto it = keys.find(key);
                if (it == keys.end()) {
                    uint32_t id = (uint32_t) (pos.size() / 3);
                    it = keys.emplace(key, id).first;
                    if (k[0] < 1 || k[0] > (long long) v.size())
                        throw Error(path + ": OBJ vertex index out of range");
                    for (int c = 0; c < 3; ++c)
                        pos.push_back((float) v[(size_t) k[0] - 1][c]);
                    if (k[1]) {
                        uv.push_back((float) vt.at((size_t) k[1] - 1)[0]);
                        const float tv = (float) vt.at((size_t) k[1] - 1)[1];
                        uv.push_back(flip_tex_coords ? 1.f - tv : tv);   // obj.cpp:151,266-267 (default true)
                    } else {
                        uv.push_back(0.f), uv.push_back(0.f);
                        all_uv = false;
                    }
                    if (k[2]) {
                        for (int c = 0; c < 3; ++c)
                            nrm.push_back((float) vn.at((size_t) k[2] - 1)[c]);
                    } else {
                        nrm.insert(nrm.end(), 3, 0.f);
                        all_n = false;
                    }
                }
                idx.push_back(it->second);
            }
            fan(idx, faces);
        }
    }
    if (!all_uv || pos.empty())
        uv.clear();
    if (!all_n || pos.empty())
        nrm.clear();
}

} // namespace

// Angle-weighted smooth normals (Thuermer & Wuethrich), as Mesh::recompute_vertex_normals; double accumulation
// in the order k = 0, 1, 2 over all faces (the Python host's np.add.at order).
void vertex_normals(const std::vector<float> &pos, const std::vector<uint32_t> &faces, std::vector<float> &out) {
    size_t nv = pos.size() / 3, nf = faces.size() / 3;
    std::vector<double> n(3 * nv, 0.0), fn(3 * nf);
    auto P = [&](uint32_t i, int c) { return (double) pos[3 * (size_t) i + c]; };
    for (size_t f = 0; f < nf; ++f) {
        uint32_t a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
        double e0[3], e1[3];
        for (int k = 0; k < 3; ++k) {
            e0[k] = P(b, k) - P(a, k);
            e1[k] = P(c, k) - P(a, k);
        }
        double cr[3] = { e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0] };
        double l = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
        for (int k = 0; k < 3; ++k)
            fn[3 * f + k] = l > 0 ? cr[k] / std::max(l, 1e-300) : 0.0;
    }
    for (int k = 0; k < 3; ++k)
        for (size_t f = 0; f < nf; ++f) {
            uint32_t a = faces[3 * f + k], b = faces[3 * f + (k + 1) % 3], c = faces[3 * f + (k + 2) % 3];
            double d0[3], d1[3];
            for (int j = 0; j < 3; ++j) {
                d0[j] = P(b, j) - P(a, j);
                d1[j] = P(c, j) - P(a, j);
            }
            double l0 = std::max(std::sqrt(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]), 1e-300);
            double l1 = std::max(std::sqrt(d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2]), 1e-300);
            for (int j = 0; j < 3; ++j) {
                d0[j] /= l0;
                d1[j] /= l1;
            }
            double dp = d0[0] * d1[0] + d0[1] * d1[1] + d0[2] * d1[2];
            double ang = std::acos(std::min(std::max(dp, -1.0), 1.0));
            for (int j = 0; j < 3; ++j)
                n[3 * (size_t) a + j] += fn[3 * f + j] * ang;
        }
    out.resize(3 * nv);
    for (size_t i = 0; i < nv; ++i) {
        double l = std::sqrt(n[3 * i] * n[3 * i] + n[3 * i + 1] * n[3 * i + 1] + n[3 * i + 2] * n[3 * i + 2]);
        for (int j = 0; j < 3; ++j)
            out[3 * i + j] = (float) (l > 0 ? n[3 * i + j] / std::max(l, 1e-300) : (j == 0 ? 1.0 : 0.0));
    }
}

namespace {

} // namespace

// One sub-mesh of a Mitsuba `.serialized` file (src/shapes/serialized.cpp:229-392): header 0x041C, version 3 / 4, one
// zlib stream per sub-mesh, end-of-file offset dictionary; float or double payload narrowed to float, optional
// normals / texture coordinates, vertex colours skipped, uint32 indices. Error texts are the reference's.
void load_serialized_file(const std::string &path, int shape_index, bool face_normals, std::vector<float> &pos,
                          std::vector<uint32_t> &faces, std::vector<float> &normals, std::vector<float> &uvs, bool compute_missing) {
    auto fail = [&](const std::string &descr) -> void {
        throw Error("Error while loading serialized file \"" + path + "\": " + descr + "!");
    };
    if (shape_index < 0)
        fail("shape index must be nonnegative");
    std::ifstream f(path, std::ios::binary);
    if (!f)
        fail("file not found");
    std::vector<unsigned char> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto rd16 = [&](size_t o) { return (uint32_t) data[o] | ((uint32_t) data[o + 1] << 8); };
    auto rd32 = [&](size_t o) { return rd16(o) | (rd16(o + 2) << 16); };
    auto rd64 = [&](size_t o) { return (uint64_t) rd32(o) | ((uint64_t) rd32(o + 4) << 32); };
    if (data.size() < 4 || rd16(0) != 0x041Cu)
        fail("encountered an invalid file format");
    const uint32_t version = rd16(2);
    if (version != 3 && version != 4)
        fail("encountered an incompatible file version");
    size_t offset = 0;
    if (shape_index != 0) {
        if (data.size() < 8)
            fail("encountered an invalid file format");
        const uint32_t count = rd32(data.size() - 4);
        if ((uint32_t) shape_index >= count)
            fail("Unable to unserialize mesh, shape index is out of range! (requested " + std::to_string(shape_index) +
                 " out of 0.." + std::to_string((long long) count - 1) + ")");
        offset = version == 4 ? (size_t) rd64(data.size() - 8 * (size_t) (count - shape_index) - 4)
                              : (size_t) rd32(data.size() - 4 * (size_t) (count - shape_index + 1));
        if (offset + 4 > data.size())
            fail("encountered an invalid file format");
    }
    // inflate the sub-mesh's stream (it ends where the next header starts; zlib stops at the stream end by itself)
    std::vector<unsigned char> raw;
    {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit(&zs) != Z_OK)
            fail("zlib initialisation failed");
        zs.next_in = data.data() + offset + 4;
        zs.avail_in = (uInt) (data.size() - offset - 4);
        unsigned char buf[1 << 16];
        int rc = Z_OK;
        while (rc != Z_STREAM_END) {
            zs.next_out = buf;
            zs.avail_out = sizeof(buf);
            rc = inflate(&zs, Z_NO_FLUSH);
            if (rc != Z_OK && rc != Z_STREAM_END) {
                inflateEnd(&zs);
                fail("the compressed stream is corrupt");
            }
            raw.insert(raw.end(), buf, buf + (sizeof(buf) - zs.avail_out));
            if (rc == Z_OK && zs.avail_in == 0 && zs.avail_out != 0)
                break;
        }
        inflateEnd(&zs);
    }
    size_t at = 0;
    auto take = [&](size_t n) -> const unsigned char * {
        if (at + n > raw.size())
            fail("unexpected end of the compressed stream");
        const unsigned char *p = raw.data() + at;
        at += n;
        return p;
    };
    uint32_t flags;
    memcpy(&flags, take(4), 4);
    if (version == 4) {
        while (*take(1) != 0) {
        }
    }
    uint64_t nv, nf;
    memcpy(&nv, take(8), 8);
    memcpy(&nf, take(8), 8);
    const bool dp = (flags & 0x2000u) != 0;
    auto array = [&](size_t dim, std::vector<float> *dst) {
        const size_t n = (size_t) nv * dim;
        const unsigned char *p = take(n * (dp ? 8 : 4));
        if (!dst)
            return;
        dst->resize(n);
        if (dp) {
            for (size_t i = 0; i < n; ++i) {
                double v;
                memcpy(&v, p + 8 * i, 8);
                (*dst)[i] = (float) v;
            }
        } else {
            memcpy(dst->data(), p, 4 * n);
        }
    };
    normals.clear();
    uvs.clear();
    array(3, &pos);
    if (flags & 0x1u)
        array(3, face_normals ? nullptr : &normals);
    if (flags & 0x2u)
        array(2, &uvs);
    if (flags & 0x8u)
        array(3, nullptr);
    faces.resize((size_t) nf * 3);
    memcpy(faces.data(), take((size_t) nf * 12), (size_t) nf * 12);
    for (uint32_t i : faces)
        if ((size_t) i >= pos.size() / 3)
            throw Error(path + ": face references a vertex out of range");
    if (normals.empty() && !face_normals && compute_missing)
        vertex_normals(pos, faces, normals);
}

void load_mesh_file(const std::string &path, bool face_normals, std::vector<float> &pos, std::vector<uint32_t> &faces,
                    std::vector<float> &normals, std::vector<float> &uvs, bool compute_missing, bool flip_tex_coords) {
    std::string lower = path;
    for (char &c : lower)
        c = (char) std::tolower((unsigned char) c);
    if (lower.size() >= 4 && lower.compare(lower.size() - 4, 4, ".ply") == 0)
        load_ply(path, pos, faces, normals, uvs);
    else
        load_obj(path, pos, faces, normals, uvs, flip_tex_coords);
    for (uint32_t i : faces)
        if ((size_t) i >= pos.size() / 3)
            throw Error(path + ": face references a vertex out of range");
    if (normals.empty() && !face_normals && compute_missing)
        vertex_normals(pos, faces, normals);
}

} // namespace dtof_host
