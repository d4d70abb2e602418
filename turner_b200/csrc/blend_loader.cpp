// Minimal .blend reader for the CLI: replaces Assimp::Importer::ReadFile +
// triangles_from_scene + the camera/light set-up of main.cpp:25-82,96-136 for the
// kind of file the reference ships (uncompressed BLENDER-v27x, 8-byte pointers,
// little endian, MPoly/MLoop meshes, no parenting).
//
// It parses the file's own SDNA ("DNA1" block) to find field offsets, then walks
// Scene.base -> Object -> Mesh/Camera/Lamp. Conventions restated from assimp@a5a5343's
// BlenderLoader (not vendored): one face per MPoly (triangle or quad), per-corner
// vertices, vertex normals MVert.no/32767, quads split (0,1,2),(0,2,3), one mesh per
// used material slot in slot order, diffuse = Material.r/g/b with alpha 1, camera FOV =
// atan2(sensor_x, 2*lens), mAspect left 0, lamp colour = rgb * energy. Triangles come out
// in world space exactly as main.cpp:53-59 transforms them (T*v, 3x3(T)*n, fp32).
#include "../../include/turner_b200.h"
#include "host_util.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <array>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Block {
    char code[5];
    int32_t size;
    uint64_t old_ptr;
    int32_t sdna;
    int32_t count;
    size_t offset;
};

struct Field {
    size_t offset;
    std::string type;
};

struct BlendFile {
    std::vector<unsigned char> data;
    std::vector<Block> blocks;
    std::unordered_map<uint64_t, size_t> by_ptr;
    std::vector<std::string> names, types;
    std::vector<int16_t> tlen;
    std::map<std::string, std::map<std::string, Field>> structs;
    std::map<std::string, size_t> struct_size;
    std::string error;

    template <typename T> T rd(size_t off) const {
        T v;
        std::memcpy(&v, data.data() + off, sizeof(T));
        return v;
    }

    bool load(const char* path) {
        FILE* f = std::fopen(path, "rb");
        if (!f) {
            error = std::string("cannot open ") + path;
            return false;
        }
        std::fseek(f, 0, SEEK_END);
        long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        data.resize(n > 0 ? static_cast<size_t>(n) : 0);
        size_t got = data.empty() ? 0 : std::fread(data.data(), 1, data.size(), f);
        std::fclose(f);
        if (got != data.size() || data.size() < 12 || std::memcmp(data.data(), "BLENDER", 7) != 0) {
            error = "not an uncompressed .blend file";
            return false;
        }
        if (data[7] != '-' || data[8] != 'v') {
            error = "only 64-bit little-endian .blend files are supported";
            return false;
        }
        size_t off = 12;
        while (off + 24 <= data.size()) {
            Block b;
            std::memcpy(b.code, data.data() + off, 4);
            b.code[4] = 0;
            b.size = rd<int32_t>(off + 4);
            b.old_ptr = rd<uint64_t>(off + 8);
            b.sdna = rd<int32_t>(off + 16);
            b.count = rd<int32_t>(off + 20);
            b.offset = off + 24;
            if (b.size < 0 || b.offset + static_cast<size_t>(b.size) > data.size()) break;
            blocks.push_back(b);
            if (b.old_ptr) by_ptr[b.old_ptr] = blocks.size() - 1;
            if (std::memcmp(b.code, "ENDB", 4) == 0) break;
            off = b.offset + static_cast<size_t>(b.size);
        }
        return parse_dna();
    }

    bool parse_dna() {
        const Block* dna = nullptr;
        for (auto& b : blocks)
            if (std::memcmp(b.code, "DNA1", 4) == 0) dna = &b;
        if (!dna) {
            error = "no DNA1 block";
            return false;
        }
        size_t off = dna->offset;
        auto expect = [&](const char* tag) {
            off = (off + 3) & ~size_t(3);
            bool ok = std::memcmp(data.data() + off, tag, 4) == 0;
            off += 4;
            return ok;
        };
        if (std::memcmp(data.data() + off, "SDNA", 4) != 0) return false;
        off += 4;
        if (!expect("NAME")) return false;
        int32_t n = rd<int32_t>(off);
        off += 4;
        for (int i = 0; i < n; ++i) {
            std::string s(reinterpret_cast<const char*>(data.data() + off));
            off += s.size() + 1;
            names.push_back(s);
        }
        if (!expect("TYPE")) return false;
        n = rd<int32_t>(off);
        off += 4;
        for (int i = 0; i < n; ++i) {
            std::string s(reinterpret_cast<const char*>(data.data() + off));
            off += s.size() + 1;
            types.push_back(s);
        }
        if (!expect("TLEN")) return false;
        for (size_t i = 0; i < types.size(); ++i) tlen.push_back(rd<int16_t>(off + 2 * i));
        off += 2 * types.size();
        if (!expect("STRC")) return false;
        n = rd<int32_t>(off);
        off += 4;
        sdna_struct_type.clear();
        for (int i = 0; i < n; ++i) {
            int16_t t = rd<int16_t>(off), nf = rd<int16_t>(off + 2);
            off += 4;
            std::map<std::string, Field> fields;
            size_t fo = 0;
            for (int k = 0; k < nf; ++k) {
                int16_t ft = rd<int16_t>(off), fn = rd<int16_t>(off + 2);
                off += 4;
                const std::string& raw = names[fn];
                bool is_ptr = raw[0] == '*' || raw.compare(0, 2, "(*") == 0;
                size_t b = raw.find_first_not_of("*(");
                size_t e = raw.find_first_of("[)", b);
                std::string base = raw.substr(b, e == std::string::npos ? std::string::npos : e - b);
                size_t mult = 1;
                for (size_t p = raw.find('['); p != std::string::npos; p = raw.find('[', p + 1))
                    mult *= static_cast<size_t>(std::atoi(raw.c_str() + p + 1));
                fields[base] = Field{fo, types[ft]};
                fo += (is_ptr ? 8 : static_cast<size_t>(tlen[ft])) * mult;
            }
            structs[types[t]] = fields;
            struct_size[types[t]] = static_cast<size_t>(tlen[t]);
            sdna_struct_type.push_back(types[t]);
        }
        return true;
    }
    std::vector<std::string> sdna_struct_type;

    const Block* at(uint64_t ptr) const {
        auto it = by_ptr.find(ptr);
        return it == by_ptr.end() ? nullptr : &blocks[it->second];
    }
    bool has(const std::string& st, const std::string& f) const {
        auto it = structs.find(st);
        return it != structs.end() && it->second.count(f);
    }
    size_t foff(const std::string& st, const std::string& f) const { return structs.at(st).at(f).offset; }
    template <typename T> T get(const Block* b, const std::string& st, const std::string& f, size_t elem = 0, size_t sub = 0) const {
        return rd<T>(b->offset + elem * struct_size.at(st) + foff(st, f) + sub * sizeof(T));
    }
};

} // namespace

extern "C" {

int32_t trn_load_blend(const char* path, trn_loaded_scene* out) {
    if (!path || !out) return trn::fail(TRN_ERR_INVALID, "null argument");
    std::memset(out, 0, sizeof *out);
    BlendFile bf;
    if (!bf.load(path)) return trn::fail(TRN_ERR_IO, bf.error.empty() ? "malformed .blend (SDNA)" : bf.error);
    for (const char* st : {"Scene", "Base", "Object", "Mesh", "MVert", "MPoly", "MLoop"})
        if (!bf.structs.count(st)) return trn::fail(TRN_ERR_IO, std::string("unsupported .blend: no struct ") + st);

    const Block* scene = nullptr;
    for (auto& b : bf.blocks)
        if (std::memcmp(b.code, "SC\0\0", 4) == 0) {
            scene = &b;
            break;
        }
    if (!scene) return trn::fail(TRN_ERR_IO, "no Scene block");

    std::vector<float> verts, norms, cols, mirs, refls;
    int num_cameras = 0;
    uint64_t base_ptr = bf.rd<uint64_t>(scene->offset + bf.foff("Scene", "base")); // ListBase.first
    while (base_ptr) {
        const Block* base = bf.at(base_ptr);
        if (!base) break;
        base_ptr = bf.get<uint64_t>(base, "Base", "next");
        const Block* ob = bf.at(bf.get<uint64_t>(base, "Base", "object"));
        if (!ob) continue;
        if (bf.get<uint64_t>(ob, "Object", "parent") != 0) return trn::fail(TRN_ERR_IO, "parented objects are not supported");
        const int16_t otype = bf.get<int16_t>(ob, "Object", "type");
        float T[4][4]; // row-major assimp a1..d4 = transpose of Blender's obmat
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) T[r][c] = bf.get<float>(ob, "Object", "obmat", 0, static_cast<size_t>(c) * 4 + r);
        const Block* data = bf.at(bf.get<uint64_t>(ob, "Object", "data"));
        if (!data) continue;
        if (otype == 11) { // camera
            const float lens = bf.get<float>(data, "Camera", "lens");
            const float sensor = bf.has("Camera", "sensor_x") ? bf.get<float>(data, "Camera", "sensor_x") : 32.f;
            if (num_cameras++ == 0) {
                out->has_camera = 1;
                for (int r = 0; r < 4; ++r)
                    for (int c = 0; c < 4; ++c) out->cam_trafo4x4[4 * r + c] = T[r][c];
                // evaluated in double and rounded once: reproduces README.md:22-36's ray count exactly
                out->cam_hfov = static_cast<float>(std::atan2(static_cast<double>(sensor), static_cast<double>(2.f * lens)));
                out->cam_aspect = 0.f;
            }
        } else if (otype == 10) { // lamp
            if (out->num_lights == 0) {
                const float e = bf.get<float>(data, "Lamp", "energy");
                out->light.pos[0] = T[0][3];
                out->light.pos[1] = T[1][3];
                out->light.pos[2] = T[2][3];
                out->light.rgba[0] = bf.get<float>(data, "Lamp", "r") * e;
                out->light.rgba[1] = bf.get<float>(data, "Lamp", "g") * e;
                out->light.rgba[2] = bf.get<float>(data, "Lamp", "b") * e;
                out->light.rgba[3] = 1.f;
            }
            out->num_lights += 1;
        } else if (otype == 1) { // mesh
            const int32_t totpoly = bf.get<int32_t>(data, "Mesh", "totpoly");
            const int16_t totcol = bf.get<int16_t>(data, "Mesh", "totcol");
            const Block* mvert = bf.at(bf.get<uint64_t>(data, "Mesh", "mvert"));
            const Block* mpoly = bf.at(bf.get<uint64_t>(data, "Mesh", "mpoly"));
            const Block* mloop = bf.at(bf.get<uint64_t>(data, "Mesh", "mloop"));
            if (!mvert || !mpoly || !mloop) continue;
            struct Mat {
                std::array<float, 4> dif, mir;
                float refl;
            };
            const Mat dflt = {{0.6f, 0.6f, 0.6f, 1.f}, {0.f, 0.f, 0.f, 0.f}, 0.f}; // assimp's default material
            std::vector<Mat> mats;
            const Block* matarr = bf.at(bf.get<uint64_t>(data, "Mesh", "mat"));
            if (matarr)
                for (int i = 0; i < totcol; ++i) {
                    const Block* ma = bf.at(bf.rd<uint64_t>(matarr->offset + 8 * static_cast<size_t>(i)));
                    if (!ma) {
                        mats.push_back(dflt);
                        continue;
                    }
                    Mat m;
                    const float r = bf.get<float>(ma, "Material", "r"), g = bf.get<float>(ma, "Material", "g"), b = bf.get<float>(ma, "Material", "b");
                    // an all-zero diffuse colour is not exported -> Get() leaves the default aiColor4D (0,0,0,0)
                    m.dif = (r || g || b) ? std::array<float, 4>{r, g, b, 1.f} : std::array<float, 4>{0.f, 0.f, 0.f, 0.f};
                    m.mir = {bf.get<float>(ma, "Material", "mirr"), bf.get<float>(ma, "Material", "mirg"), bf.get<float>(ma, "Material", "mirb"), 1.f};
                    // AI_MATKEY_REFLECTIVITY is only exported when ray mirroring is on (MA_RAYMIRROR = 0x40000)
                    m.refl = (bf.get<int32_t>(ma, "Material", "mode") & 0x40000) ? bf.get<float>(ma, "Material", "ray_mirror") : 0.f;
                    mats.push_back(m);
                }
            if (mats.empty()) mats.push_back(dflt);
            int max_slot = 0;
            for (int i = 0; i < totpoly; ++i) max_slot = std::max<int>(max_slot, bf.get<int16_t>(mpoly, "MPoly", "mat_nr", i));
            for (int slot = 0; slot <= max_slot; ++slot) {
                const Mat& col = mats[static_cast<size_t>(slot) < mats.size() ? slot : 0];
                for (int i = 0; i < totpoly; ++i) {
                    if (bf.get<int16_t>(mpoly, "MPoly", "mat_nr", i) != slot) continue;
                    const int32_t ls = bf.get<int32_t>(mpoly, "MPoly", "loopstart", i);
                    const int32_t tl = bf.get<int32_t>(mpoly, "MPoly", "totloop", i);
                    if (tl != 3 && tl != 4) return trn::fail(TRN_ERR_IO, "n-gons are not supported");
                    int32_t vi[4];
                    for (int k = 0; k < tl; ++k) vi[k] = bf.get<int32_t>(mloop, "MLoop", "v", static_cast<size_t>(ls + k));
                    const int order[2][3] = {{0, 1, 2}, {0, 2, 3}};
                    for (int tri = 0; tri < (tl == 4 ? 2 : 1); ++tri) {
                        for (int k = 0; k < 3; ++k) {
                            const size_t v = static_cast<size_t>(vi[order[tri][k]]);
                            float co[3], no[3];
                            for (int c = 0; c < 3; ++c) {
                                co[c] = bf.get<float>(mvert, "MVert", "co", v, c);
                                no[c] = bf.get<int16_t>(mvert, "MVert", "no", v, c) / 32767.f;
                            }
                            for (int r = 0; r < 3; ++r) {
                                verts.push_back(T[r][0] * co[0] + T[r][1] * co[1] + T[r][2] * co[2] + T[r][3]); // aiMatrix4x4 * v
                                norms.push_back(T[r][0] * no[0] + T[r][1] * no[1] + T[r][2] * no[2]);           // aiMatrix3x3(T) * n
                            }
                        }
                        cols.insert(cols.end(), col.dif.begin(), col.dif.end());
                        mirs.insert(mirs.end(), col.mir.begin(), col.mir.end());
                        refls.push_back(col.refl);
                    }
                }
            }
        }
    }
    if (num_cameras != 1) return trn::fail(TRN_ERR_IO, "scene must contain exactly one camera (main.cpp:110)");
    if (out->num_lights > 1) return trn::fail(TRN_ERR_IO, "scene must contain at most one light (main.cpp:123)");
    const size_t n = cols.size() / 4;
    if (n == 0) return trn::fail(TRN_ERR_IO, "scene has no triangles");
    out->num_triangles = static_cast<uint32_t>(n);
    out->verts = static_cast<float*>(std::malloc(n * 9 * sizeof(float)));
    out->normals = static_cast<float*>(std::malloc(n * 9 * sizeof(float)));
    out->diffuse = static_cast<float*>(std::malloc(n * 4 * sizeof(float)));
    // per-corner layout above is v0.xyz n0.xyz interleaved by rows; regroup into 9+9
    for (size_t t = 0; t < n; ++t)
        for (int k = 0; k < 9; ++k) {
            out->verts[t * 9 + k] = verts[t * 9 + k];
            out->normals[t * 9 + k] = norms[t * 9 + k];
        }
    std::memcpy(out->diffuse, cols.data(), n * 4 * sizeof(float));
    out->reflective = static_cast<float*>(std::malloc(n * 4 * sizeof(float)));
    out->reflectivity = static_cast<float*>(std::malloc(n * sizeof(float)));
    std::memcpy(out->reflective, mirs.data(), n * 4 * sizeof(float));
    std::memcpy(out->reflectivity, refls.data(), n * sizeof(float));
    return TRN_OK;
}

// Neutral triangle-soup text format, one record per line, '#' starts a comment:
//   camera <16 floats: node transformation, row-major a1..d4> <hfov radians>
//   light  <x y z> <r g b a>                       (optional, at most one)
//   tri    <v0 v1 v2: 9 floats> <n0 n1 n2: 9 floats> <r g b a> [<reflective r g b a> <reflectivity>]
// Values are what main.cpp:25-82,109-136 would hand to the renderer (world space). Floats are parsed with strtof,
// so "%.9g" text round-trips fp32 exactly.
int32_t trn_load_soup(const char* path, trn_loaded_scene* out) {
    if (!path || !out) return trn::fail(TRN_ERR_INVALID, "null argument");
    std::memset(out, 0, sizeof *out);
    FILE* f = std::fopen(path, "r");
    if (!f) return trn::fail(TRN_ERR_IO, std::string("cannot open ") + path);
    std::vector<float> verts, norms, cols, mirs, refls;
    std::vector<char> line(1 << 16);
    int lineno = 0;
    auto parse = [](char* p, float* dst, int n) {
        for (int i = 0; i < n; ++i) {
            char* e = nullptr;
            dst[i] = std::strtof(p, &e);
            if (e == p) return false;
            p = e;
        }
        return true;
    };
    std::string err;
    while (std::fgets(line.data(), static_cast<int>(line.size()), f)) {
        ++lineno;
        char* p = line.data();
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '#' || *p == '\n' || *p == 0) continue;
        float v[27];
        if (std::strncmp(p, "tri", 3) == 0 && parse(p + 3, v, 22)) {
            verts.insert(verts.end(), v, v + 9);
            norms.insert(norms.end(), v + 9, v + 18);
            cols.insert(cols.end(), v + 18, v + 22);
            if (!parse(p + 3, v, 27)) std::fill(v + 22, v + 27, 0.f); // optional mirror material
            mirs.insert(mirs.end(), v + 22, v + 26);
            refls.push_back(v[26]);
        } else if (std::strncmp(p, "camera", 6) == 0 && parse(p + 6, v, 17)) {
            std::memcpy(out->cam_trafo4x4, v, 16 * sizeof(float));
            out->cam_hfov = v[16];
            out->cam_aspect = 0.f;
            out->has_camera += 1;
        } else if (std::strncmp(p, "light", 5) == 0 && parse(p + 5, v, 7)) {
            std::memcpy(out->light.pos, v, 3 * sizeof(float));
            std::memcpy(out->light.rgba, v + 3, 4 * sizeof(float));
            out->num_lights += 1;
        } else {
            err = std::string(path) + ":" + std::to_string(lineno) + ": malformed record";
            break;
        }
    }
    std::fclose(f);
    if (!err.empty()) return trn::fail(TRN_ERR_IO, err);
    if (out->has_camera != 1) return trn::fail(TRN_ERR_IO, "scene must contain exactly one camera (main.cpp:110)");
    if (out->num_lights > 1) return trn::fail(TRN_ERR_IO, "scene must contain at most one light (main.cpp:123)");
    const size_t n = cols.size() / 4;
    if (n == 0) return trn::fail(TRN_ERR_IO, "scene has no triangles");
    out->num_triangles = static_cast<uint32_t>(n);
    out->verts = static_cast<float*>(std::malloc(n * 9 * sizeof(float)));
    out->normals = static_cast<float*>(std::malloc(n * 9 * sizeof(float)));
    out->diffuse = static_cast<float*>(std::malloc(n * 4 * sizeof(float)));
    std::memcpy(out->verts, verts.data(), n * 9 * sizeof(float));
    std::memcpy(out->normals, norms.data(), n * 9 * sizeof(float));
    std::memcpy(out->diffuse, cols.data(), n * 4 * sizeof(float));
    out->reflective = static_cast<float*>(std::malloc(n * 4 * sizeof(float)));
    out->reflectivity = static_cast<float*>(std::malloc(n * sizeof(float)));
    std::memcpy(out->reflective, mirs.data(), n * 4 * sizeof(float));
    std::memcpy(out->reflectivity, refls.data(), n * sizeof(float));
    return TRN_OK;
}

void trn_loaded_scene_free(trn_loaded_scene* s) {
    if (!s) return;
    std::free(s->verts);
    std::free(s->normals);
    std::free(s->diffuse);
    std::free(s->reflective);
    std::free(s->reflectivity);
    std::memset(s, 0, sizeof *s);
}

} // extern "C"
