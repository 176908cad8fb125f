// host/scene_io.cpp - scene ingestion for the B200 backend: C-ABI description, RTSC container, .crtscene JSON.
//
// .crtscene semantics follow the reference loader, io/json/loader.hpp (cited per function).  Written from the
// schema, on our own JSON reader (json.hpp) because simdjson is not available offline.
#include "scene.hpp"

#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <sstream>

#include "json.hpp"

namespace rtb {

namespace {

std::string read_file(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw rt_error(RT_ERR_IO, "cannot open " + path);
    return std::string(std::istreambuf_iterator<char>(in), {});
}

// loader.hpp:9-17: every scalar goes double -> float
float narrow(const json::value& v) { return static_cast<float>(v.as_double()); }

void load3(const json::value& v, float* out) {                       // loader.hpp:19-26, :39-44
    const auto& a = v.as_array();
    if (a.size() < 3) throw json::parse_error("expected 3 numbers");
    for (int i = 0; i < 3; ++i) out[i] = narrow(a[i]);
}

}  // namespace

void validate_scene(const HostScene& s) {
    if (s.width == 0 || s.height == 0) throw rt_error(RT_ERR_BAD_ARG, "image size is zero");
    for (const auto& t : s.textures) {
        if (t.kind > RT_TEX_BITMAP) throw rt_error(RT_ERR_PARSE, "texture type unknown");
        if (t.kind == RT_TEX_BITMAP) {
            if (t.bmp_w == 0 || t.bmp_h == 0 || uint64_t(t.bmp_off) + uint64_t(t.bmp_w) * t.bmp_h * 3 > s.texels.size())
                throw rt_error(RT_ERR_BAD_ARG, "bitmap texels out of range");
        }
    }
    for (const auto& m : s.materials) {
        if (m.kind > RT_MAT_TEXTURE) throw rt_error(RT_ERR_PARSE, "material type unknown");
        if (m.kind == RT_MAT_TEXTURE && (m.texture < 0 || uint64_t(m.texture) >= s.textures.size()))
            throw rt_error(RT_ERR_BAD_ARG, "material references a missing texture");
    }
    for (const auto& m : s.meshes) {
        if (m.material >= s.materials.size()) throw rt_error(RT_ERR_BAD_ARG, "mesh material_index out of range");
        if (m.vertices.size() % 3 || m.uvs.size() % 2 || m.triangles.size() % 3)
            throw rt_error(RT_ERR_BAD_ARG, "mesh array length");
        const uint64_t nv = m.vertices.size() / 3, nuv = m.uvs.size() / 2;
        for (uint32_t i : m.triangles) {
            if (i >= nv) throw rt_error(RT_ERR_BAD_ARG, "triangle vertex index out of range");
            if (nuv && i >= nuv) throw rt_error(RT_ERR_BAD_ARG, "triangle uv index out of range");
        }
    }
    if (s.triangle_count() >= (1ull << 31)) throw rt_error(RT_ERR_BAD_ARG, "too many triangles");
}

HostScene scene_from_desc(const rt_scene_desc& d) {
    HostScene s;
    std::memcpy(s.background, d.background, 12);
    s.width = d.width; s.height = d.height; s.bucket_size = d.bucket_size;
    std::memcpy(s.camera_position, d.camera_position, 12);
    std::memcpy(s.camera_matrix, d.camera_matrix, 36);
    if ((d.n_lights && !d.lights) || (d.n_textures && !d.textures) || (d.n_materials && !d.materials) ||
        (d.n_meshes && !d.meshes) || (d.n_texel_bytes && !d.texels))
        throw rt_error(RT_ERR_BAD_ARG, "null array in scene description");
    s.lights.assign(d.lights, d.lights + d.n_lights);
    s.textures.assign(d.textures, d.textures + d.n_textures);
    s.materials.assign(d.materials, d.materials + d.n_materials);
    s.texels.assign(d.texels, d.texels + d.n_texel_bytes);
    s.meshes.resize(d.n_meshes);
    for (uint32_t i = 0; i < d.n_meshes; ++i) {
        const rt_mesh_desc& m = d.meshes[i];
        if ((m.n_vertices && !m.vertices) || (m.n_uvs && !m.uvs) || (m.n_triangles && !m.triangles))
            throw rt_error(RT_ERR_BAD_ARG, "null array in mesh description");
        s.meshes[i].material = m.material;
        s.meshes[i].vertices.assign(m.vertices, m.vertices + 3ull * m.n_vertices);
        if (m.n_uvs) s.meshes[i].uvs.assign(m.uvs, m.uvs + 2ull * m.n_uvs);
        s.meshes[i].triangles.assign(m.triangles, m.triangles + 3ull * m.n_triangles);
    }
    validate_scene(s);
    return s;
}

// the inverse of scene_from_rtsc: what a scene loaded from a .crtscene file looks like in the flat container
std::vector<uint8_t> scene_to_rtsc(const HostScene& s) {
    std::vector<uint8_t> out;
    auto put = [&](const void* src, uint64_t k) { const auto* b = static_cast<const uint8_t*>(src); out.insert(out.end(), b, b + k); };
    auto u32 = [&](uint64_t v) { const uint32_t w = uint32_t(v); put(&w, 4); };
    put("RTSC", 4); u32(1);
    put(s.background, 12);
    u32(s.width); u32(s.height); u32(s.bucket_size);
    put(s.camera_position, 12);
    put(s.camera_matrix, 36);
    u32(s.lights.size()); put(s.lights.data(), s.lights.size() * sizeof(rt_light_desc));
    u32(s.textures.size()); put(s.textures.data(), s.textures.size() * sizeof(rt_texture_desc));
    u32(s.materials.size()); put(s.materials.data(), s.materials.size() * sizeof(rt_material_desc));
    u32(s.meshes.size());
    for (const auto& m : s.meshes) { u32(m.material); u32(m.vertices.size() / 3); u32(m.uvs.size() / 2); u32(m.triangles.size() / 3); }
    for (const auto& m : s.meshes) { put(m.vertices.data(), m.vertices.size() * 4); put(m.uvs.data(), m.uvs.size() * 4); put(m.triangles.data(), m.triangles.size() * 4); }
    u32(s.texels.size()); put(s.texels.data(), s.texels.size());
    return out;
}

// RTSC v1 - layout documented in tests/helpers/crtscene.py
HostScene scene_from_rtsc(const void* bytes, uint64_t n) {
    const auto* p = static_cast<const uint8_t*>(bytes);
    uint64_t off = 0;
    auto get = [&](void* dst, uint64_t k) {
        if (off + k > n) throw rt_error(RT_ERR_PARSE, "RTSC truncated");
        std::memcpy(dst, p + off, k);
        off += k;
    };
    auto u32 = [&]() { uint32_t v; get(&v, 4); return v; };
    if (n < 8 || std::memcmp(p, "RTSC", 4) != 0) throw rt_error(RT_ERR_PARSE, "RTSC bad magic");
    off = 4;
    if (u32() != 1) throw rt_error(RT_ERR_PARSE, "RTSC unsupported version");
    HostScene s;
    get(s.background, 12);
    s.width = u32(); s.height = u32(); s.bucket_size = u32();
    get(s.camera_position, 12);
    get(s.camera_matrix, 36);
    static_assert(sizeof(rt_light_desc) == 16 && sizeof(rt_texture_desc) == 44 && sizeof(rt_material_desc) == 28);
    // a count is only believed when the bytes it promises are there (a 40-byte file must not ask for gigabytes)
    auto counted = [&](uint64_t elem_bytes) {
        const uint32_t k = u32();
        if (uint64_t(k) * elem_bytes > n - off) throw rt_error(RT_ERR_PARSE, "RTSC truncated");
        return k;
    };
    s.lights.resize(counted(sizeof(rt_light_desc)));
    get(s.lights.data(), s.lights.size() * sizeof(rt_light_desc));
    s.textures.resize(counted(sizeof(rt_texture_desc)));
    get(s.textures.data(), s.textures.size() * sizeof(rt_texture_desc));
    s.materials.resize(counted(sizeof(rt_material_desc)));
    get(s.materials.data(), s.materials.size() * sizeof(rt_material_desc));
    s.meshes.resize(counted(16));
    struct head { uint32_t mat, nv, nuv, nt; };
    std::vector<head> heads(s.meshes.size());
    get(heads.data(), heads.size() * sizeof(head));
    for (std::size_t i = 0; i < heads.size(); ++i) {
        auto& m = s.meshes[i];
        m.material = heads[i].mat;
        if (off + 12ull * heads[i].nv + 8ull * heads[i].nuv + 12ull * heads[i].nt > n) throw rt_error(RT_ERR_PARSE, "RTSC truncated");
        m.vertices.resize(3ull * heads[i].nv); get(m.vertices.data(), m.vertices.size() * 4);
        m.uvs.resize(2ull * heads[i].nuv); get(m.uvs.data(), m.uvs.size() * 4);
        m.triangles.resize(3ull * heads[i].nt); get(m.triangles.data(), m.triangles.size() * 4);
    }
    const uint32_t ntex = u32();
    if (off + ntex > n) throw rt_error(RT_ERR_PARSE, "RTSC truncated");
    s.texels.resize(ntex);
    get(s.texels.data(), ntex);
    if (off != n) throw rt_error(RT_ERR_PARSE, "RTSC trailing bytes");
    validate_scene(s);
    return s;
}

HostScene scene_from_crtscene(const std::string& path, const std::string& asset_root) {
    const std::string text = read_file(path);
    json::value doc;
    try {
        doc = json::parse(text);
        HostScene s;

        // load_settings, loader.hpp:46-60: bucket_size optional (must be an unsigned integer), default 64
        const json::value& st = doc.at("settings");
        load3(st.at("background_color"), s.background);
        const json::value& ims = st.at("image_settings");
        s.width = static_cast<uint32_t>(ims.at("width").as_u64());
        s.height = static_cast<uint32_t>(ims.at("height").as_u64());
        s.bucket_size = 64;
        if (const json::value* b = ims.find("bucket_size"); b && b->is_number() && b->is_integer)
            s.bucket_size = static_cast<uint32_t>(b->u64);

        // load_camera, loader.hpp:62-68
        const json::value& cam = doc.at("camera");
        load3(cam.at("position"), s.camera_position);
        const auto& cm = cam.at("matrix").as_array();
        if (cm.size() < 9) throw json::parse_error("camera.matrix needs 9 numbers");
        for (int i = 0; i < 9; ++i) s.camera_matrix[i] = narrow(cm[i]);

        // lights, loader.hpp:70-76, :246-248
        for (const json::value& l : doc.at("lights").as_array()) {
            rt_light_desc ld{};
            load3(l.at("position"), ld.position);
            ld.intensity = narrow(l.at("intensity"));
            s.lights.push_back(ld);
        }

        // textures, loader.hpp:78-106, :250-254 (optional array; unordered_map::emplace keeps the first of a name)
        std::map<std::string, int32_t> texture_by_name;
        if (const json::value* texs = doc.find("textures"); texs && texs->is_array()) {
            for (const json::value& t : *texs->arr) {
                rt_texture_desc td{};
                const std::string& type = t.at("type").as_string();
                if (type == "albedo") {
                    td.kind = RT_TEX_ALBEDO;
                    load3(t.at("albedo"), td.c0);
                } else if (type == "edges") {
                    td.kind = RT_TEX_EDGES;
                    load3(t.at("edge_color"), td.c0);
                    load3(t.at("inner_color"), td.c1);
                    td.scalar = narrow(t.at("edge_width"));
                } else if (type == "checker") {
                    td.kind = RT_TEX_CHECKER;
                    load3(t.at("color_A"), td.c0);
                    load3(t.at("color_B"), td.c1);
                    td.scalar = narrow(t.at("square_size"));
                } else if (type == "bitmap") {
                    td.kind = RT_TEX_BITMAP;
                    std::string fp = t.at("file_path").as_string();
                    if (!asset_root.empty() && !fp.empty() && fp[0] != '/') fp = asset_root + "/" + fp;
                    const Bitmap bm = load_bitmap_file(fp);
                    td.bmp_w = bm.w; td.bmp_h = bm.h; td.bmp_off = static_cast<uint32_t>(s.texels.size());
                    s.texels.insert(s.texels.end(), bm.rgb.begin(), bm.rgb.end());
                } else {
                    throw rt_error(RT_ERR_PARSE, "texture type unknown");                    // loader.hpp:104
                }
                texture_by_name.emplace(t.at("name").as_string(), static_cast<int32_t>(s.textures.size()));
                s.textures.push_back(td);
            }
        }

        // materials, loader.hpp:108-147, :256-258
        for (const json::value& m : doc.at("materials").as_array()) {
            rt_material_desc md{};
            md.texture = -1;
            md.ior = 1.0f;
            const std::string& type = m.at("type").as_string();
            if (type == "diffuse") {
                const json::value& alb = m.at("albedo");
                if (alb.is_array()) {
                    md.kind = RT_MAT_DIFFUSE;
                    load3(alb, md.albedo);
                } else if (alb.is_string()) {               // a named texture turns it into a texture material
                    md.kind = RT_MAT_TEXTURE;
                    auto it = texture_by_name.find(alb.str);
                    if (it == texture_by_name.end())        // the reference would throw from textures.at() mid-render
                        throw rt_error(RT_ERR_PARSE, "material references unknown texture \"" + alb.str + "\"");
                    md.texture = it->second;
                } else {
                    throw rt_error(RT_ERR_PARSE, "albedo neither array nor string");         // loader.hpp:126
                }
            } else if (type == "reflective") {
                md.kind = RT_MAT_REFLECTIVE;
                load3(m.at("albedo"), md.albedo);
            } else if (type == "refractive") {
                md.kind = RT_MAT_REFRACTIVE;
                md.ior = narrow(m.at("ior"));
            } else if (type == "constant") {
                md.kind = RT_MAT_CONSTANT;
                load3(m.at("albedo"), md.albedo);
            } else {
                throw rt_error(RT_ERR_PARSE, "material type unknown");                       // loader.hpp:145
            }
            md.smooth_shading = m.at("smooth_shading").as_bool() ? 1u : 0u;
            s.materials.push_back(md);
        }

        // objects, loader.hpp:149-233, :260-262
        for (const json::value& o : doc.at("objects").as_array()) {
            HostMesh mesh;
            mesh.material = static_cast<uint32_t>(o.at("material_index").as_u64());
            const auto& verts = o.at("vertices").as_array();
            if (verts.size() % 3) throw rt_error(RT_ERR_PARSE, "vertex coordinates not multiple of 3");
            mesh.vertices.reserve(verts.size());
            for (const auto& c : verts) mesh.vertices.push_back(narrow(c));
            if (const json::value* uvs = o.find("uvs"); uvs && uvs->is_array()) {
                if (uvs->arr->size() % 3) throw rt_error(RT_ERR_PARSE, "uv coordinates not multiple of 3");
                for (std::size_t i = 0; i + 2 < uvs->arr->size(); i += 3) {     // third component dropped, :179-186
                    mesh.uvs.push_back(narrow((*uvs->arr)[i]));
                    mesh.uvs.push_back(narrow((*uvs->arr)[i + 1]));
                    (void)(*uvs->arr)[i + 2].as_double();
                }
            }
            const auto& tris = o.at("triangles").as_array();
            if (tris.size() % 3) throw rt_error(RT_ERR_PARSE, "triangle indices not multiple of 3");
            mesh.triangles.reserve(tris.size());
            for (const auto& c : tris) mesh.triangles.push_back(static_cast<uint32_t>(c.as_u64()));
            s.meshes.push_back(std::move(mesh));
        }
        validate_scene(s);
        return s;
    } catch (const json::parse_error& e) {
        throw rt_error(RT_ERR_PARSE, path + ": " + e.what());
    }
}

// ---------------------------------------------------------------------------------------------------------------
// bitmap files.  The reference decodes through stb_image (scene/texture/bitmap.hpp:11-37), which is not vendored
// and not available offline.  JPEG (the format of the one bitmap the reference ships) is decoded by
// host/jpeg_decode.cpp, whose arithmetic reproduces the bitmap quadrant of the published outputs/textures.png
// exactly; binary PPM (P6) is read natively; for any other file a pre-decoded sidecar "<file>.ppm" is used when
// present.  Everything downstream of the texel bytes is exact.
// ---------------------------------------------------------------------------------------------------------------
namespace {
bool parse_ppm(const std::string& data, Bitmap& out) {
    if (data.size() < 2 || data[0] != 'P' || data[1] != '6') return false;
    std::size_t pos = 2;
    auto next_int = [&](uint32_t& v) -> bool {
        for (;;) {
            while (pos < data.size() && (data[pos] == ' ' || data[pos] == '\n' || data[pos] == '\r' || data[pos] == '\t')) ++pos;
            if (pos < data.size() && data[pos] == '#') { while (pos < data.size() && data[pos] != '\n') ++pos; continue; }
            break;
        }
        if (pos >= data.size() || data[pos] < '0' || data[pos] > '9') return false;
        uint64_t acc = 0;
        while (pos < data.size() && data[pos] >= '0' && data[pos] <= '9') { acc = acc * 10 + uint64_t(data[pos] - '0'); if (acc > 1u << 30) return false; ++pos; }
        v = static_cast<uint32_t>(acc);
        return true;
    };
    uint32_t w, h, maxv;
    if (!next_int(w) || !next_int(h) || !next_int(maxv) || maxv != 255 || pos >= data.size()) return false;
    ++pos;  // single whitespace after maxval
    const uint64_t need = uint64_t(w) * h * 3;
    if (data.size() - pos < need) return false;
    out.w = w; out.h = h;
    out.rgb.assign(data.begin() + pos, data.begin() + pos + need);
    return true;
}
}  // namespace

Bitmap load_bitmap_file(const std::string& path) {
    Bitmap bm;
    std::string data;
    bool opened = true;
    try { data = read_file(path); } catch (const rt_error&) { opened = false; }
    if (opened && parse_ppm(data, bm)) return bm;
    if (opened && data.size() > 3 && uint8_t(data[0]) == 0xFF && uint8_t(data[1]) == 0xD8) return decode_jpeg(data, path);
    try {
        data = read_file(path + ".ppm");
        if (parse_ppm(data, bm)) return bm;
    } catch (const rt_error&) {}
    if (!opened) throw rt_error(RT_ERR_IO, "cannot open bitmap " + path);
    throw rt_error(RT_ERR_UNSUPPORTED, "bitmap " + path + ": only JPEG (baseline), binary PPM (P6), or a '" + path + ".ppm' sidecar, can be decoded");
}

}  // namespace rtb
