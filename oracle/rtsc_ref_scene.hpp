// oracle/rtsc_ref_scene.hpp - TEST INFRASTRUCTURE.  RTSC container (tests/helpers/crtscene.py) -> the REFERENCE's own scene<float>,
// built through the reference's own constructors (mirrors io/json/loader.hpp:149-265, which needs simdjson - absent here), plus
// the stbi_load stub the reference's bitmap loader calls (scene/texture/bitmap.hpp:15).  Shared by oracle/ref_harness.cpp (the
// compiled reference behind oracle/_ref) and tests/helpers/adapter_gpu.cpp (the reference's main.cpp flow over b200_accel).
// Include after <raytracer/scene/scene.hpp>; one translation unit per program (it defines stbi_load).
#pragma once

#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

using F = float;

// ---------------------------------------------------------------------------------------------------------------
// stbi stub: bitmaps come pre-decoded out of the RTSC texel blob
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct bitmap_blob { int w, h; std::vector<unsigned char> rgb; };
std::map<std::string, bitmap_blob>& bitmap_registry() { static std::map<std::string, bitmap_blob> r; return r; }
std::mutex registry_mutex;
}

extern "C" unsigned char* stbi_load(const char* filename, int* x, int* y, int* channels_in_file, int) {
    std::lock_guard g(registry_mutex);
    auto it = bitmap_registry().find(filename);
    if (it == bitmap_registry().end()) return nullptr;
    *x = it->second.w; *y = it->second.h; *channels_in_file = 3;
    auto* p = static_cast<unsigned char*>(std::malloc(it->second.rgb.size()));
    std::memcpy(p, it->second.rgb.data(), it->second.rgb.size());
    return p;
}
extern "C" void stbi_image_free(void* p) { std::free(p); }

// ---------------------------------------------------------------------------------------------------------------
// RTSC -> reference scene<float>
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct reader {
    std::vector<unsigned char> buf; std::size_t off = 0;
    template <typename T> T get() { T v; std::memcpy(&v, buf.data() + off, sizeof(T)); off += sizeof(T); return v; }
    void get_n(void* dst, std::size_t bytes) { std::memcpy(dst, buf.data() + off, bytes); off += bytes; }
};

struct tex_rec { uint32_t kind; float c0[3], c1[3], scalar; uint32_t w, h, off; };
struct mat_rec { uint32_t kind; float albedo[3], ior; uint32_t smooth; int32_t texture; };

scene<F> scene_from_rtsc(const char* path, const void* handle_tag) {
    reader r;
    {
        std::ifstream in(path, std::ios::binary);
        if (!in) throw std::runtime_error("cannot open RTSC");
        r.buf.assign(std::istreambuf_iterator<char>(in), {});
    }
    if (r.buf.size() < 8 || std::memcmp(r.buf.data(), "RTSC", 4) != 0) throw std::runtime_error("bad RTSC magic");
    r.off = 4;
    if (r.get<uint32_t>() != 1) throw std::runtime_error("bad RTSC version");

    scene<F> s{};
    float bg[3]; r.get_n(bg, 12);
    const uint32_t width = r.get<uint32_t>(), height = r.get<uint32_t>(), bucket = r.get<uint32_t>();
    s.config = settings<F>{color<F>{bg[0], bg[1], bg[2]}, height, width, bucket};  // field order: settings.hpp:9-12
    float cp[3]; r.get_n(cp, 12);
    std::array<F, 9> cm; r.get_n(cm.data(), 36);
    s.viewpoint = camera<F>{vec3<F>{cp[0], cp[1], cp[2]}, mat3<F>{cm}};
    const uint32_t nl = r.get<uint32_t>();
    for (uint32_t i = 0; i < nl; ++i) {
        float l[4]; r.get_n(l, 16);
        s.lights.push_back(light<F>{vec3<F>{l[0], l[1], l[2]}, l[3]});
    }
    const uint32_t nt = r.get<uint32_t>();
    std::vector<tex_rec> texs(nt);
    for (auto& t : texs) {
        t.kind = r.get<uint32_t>(); r.get_n(t.c0, 12); r.get_n(t.c1, 12); t.scalar = r.get<float>();
        t.w = r.get<uint32_t>(); t.h = r.get<uint32_t>(); t.off = r.get<uint32_t>();
    }
    const uint32_t nm = r.get<uint32_t>();
    std::vector<mat_rec> mats(nm);
    for (auto& m : mats) {
        m.kind = r.get<uint32_t>(); r.get_n(m.albedo, 12); m.ior = r.get<float>(); m.smooth = r.get<uint32_t>();
        m.texture = r.get<int32_t>();
    }
    const uint32_t nmesh = r.get<uint32_t>();
    struct head { uint32_t mat, nv, nuv, ntri; };
    std::vector<head> heads(nmesh);
    for (auto& h : heads) { h.mat = r.get<uint32_t>(); h.nv = r.get<uint32_t>(); h.nuv = r.get<uint32_t>(); h.ntri = r.get<uint32_t>(); }

    for (uint32_t mi = 0; mi < nmesh; ++mi) {
        const auto& h = heads[mi];
        std::vector<float> vb(3 * std::size_t(h.nv)), ub(2 * std::size_t(h.nuv));
        std::vector<uint32_t> tb(3 * std::size_t(h.ntri));
        r.get_n(vb.data(), vb.size() * 4); r.get_n(ub.data(), ub.size() * 4); r.get_n(tb.data(), tb.size() * 4);
        std::vector<vec3<F>> vertices(h.nv);
        for (uint32_t i = 0; i < h.nv; ++i) vertices[i] = vec3<F>{vb[3 * i], vb[3 * i + 1], vb[3 * i + 2]};
        std::vector<vec2<F>> uvs(h.nuv);
        for (uint32_t i = 0; i < h.nuv; ++i) uvs[i] = vec2<F>{ub[2 * i], ub[2 * i + 1]};
        std::vector<triangle<F>> tris;
        tris.reserve(h.ntri);
        for (uint32_t i = 0; i < h.ntri; ++i) {
            const std::size_t i0 = tb[3 * i], i1 = tb[3 * i + 1], i2 = tb[3 * i + 2];
            vec3<vec2<F>> tuv{};
            if (!uvs.empty()) tuv = vec3<vec2<F>>{uvs[i0], uvs[i1], uvs[i2]};       // loader.hpp:203-209
            tris.push_back(triangle<F>{vertices[i0], vertices[i1], vertices[i2], {i0, i1, i2}, mi, tuv});
        }
        s.meshes.emplace_back(mesh_object<F>{h.mat, vertices, uvs, tris});          // loader.hpp:226-232
    }
    const uint32_t ntexel = r.get<uint32_t>();
    const unsigned char* texels = r.buf.data() + r.off;
    r.off += ntexel;

    char tag[64];
    for (uint32_t i = 0; i < nt; ++i) {
        const auto& t = texs[i];
        const std::string name = "tex" + std::to_string(i);
        switch (t.kind) {
            case 0: s.textures.emplace(name, albedo_texture<F>{color<F>{t.c0[0], t.c0[1], t.c0[2]}}); break;
            case 1: s.textures.emplace(name, edge_texture<F>{color<F>{t.c0[0], t.c0[1], t.c0[2]},
                                                           color<F>{t.c1[0], t.c1[1], t.c1[2]}, t.scalar}); break;
            case 2: s.textures.emplace(name, checker_texture<F>{color<F>{t.c0[0], t.c0[1], t.c0[2]},
                                                              color<F>{t.c1[0], t.c1[1], t.c1[2]}, t.scalar}); break;
            case 3: {
                std::snprintf(tag, sizeof tag, "rtsc:%p:%u", handle_tag, i);
                {
                    std::lock_guard g(registry_mutex);
                    bitmap_blob b{int(t.w), int(t.h), {}};
                    b.rgb.assign(texels + t.off, texels + t.off + std::size_t(t.w) * t.h * 3);
                    bitmap_registry()[tag] = std::move(b);
                }
                s.textures.emplace(name, bitmap_texture<F>{std::string{tag}});        // -> load_bitmap -> stbi_load stub
                break;
            }
            default: throw std::invalid_argument("texture type unknown");
        }
    }
    for (const auto& m : mats) {
        const color<F> alb{m.albedo[0], m.albedo[1], m.albedo[2]};
        const bool smooth = m.smooth != 0;
        switch (m.kind) {
            case 0: s.materials.emplace_back(diffuse_material<F>{alb, smooth}); break;
            case 1: s.materials.emplace_back(reflective_material<F>{alb, smooth}); break;
            case 2: s.materials.emplace_back(refractive_material<F>{m.ior, smooth}); break;
            case 3: s.materials.emplace_back(constant_material<F>{alb, smooth}); break;
            case 4: s.materials.emplace_back(texture_material<F>{"tex" + std::to_string(m.texture), smooth}); break;
            default: throw std::invalid_argument("material type unknown");
        }
    }
    return s;
}

}  // namespace
