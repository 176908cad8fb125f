// host/scene.hpp - host-side scene model of the B200 backend.
//
// Plays the role of the reference's scene<F> (scene/scene.hpp:14-22) for F = float, already flattened to the
// plain arrays the C ABI carries (include/rt_b200.h, rt_scene_desc): textures are indexed instead of keyed by name,
// all bitmap texels live in one RGB8 blob.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"

namespace rtb {

struct rt_error : std::runtime_error {
    int status;
    rt_error(int status, const std::string& what) : std::runtime_error(what), status(status) {}
};

struct HostMesh {
    uint32_t material = 0;
    std::vector<float> vertices;     // 3 per vertex
    std::vector<float> uvs;          // 2 per uv
    std::vector<uint32_t> triangles; // 3 per triangle
};

struct HostScene {
    float background[3] = {0, 0, 0};
    uint32_t width = 0, height = 0, bucket_size = 64;
    float camera_position[3] = {0, 0, 0};
    float camera_matrix[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::vector<rt_light_desc> lights;
    std::vector<rt_texture_desc> textures;
    std::vector<rt_material_desc> materials;
    std::vector<HostMesh> meshes;
    std::vector<uint8_t> texels;

    uint64_t triangle_count() const {
        uint64_t n = 0;
        for (const auto& m : meshes) n += m.triangles.size() / 3;
        return n;
    }
    uint64_t vertex_count() const {
        uint64_t n = 0;
        for (const auto& m : meshes) n += m.vertices.size() / 3;
        return n;
    }
};

// validates indices (the reference indexes unchecked and would crash) and copies the caller's arrays
HostScene scene_from_desc(const rt_scene_desc& d);
HostScene scene_from_rtsc(const void* bytes, uint64_t n);
std::vector<uint8_t> scene_to_rtsc(const HostScene& s);
// .crtscene JSON, semantics of io/json/loader.hpp:235-265
HostScene scene_from_crtscene(const std::string& path, const std::string& asset_root);
void validate_scene(const HostScene& s);

// decoded image for bitmap textures: RGB8 row major
struct Bitmap { uint32_t w = 0, h = 0; std::vector<uint8_t> rgb; };
Bitmap load_bitmap_file(const std::string& path);
// baseline JPEG bytes -> RGB8 (host/jpeg_decode.cpp); `what` names the file in error messages
Bitmap decode_jpeg(const std::string& file, const std::string& what);

}  // namespace rtb
