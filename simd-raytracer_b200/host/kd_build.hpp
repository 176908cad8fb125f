// host/kd_build.hpp - host kd-tree builder and flattener of the B200 backend.
//
// The tree must be THE reference's tree (same median splits, same closed-interval overlap test, same DFS
// pre-order numbering, same leaf lists): tie-breaking between exactly-equal hit distances depends on leaf order
// (SURVEY.md section 7 "tie-breaking").  Reference: kd_tree_simd_accel ctor / build_tree / build_tree_leaf,
// render/accel/kd_tree_simd.hpp:100-185, and aabb3::split / intersect(aabb), core/math/aabb3.hpp:43-72.
#pragma once

#include <cstdint>
#include <vector>

#include "scene.hpp"

namespace rtb {

constexpr uint64_t KD_NONE = ~0ull;

struct TriGeom {                 // scene/primitive/triangle.hpp:10-30, the fields the hot path and shading read
    float v0[3], e1[3], e2[3];
    float normal[3];
    float bmin[3], bmax[3];
    uint32_t vi[3];              // indices into the mesh-concatenated vertex-normal table
    uint32_t mesh;
    float uv[6];
};

struct Geometry {
    std::vector<TriGeom> tris;           // global triangle id = position in the mesh-order concatenation
    std::vector<float> vertex_normals;   // 3 per vertex, scene/object/mesh.hpp:23-44
    float root_min[3], root_max[3];      // union of the mesh boxes, kd_tree_simd.hpp:101-104
};

struct KdNode {                  // kd_tree_simd_accel::node, kd_tree_simd.hpp:75-84 (packs counted in triangle refs)
    uint64_t parent, child0, child1, first_ref, ref_count;
    float bmin[3], bmax[3];
    uint32_t axis;               // split axis actually used (after the degenerate-axis fall-through)
    float split;
};

struct KdTree {
    std::vector<KdNode> nodes;
    std::vector<uint32_t> refs;  // leaf triangle lists, concatenated in leaf creation order
    uint64_t n_leaves = 0, max_leaf_refs = 0, depth = 0;
};

Geometry prepare_geometry(const HostScene& s);
KdTree build_kd_tree(const Geometry& g, uint32_t max_depth, uint32_t max_leaf_size);

// ---- device layout -------------------------------------------------------------------------------------------
//
// nodes8    8 bytes per node: { f32 split | u32 first_packet, u32 word }
//             word[1:0] = split axis, 3 = leaf
//             inner: word[2] = has child0 (always at index+1), word[3] = has child1, word[31:4] = child1 index
//             leaf : word[31:2] = packet count, first u32 = first packet index
// nodes32   the same node with its box, for the reference-order traversal (which slab-tests every node box,
//           kd_tree_simd.hpp:202): 2 x float4 = { min.xyz, first_u32 } { max.xyz, word }
// packets   leaf triangles in groups of four, structure-of-arrays, every row one aligned 16-byte vector:
//             v0x[4] v0y[4] v0z[4] e1x[4] e1y[4] e1z[4] e2x[4] e2y[4] e2z[4] tri_id[4]     (160 B)
//           the last packet of a leaf repeats the leaf's last triangle, as the reference pads its W-wide packs
//           (kd_tree_simd.hpp:123)
// tri_shade per triangle: { vi0, vi1, vi2, material } + face normal (float4) + uvs (2 x float4)
// vnormals  float4 per vertex
constexpr uint32_t PACKET_WORDS = 40;
constexpr uint32_t PACKET_LANES = 4;

struct DeviceLayout {
    std::vector<uint32_t> nodes8;      // 2 words per node
    std::vector<uint32_t> nodes32;     // 8 words per node
    std::vector<uint32_t> packets;     // PACKET_WORDS per packet
    std::vector<uint32_t> tri_index;   // 4 words per triangle
    std::vector<float> tri_normal;     // 4 floats per triangle
    std::vector<float> tri_uv;         // 8 floats per triangle
    std::vector<float> vnormals;       // 4 floats per vertex
    uint64_t n_packets = 0;
    bool has_transmissive = false;     // any refractive material (shadow rays may need the pass-through loop)
};

DeviceLayout flatten(const HostScene& s, const Geometry& g, const KdTree& t);
DeviceLayout flatten_tree_only(const Geometry& g, const KdTree& t);   // nodes8 / nodes32 / packets only

// the bounding-volume hierarchy of the accelerated mode (bvh_build.cpp, csrc/rt_bvh.cuh): same KdTree container (node boxes
// are the tight bounds, `axis`/`split` record the SAH decision), flattened to 64-byte two-child nodes
KdTree build_bvh(const Geometry& g, uint32_t max_leaf);
struct BvhLayout {
    std::vector<uint32_t> nodes;       // 16 words per inner node
    std::vector<uint32_t> tris;        // 12 words per triangle record, leaf order
    uint64_t n_nodes = 0, n_refs = 0;
    float root_min[3] = {3.0e38f, 3.0e38f, 3.0e38f}, root_max[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
};
BvhLayout flatten_bvh(const Geometry& g, const KdTree& t);

}  // namespace rtb
