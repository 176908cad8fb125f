// host/kd_build.cpp - geometry preparation, kd-tree build, flattening.  See kd_build.hpp.
// Compiled with -ffp-contract=off: triangle / vertex normals must carry the same bits as the canonical
// (uncontracted) reference build, they feed the shading of every hit.
#include "kd_build.hpp"
#include "kd_parallel.hpp"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <utility>

namespace rtb {

namespace {

struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// core/math/vec3.hpp:104-108: x * (1 / sqrt(x*x + y*y + z*z)), sum left to right
inline V3 normalized(V3 a) {
    const float inv = 1.0f / std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return {a.x * inv, a.y * inv, a.z * inv};
}
// std::min / std::max argument order matters for NaNs and signed zeros: aabb3.hpp:25-41
inline float min_std(float a, float b) { return (b < a) ? b : a; }
inline float max_std(float a, float b) { return (a < b) ? b : a; }
inline void grow(float* lo, float* hi, V3 p) {
    lo[0] = min_std(lo[0], p.x); lo[1] = min_std(lo[1], p.y); lo[2] = min_std(lo[2], p.z);
    hi[0] = max_std(hi[0], p.x); hi[1] = max_std(hi[1], p.y); hi[2] = max_std(hi[2], p.z);
}
inline void empty_box(float* lo, float* hi) {
    for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
}
// aabb3.hpp:68-72, closed intervals
inline bool overlaps(const float* alo, const float* ahi, const float* blo, const float* bhi) {
    return (blo[0] <= ahi[0] && alo[0] <= bhi[0]) && (blo[1] <= ahi[1] && alo[1] <= bhi[1]) &&
           (blo[2] <= ahi[2] && alo[2] <= bhi[2]);
}

}  // namespace

Geometry prepare_geometry(const HostScene& s) {
    Geometry g;
    g.tris.reserve(s.triangle_count());
    g.vertex_normals.assign(3 * s.vertex_count(), 0.0f);
    empty_box(g.root_min, g.root_max);

    uint64_t vertex_base = 0;
    for (uint32_t mi = 0; mi < s.meshes.size(); ++mi) {
        const HostMesh& m = s.meshes[mi];
        const uint64_t nv = m.vertices.size() / 3, nt = m.triangles.size() / 3;
        const bool has_uv = !m.uvs.empty();
        float mlo[3], mhi[3];
        empty_box(mlo, mhi);
        float* vn = g.vertex_normals.data() + 3 * vertex_base;
        auto vertex = [&](uint32_t i) { return V3{m.vertices[3 * i], m.vertices[3 * i + 1], m.vertices[3 * i + 2]}; };
        for (uint64_t t = 0; t < nt; ++t) {
            const uint32_t i0 = m.triangles[3 * t], i1 = m.triangles[3 * t + 1], i2 = m.triangles[3 * t + 2];
            const V3 a = vertex(i0), b = vertex(i1), c = vertex(i2);
            TriGeom tg{};
            const V3 e1 = sub(b, a), e2 = sub(c, a);
            const V3 n = normalized(cross(e1, e2));                       // triangle.hpp:22, mesh.hpp:34
            tg.v0[0] = a.x; tg.v0[1] = a.y; tg.v0[2] = a.z;
            tg.e1[0] = e1.x; tg.e1[1] = e1.y; tg.e1[2] = e1.z;
            tg.e2[0] = e2.x; tg.e2[1] = e2.y; tg.e2[2] = e2.z;
            tg.normal[0] = n.x; tg.normal[1] = n.y; tg.normal[2] = n.z;
            empty_box(tg.bmin, tg.bmax);
            grow(tg.bmin, tg.bmax, a); grow(tg.bmin, tg.bmax, b); grow(tg.bmin, tg.bmax, c);
            tg.vi[0] = uint32_t(vertex_base + i0); tg.vi[1] = uint32_t(vertex_base + i1); tg.vi[2] = uint32_t(vertex_base + i2);
            tg.mesh = mi;
            if (has_uv) {
                tg.uv[0] = m.uvs[2 * i0]; tg.uv[1] = m.uvs[2 * i0 + 1];
                tg.uv[2] = m.uvs[2 * i1]; tg.uv[3] = m.uvs[2 * i1 + 1];
                tg.uv[4] = m.uvs[2 * i2]; tg.uv[5] = m.uvs[2 * i2 + 1];
            }
            g.tris.push_back(tg);
            grow(mlo, mhi, a); grow(mlo, mhi, b); grow(mlo, mhi, c);
            // un-weighted accumulation in triangle order, mesh.hpp:36-38
            for (uint32_t vi : {i0, i1, i2}) { vn[3 * vi] += n.x; vn[3 * vi + 1] += n.y; vn[3 * vi + 2] += n.z; }
        }
        for (uint64_t v = 0; v < nv; ++v) {                                // mesh.hpp:41-43
            const V3 q = normalized(V3{vn[3 * v], vn[3 * v + 1], vn[3 * v + 2]});
            vn[3 * v] = q.x; vn[3 * v + 1] = q.y; vn[3 * v + 2] = q.z;
        }
        for (int a = 0; a < 3; ++a) {                                      // unite, aabb3.hpp:34-41
            g.root_min[a] = min_std(g.root_min[a], mlo[a]);
            g.root_max[a] = max_std(g.root_max[a], mhi[a]);
        }
        vertex_base += nv;
    }
    return g;
}

namespace {

// The reference's builder as a policy of the parallel driver (kd_parallel.hpp): build_tree, kd_tree_simd.hpp:146-185.
struct MedianPolicy {
    struct Work {
        uint64_t depth = 0;
        float lo[3], hi[3];
        std::vector<uint32_t> tris;
        size_t size() const { return tris.size(); }
        bool empty() const { return tris.empty(); }
    };
    const Geometry& g;
    uint32_t max_depth, max_leaf_size;

    bool expand(Work& w, uint32_t& axis_out, float& split_out, Work& c0, Work& c1) const {
        if (w.depth == max_depth || w.tris.size() <= max_leaf_size) return false;         // kd_tree_simd.hpp:147
        // aabb3::split, aabb3.hpp:43-60: median of the box on axis depth%3; a zero-width axis defers to the next.
        // (The reference would recurse forever on a point-sized box; we stop after trying all three.)
        uint32_t axis = uint32_t(w.depth % 3);
        for (int tries = 0; tries < 3 && w.lo[axis] == w.hi[axis]; ++tries) axis = (axis + 1u) % 3u;
        const float mid = w.lo[axis] + ((w.hi[axis] - w.lo[axis]) / 2.0f);
        axis_out = axis; split_out = mid;
        c0.depth = c1.depth = w.depth + 1;
        std::memcpy(c0.lo, w.lo, 12); std::memcpy(c0.hi, w.hi, 12);
        std::memcpy(c1.lo, w.lo, 12); std::memcpy(c1.hi, w.hi, 12);
        c0.hi[axis] = mid;
        c1.lo[axis] = mid;
        c0.tris.reserve(w.tris.size());
        c1.tris.reserve(w.tris.size());
        for (uint32_t id : w.tris) {                                       // no clipping: refs are duplicated (:160-170)
            const TriGeom& tg = g.tris[id];
            if (overlaps(c0.lo, c0.hi, tg.bmin, tg.bmax)) c0.tris.push_back(id);
            if (overlaps(c1.lo, c1.hi, tg.bmin, tg.bmax)) c1.tris.push_back(id);
        }
        c0.tris.shrink_to_fit();
        c1.tris.shrink_to_fit();
        return true;
    }
    void emit(const Work& w, std::vector<uint32_t>& refs) const { refs.insert(refs.end(), w.tris.begin(), w.tris.end()); }
};

}  // namespace

KdTree build_kd_tree(const Geometry& g, uint32_t max_depth, uint32_t max_leaf_size) {
    MedianPolicy pol{g, max_depth, max_leaf_size};
    MedianPolicy::Work root;
    root.depth = 0;
    std::memcpy(root.lo, g.root_min, 12); std::memcpy(root.hi, g.root_max, 12);
    root.tris.resize(g.tris.size());
    for (uint32_t i = 0; i < root.tris.size(); ++i) root.tris[i] = i;
    return kd_build_parallel(pol, std::move(root));
}

DeviceLayout flatten_tree_only(const Geometry& g, const KdTree& t) {
    DeviceLayout d;
    auto bits = [](float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; };
    const uint64_t n_nodes = t.nodes.size();
    d.nodes8.resize(2 * n_nodes);
    d.nodes32.resize(8 * n_nodes);

    // packets are laid out leaf after leaf in node order (= leaf creation order)
    uint64_t n_packets = 0;
    for (const KdNode& n : t.nodes)
        if (n.first_ref != KD_NONE) n_packets += (n.ref_count + PACKET_LANES - 1) / PACKET_LANES;
    d.n_packets = n_packets;
    d.packets.assign(uint64_t(PACKET_WORDS) * (n_packets ? n_packets : 1), 0u);

    std::vector<uint64_t> first_packet(n_nodes, 0);
    {
        uint64_t next = 0;
        for (uint64_t i = 0; i < n_nodes; ++i) {
            const KdNode& n = t.nodes[i];
            first_packet[i] = next;
            if (n.first_ref != KD_NONE) next += (n.ref_count + PACKET_LANES - 1) / PACKET_LANES;
        }
        if (next >= (1ull << 32)) throw rt_error(RT_ERR_UNSUPPORTED, "too many leaf packets");
    }
    for (const KdNode& n : t.nodes) {
        if (n.first_ref == KD_NONE) {
            if (n.child0 != KD_NONE && n.child0 != uint64_t(&n - t.nodes.data()) + 1) throw rt_error(RT_ERR_BAD_ARG, "kd-tree is not in DFS pre-order");
            if (n.child1 != KD_NONE && n.child1 >= (1ull << 28)) throw rt_error(RT_ERR_UNSUPPORTED, "kd-tree has too many nodes");
        } else if ((n.ref_count + PACKET_LANES - 1) / PACKET_LANES >= (1ull << 30)) throw rt_error(RT_ERR_UNSUPPORTED, "too many leaf packets");
    }
    parallel_for(n_nodes, 4096, [&](uint64_t nb, uint64_t ne) {
    for (uint64_t i = nb; i < ne; ++i) {
        const KdNode& n = t.nodes[i];
        uint32_t first, word;
        if (n.first_ref == KD_NONE) {
            first = bits(n.split);
            word = (n.axis & 3u) | (n.child0 != KD_NONE ? 4u : 0u) | (n.child1 != KD_NONE ? 8u : 0u) |
                   (n.child1 != KD_NONE ? uint32_t(n.child1) << 4 : 0u);
        } else {
            const uint64_t count = (n.ref_count + PACKET_LANES - 1) / PACKET_LANES;
            const uint64_t next_packet = first_packet[i];
            first = uint32_t(next_packet);
            word = 3u | (uint32_t(count) << 2);
            for (uint64_t k = 0; k < count * PACKET_LANES; ++k) {
                const uint64_t r = k < n.ref_count ? k : n.ref_count - 1;     // pad with the leaf's last triangle
                const uint32_t id = t.refs[n.first_ref + r];
                const TriGeom& tg = g.tris[id];
                uint32_t* p = d.packets.data() + (next_packet + k / PACKET_LANES) * PACKET_WORDS;
                const uint32_t lane = uint32_t(k % PACKET_LANES);
                for (int c = 0; c < 3; ++c) {
                    p[(0 + c) * 4 + lane] = bits(tg.v0[c]);
                    p[(3 + c) * 4 + lane] = bits(tg.e1[c]);
                    p[(6 + c) * 4 + lane] = bits(tg.e2[c]);
                }
                p[9 * 4 + lane] = id;
            }
        }
        d.nodes8[2 * i] = first;
        d.nodes8[2 * i + 1] = word;
        uint32_t* f = d.nodes32.data() + 8 * i;
        f[0] = bits(n.bmin[0]); f[1] = bits(n.bmin[1]); f[2] = bits(n.bmin[2]); f[3] = first;
        f[4] = bits(n.bmax[0]); f[5] = bits(n.bmax[1]); f[6] = bits(n.bmax[2]); f[7] = word;
    }
    });

    return d;
}

DeviceLayout flatten(const HostScene& s, const Geometry& g, const KdTree& t) {
    DeviceLayout d = flatten_tree_only(g, t);
    const uint64_t nt = g.tris.size();
    d.tri_index.resize(4 * (nt ? nt : 1));
    d.tri_normal.resize(4 * (nt ? nt : 1));
    d.tri_uv.resize(8 * (nt ? nt : 1));
    parallel_for(nt, 1 << 16, [&](uint64_t tb, uint64_t te) {
    for (uint64_t i = tb; i < te; ++i) {
        const TriGeom& tg = g.tris[i];
        d.tri_index[4 * i] = tg.vi[0]; d.tri_index[4 * i + 1] = tg.vi[1]; d.tri_index[4 * i + 2] = tg.vi[2];
        d.tri_index[4 * i + 3] = s.meshes[tg.mesh].material;
        d.tri_normal[4 * i] = tg.normal[0]; d.tri_normal[4 * i + 1] = tg.normal[1]; d.tri_normal[4 * i + 2] = tg.normal[2];
        d.tri_normal[4 * i + 3] = 0.0f;
        for (int c = 0; c < 6; ++c) d.tri_uv[8 * i + c] = tg.uv[c];
        d.tri_uv[8 * i + 6] = d.tri_uv[8 * i + 7] = 0.0f;
    }
    });
    const uint64_t nv = g.vertex_normals.size() / 3;
    d.vnormals.resize(4 * (nv ? nv : 1));
    parallel_for(nv, 1 << 16, [&](uint64_t vb, uint64_t ve) {
    for (uint64_t i = vb; i < ve; ++i) {
        d.vnormals[4 * i] = g.vertex_normals[3 * i]; d.vnormals[4 * i + 1] = g.vertex_normals[3 * i + 1];
        d.vnormals[4 * i + 2] = g.vertex_normals[3 * i + 2]; d.vnormals[4 * i + 3] = 0.0f;
    }
    });
    for (const auto& m : s.materials) if (m.kind == RT_MAT_REFRACTIVE) d.has_transmissive = true;
    return d;
}

}  // namespace rtb
