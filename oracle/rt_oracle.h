/* oracle/rt_oracle.h - TEST INFRASTRUCTURE.
 *
 * CPU restatement, in plain C, of the reference's kd_tree_simd_accel hot path and of the render loops that call it
 * (/root/reference/include/raytracer/render/accel/kd_tree_simd.hpp, core/math/aabb3.hpp, render/render.hpp,
 * scene/texture/ headers).  Every function in rt_oracle.c cites the reference lines it follows.
 *
 * Pinned (see DESIGN.md "Oracle"): bit-exact against the UNMODIFIED reference compiled into oracle/_ref (every
 * closest-hit query of a frame, inputs and outputs, on all four reference configs, incl. the GI / multi-sample
 * paths via the reference's own minstd sequence), and against the reference's published golden
 * outputs/refractive_dragon.png (0 differing pixels) through tests/golden/golden.json.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (simd-raytracer_b200/) never links, loads or calls it.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ro_scene ro_scene;

enum { RO_RNG_MINSTD = 0, RO_RNG_PHILOX = 1 };
enum { RO_KIND_PRIMARY = 0, RO_KIND_SHADOW = 1, RO_KIND_REFLECT = 2, RO_KIND_REFRACT = 3, RO_KIND_GI = 4 };

typedef struct ro_params {
    double fov_degrees;        /* config.hpp:6  */
    float eps;                 /* config.hpp:8, narrowed as main.cpp:37 */
    float shadow_bias;         /* config.hpp:9  */
    float reflection_bias;     /* config.hpp:10 */
    float refraction_bias;     /* config.hpp:11 */
    uint32_t spp;              /* config.hpp:13 */
    uint32_t max_ray_depth;    /* config.hpp:14 */
    uint32_t gi_rays;          /* config.hpp:15 */
    uint32_t seed;             /* config.hpp:17 */
    uint32_t rng;              /* RO_RNG_*: minstd = the reference's sequence (single thread only) */
    uint32_t sample_offset;    /* philox only: first sample index of this slice (multi-GPU spp sharding) */
    uint32_t spp_total;        /* philox only: divisor / jitter switch for sliced renders; 0 = spp */
    uint32_t raw_sum;          /* 1: write the slice's sample SUM (the multi-GPU combine divides after reducing) */
} ro_params;

typedef struct ro_record {
    float o[3], d[3];
    float t, u, v;
    int32_t tri;               /* -1 = miss */
    uint32_t cull;
    uint32_t kind;             /* RO_KIND_* */
} ro_record;

/* counts[]: 0 primary, 1 primary hits, 2 shadow queries, 3 shadow-query hits, 4 secondary (reflect+refract+GI),
 *           5 secondary hits, 6 node visits (slab tests), 7 triangle tests (real, unpadded) - all kinds;
 *           8 / 9 node visits / triangle tests of the shadow queries, 10 / 11 of the secondary queries
 *           (primary = total - shadow - secondary).  These are the reference algorithm's own work counts: the
 *           fixed denominators of the roofline. */
enum { RO_N_COUNTS = 12 };

void ro_default_params(ro_params* p);

ro_scene* ro_scene_load_rtsc(const char* path, uint32_t kd_max_depth, uint32_t kd_max_leaf);
ro_scene* ro_scene_from_rtsc_bytes(const void* bytes, uint64_t n, uint32_t kd_max_depth, uint32_t kd_max_leaf);
void ro_scene_free(ro_scene* s);

/* info[]: width, height, n_tris, n_nodes, n_leaf_refs, n_leaves, max_leaf_refs, tree_depth */
void ro_scene_info(const ro_scene* s, uint64_t* info);

/* tree dump: node5 = parent, child0, child1, first_ref, ref_count per node (UINT64_MAX = none / inner),
 * boxes = 6 floats per node, refs = triangle index per leaf reference */
void ro_tree(const ro_scene* s, uint64_t* node5, float* boxes, uint32_t* refs);

/* derived per-triangle / per-vertex data, for flattener parity: tri9 = v0,e1,e2; normals = face normal;
 * vnormals = per (mesh-concatenated) vertex normal */
void ro_geometry(const ro_scene* s, float* tri9, float* face_normals, float* vertex_normals, uint32_t* tri_vidx,
                 uint32_t* tri_mesh);

void ro_trace(const ro_scene* s, float eps, const float* rays6, uint64_t n, int cull, float* tuv, int32_t* tri,
              uint64_t* counts);
void ro_occluded(const ro_scene* s, const ro_params* p, const float* rays6, const float* max_t, uint64_t n,
                 uint8_t* out, uint64_t* counts);
void ro_primary_rays(const ro_scene* s, const ro_params* p, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                     float* rays6);

/* full frame (or the tile [x0,x1)x[y0,y1) of it; pixels outside stay untouched).  n_threads<=0 -> all cores.
 * RO_RNG_MINSTD forces one thread and row-major order (= the reference in SINGLE_TILE mode). */
void ro_render(const ro_scene* s, const ro_params* p, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
               float* rgb, int n_threads, uint64_t* counts);

/* single-threaded frame that logs every closest-hit query in issue order */
uint64_t ro_record_frame(const ro_scene* s, const ro_params* p, ro_record* out, uint64_t cap, float* rgb);

/* the counter-based RNG shared (by specification, not by code) with the CUDA path */
void ro_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
