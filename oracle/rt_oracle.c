/* oracle/rt_oracle.c - TEST INFRASTRUCTURE: CPU restatement of the reference hot path.  See rt_oracle.h.
 *
 * Build: oracle/Makefile (gcc -std=c11 -O2 -ffp-contract=off; NEVER -ffast-math, never FMA contraction: the
 * reference's published golden is only reproduced without contraction).
 *
 * All reference citations are path:line under /root/reference/include/raytracer/.  "L->R" = left-associative as
 * the reference spells the expression.  Scalar float arithmetic on x86-64 is SSE, i.e. IEEE binary32 with
 * round-to-nearest-even, subnormals kept.
 */
#define _GNU_SOURCE
#include "rt_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define RO_EMPTY UINT64_MAX

typedef struct { float x, y, z; } v3;
typedef struct { float r, g, b; } col;
typedef struct { v3 o, d, inv; } ray;

typedef struct {
    v3 v0, e1, e2, normal;
    uint32_t vi[3];   /* index into the concatenated vertex-normal table */
    uint32_t mesh;
    float bmin[3], bmax[3];
    float uv[6];      /* uv0.xy uv1.xy uv2.xy */
} tri_t;

typedef struct {
    uint64_t parent, child0, child1, first_ref, ref_count;
    float bmin[3], bmax[3];
} node_t;

typedef struct { uint32_t kind; float c0[3], c1[3], scalar; uint32_t w, h, off; } tex_t;
typedef struct { uint32_t kind; float albedo[3], ior; uint32_t smooth; int32_t texture; } mat_t;
typedef struct { float pos[3], intensity; } light_t;

struct ro_scene {
    float bg[3];
    uint32_t width, height, bucket;
    float cam_pos[3], cam_m[9];
    uint32_t n_lights, n_tex, n_mat, n_mesh;
    light_t* lights; tex_t* tex; mat_t* mat;
    uint32_t* mesh_mat;
    uint8_t* texels; uint64_t n_texels;
    uint64_t n_tris, n_verts;
    tri_t* tris; v3* vnormals;
    node_t* nodes; uint64_t n_nodes, cap_nodes;
    uint32_t* refs; uint64_t n_refs, cap_refs;
    uint32_t kd_max_depth, kd_max_leaf;
    uint64_t n_leaves, max_leaf_refs, tree_depth;
};

/* ------------------------------------------------------------------------------------------------------------
 * vec3 algebra - core/math/vec3.hpp
 * ---------------------------------------------------------------------------------------------------------- */
static inline v3 v3_add(v3 a, v3 b) { v3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }            /* :76-78 */
static inline v3 v3_sub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }            /* :80-82 */
static inline v3 v3_scale(float s, v3 a) { v3 r = {s * a.x, s * a.y, s * a.z}; return r; }             /* :94-97 */
static inline v3 v3_neg(v3 a) { v3 r = {-a.x, -a.y, -a.z}; return r; }                                 /* :43-45 */
static inline float v3_len2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }                        /* :84-86 L->R */
static inline float v3_len(v3 a) { return sqrtf(v3_len2(a)); }                                         /* :88-90 */
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }             /* :119-122 */
static inline v3 v3_cross(v3 a, v3 b) {                                                                 /* :124-131 */
    v3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
static inline v3 v3_normalized(v3 a) {                                                                  /* :104-108 */
    const float inv = 1.0f / v3_len(a);      /* true division of 1 by sqrt, not rsqrt */
    v3 r = {a.x * inv, a.y * inv, a.z * inv};
    return r;
}
static inline float v3_get(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

static inline ray make_ray(v3 o, v3 d) {                                                                /* ray3.hpp:11-14 */
    ray r; r.o = o; r.d = d;
    r.inv.x = 1.0f / d.x; r.inv.y = 1.0f / d.y; r.inv.z = 1.0f / d.z;                                   /* vec3.hpp:99-102 */
    return r;
}

static inline float fminf_std(float a, float b) { return (b < a) ? b : a; }   /* std::min(a,b) */
static inline float fmaxf_std(float a, float b) { return (a < b) ? b : a; }   /* std::max(a,b) */

/* ------------------------------------------------------------------------------------------------------------
 * RNG
 * ---------------------------------------------------------------------------------------------------------- */
static inline void mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
    const uint64_t p = (uint64_t)a * b; *hi = (uint32_t)(p >> 32); *lo = (uint32_t)p;
}
/* Philox4x32-10, Salmon et al. SC'11 (the published algorithm; counter-based, so the wavefront CUDA path and this
 * recursive restatement can draw the same numbers for the same (pixel, sample, path-node)). */
void ro_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c0, &hi0, &lo0);
        mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline float u01_from_u32(uint32_t x) { return (float)(x >> 8) * 0x1p-24f; }

#define RO_TAG_ROOT  0x52544230u /* "RTB0" */
#define RO_TAG_CHILD 0x4348494Cu /* "CHIL" */
#define RO_TAG_GI    0x47495F5Fu /* "GI__" */
enum { RO_SLOT_REFRACT = 0, RO_SLOT_REFLECT = 1, RO_SLOT_GI0 = 2 };

typedef struct {
    uint32_t mode;
    uint32_t minstd;      /* engine state, utils/rand.hpp:16 (std::minstd_rand seeded with fixed_rng_seed) */
} rng_t;

/* utils/rand.hpp:5-19: std::generate_canonical<float, 24>(std::minstd_rand).  libstdc++ 13 bits/random.tcc:
 * range r = 2147483646 -> one draw; sum = float(x - min); tmp = float(1 * r) = 2147483648.f; ret = sum / tmp,
 * nudged below 1 when it rounds to 1. */
static float minstd_urand01(rng_t* g) {
    g->minstd = (uint32_t)(((uint64_t)g->minstd * 48271u) % 2147483647u);
    const float sum = (float)(g->minstd - 1u);
    float ret = sum / 2147483648.0f;
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    return ret;
}

/* ------------------------------------------------------------------------------------------------------------
 * scene construction
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { const uint8_t* p; uint64_t n, off; int bad; } rd_t;
static void rd_get(rd_t* r, void* dst, uint64_t bytes) {
    if (r->off + bytes > r->n) { r->bad = 1; memset(dst, 0, bytes); return; }
    memcpy(dst, r->p + r->off, bytes); r->off += bytes;
}
static uint32_t rd_u32(rd_t* r) { uint32_t v; rd_get(r, &v, 4); return v; }

static void box_expand(float* bmin, float* bmax, v3 p) {                                               /* aabb3.hpp:25-32 */
    bmin[0] = fminf_std(bmin[0], p.x); bmin[1] = fminf_std(bmin[1], p.y); bmin[2] = fminf_std(bmin[2], p.z);
    bmax[0] = fmaxf_std(bmax[0], p.x); bmax[1] = fmaxf_std(bmax[1], p.y); bmax[2] = fmaxf_std(bmax[2], p.z);
}
static void box_init(float* bmin, float* bmax) {                                                       /* aabb3.hpp:21-23 */
    for (int i = 0; i < 3; ++i) { bmin[i] = FLT_MAX; bmax[i] = -FLT_MAX; }
}

static uint64_t push_node(ro_scene* s, uint64_t parent, const float* bmin, const float* bmax) {
    if (s->n_nodes == s->cap_nodes) {
        s->cap_nodes = s->cap_nodes ? s->cap_nodes * 2 : 256;
        s->nodes = (node_t*)realloc(s->nodes, s->cap_nodes * sizeof(node_t));
    }
    node_t* n = &s->nodes[s->n_nodes];
    n->parent = parent; n->child0 = n->child1 = n->first_ref = RO_EMPTY; n->ref_count = 0;
    memcpy(n->bmin, bmin, 12); memcpy(n->bmax, bmax, 12);
    return s->n_nodes++;
}

/* aabb3.hpp:68-72 - closed-interval box/box overlap, `this` = a, `other` = b */
static int box_overlap(const float* amin, const float* amax, const float* bmin, const float* bmax) {
    return (bmin[0] <= amax[0] && amin[0] <= bmax[0]) && (bmin[1] <= amax[1] && amin[1] <= bmax[1]) &&
           (bmin[2] <= amax[2] && amin[2] <= bmax[2]);
}

/* kd_tree_simd.hpp:146-185 (build_tree) + :117-144 (build_tree_leaf; the W-lane packing is a storage detail:
 * the leaf keeps its triangle list in order, padding lanes repeat the last triangle and can never win, see
 * leaf_intersect below) + aabb3.hpp:43-60 (split). */
static void build_tree(ro_scene* s, uint64_t parent_idx, uint64_t depth, const uint32_t* idx, uint64_t n) {
    if (depth > s->tree_depth) s->tree_depth = depth;
    if (depth == s->kd_max_depth || n <= s->kd_max_leaf) {
        if (s->n_refs + n > s->cap_refs) {
            s->cap_refs = (s->n_refs + n) * 2 + 64;
            s->refs = (uint32_t*)realloc(s->refs, s->cap_refs * sizeof(uint32_t));
        }
        memcpy(s->refs + s->n_refs, idx, n * sizeof(uint32_t));
        s->nodes[parent_idx].first_ref = s->n_refs;
        s->nodes[parent_idx].ref_count = n;
        s->n_refs += n;
        s->n_leaves++;
        if (n > s->max_leaf_refs) s->max_leaf_refs = n;
        return;
    }
    float min0[3], max0[3], min1[3], max1[3];
    memcpy(min0, s->nodes[parent_idx].bmin, 12); memcpy(max0, s->nodes[parent_idx].bmax, 12);
    memcpy(min1, min0, 12); memcpy(max1, max0, 12);
    uint32_t axis = (uint32_t)(depth % 3);
    /* :47-49 - a degenerate axis falls through to the next one (the reference recurses forever on a point box;
     * three tries is where we stop and split the zero-width slab anyway) */
    for (int tries = 0; tries < 3 && min0[axis] == max0[axis]; ++tries) axis = (axis + 1u) % 3u;
    const float mid = min0[axis] + ((max0[axis] - min0[axis]) / 2.0f);                                  /* :51 */
    max0[axis] = mid;                                                                                   /* :56 */
    min1[axis] = mid;                                                                                   /* :57 */

    uint32_t* c0 = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t* c1 = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    uint64_t n0 = 0, n1 = 0;
    for (uint64_t i = 0; i < n; ++i) {                                                                  /* :160-170 */
        const tri_t* t = &s->tris[idx[i]];
        if (box_overlap(min0, max0, t->bmin, t->bmax)) c0[n0++] = idx[i];
        if (box_overlap(min1, max1, t->bmin, t->bmax)) c1[n1++] = idx[i];
    }
    if (n0) {                                                                                           /* :172-177 */
        const uint64_t c = push_node(s, parent_idx, min0, max0);
        s->nodes[parent_idx].child0 = c;
        build_tree(s, c, depth + 1, c0, n0);
    }
    if (n1) {                                                                                           /* :179-184 */
        const uint64_t c = push_node(s, parent_idx, min1, max1);
        s->nodes[parent_idx].child1 = c;
        build_tree(s, c, depth + 1, c1, n1);
    }
    free(c0); free(c1);
}

ro_scene* ro_scene_from_rtsc_bytes(const void* bytes, uint64_t nbytes, uint32_t kd_max_depth, uint32_t kd_max_leaf) {
    rd_t r = {(const uint8_t*)bytes, nbytes, 0, 0};
    if (nbytes < 8 || memcmp(bytes, "RTSC", 4) != 0) return NULL;
    r.off = 4;
    if (rd_u32(&r) != 1) return NULL;
    ro_scene* s = (ro_scene*)calloc(1, sizeof(ro_scene));
    s->kd_max_depth = kd_max_depth; s->kd_max_leaf = kd_max_leaf;
    rd_get(&r, s->bg, 12);
    s->width = rd_u32(&r); s->height = rd_u32(&r); s->bucket = rd_u32(&r);
    rd_get(&r, s->cam_pos, 12); rd_get(&r, s->cam_m, 36);
    s->n_lights = rd_u32(&r);
    s->lights = (light_t*)calloc(s->n_lights ? s->n_lights : 1, sizeof(light_t));
    rd_get(&r, s->lights, (uint64_t)s->n_lights * sizeof(light_t));
    s->n_tex = rd_u32(&r);
    s->tex = (tex_t*)calloc(s->n_tex ? s->n_tex : 1, sizeof(tex_t));
    rd_get(&r, s->tex, (uint64_t)s->n_tex * sizeof(tex_t));
    s->n_mat = rd_u32(&r);
    s->mat = (mat_t*)calloc(s->n_mat ? s->n_mat : 1, sizeof(mat_t));
    rd_get(&r, s->mat, (uint64_t)s->n_mat * sizeof(mat_t));
    s->n_mesh = rd_u32(&r);
    uint32_t* heads = (uint32_t*)calloc(4 * (s->n_mesh ? s->n_mesh : 1), 4);
    rd_get(&r, heads, 16ull * s->n_mesh);
    s->mesh_mat = (uint32_t*)calloc(s->n_mesh ? s->n_mesh : 1, 4);
    for (uint32_t m = 0; m < s->n_mesh; ++m) {
        s->mesh_mat[m] = heads[4 * m]; s->n_verts += heads[4 * m + 1]; s->n_tris += heads[4 * m + 3];
    }
    if (r.bad) { ro_scene_free(s); free(heads); return NULL; }
    s->tris = (tri_t*)calloc(s->n_tris ? s->n_tris : 1, sizeof(tri_t));
    s->vnormals = (v3*)calloc(s->n_verts ? s->n_verts : 1, sizeof(v3));

    float root_min[3], root_max[3];
    box_init(root_min, root_max);
    uint64_t tri_base = 0, vert_base = 0;
    for (uint32_t m = 0; m < s->n_mesh; ++m) {
        const uint32_t nv = heads[4 * m + 1], nuv = heads[4 * m + 2], nt = heads[4 * m + 3];
        float* vb = (float*)malloc(12ull * (nv ? nv : 1));
        float* ub = (float*)malloc(8ull * (nuv ? nuv : 1));
        uint32_t* tb = (uint32_t*)malloc(12ull * (nt ? nt : 1));
        rd_get(&r, vb, 12ull * nv); rd_get(&r, ub, 8ull * nuv); rd_get(&r, tb, 12ull * nt);
        float mmin[3], mmax[3];
        box_init(mmin, mmax);
        for (uint32_t i = 0; i < nt; ++i) {
            tri_t* t = &s->tris[tri_base + i];
            const uint32_t i0 = tb[3 * i], i1 = tb[3 * i + 1], i2 = tb[3 * i + 2];
            const v3 v0 = {vb[3 * i0], vb[3 * i0 + 1], vb[3 * i0 + 2]};
            const v3 v1 = {vb[3 * i1], vb[3 * i1 + 1], vb[3 * i1 + 2]};
            const v3 v2 = {vb[3 * i2], vb[3 * i2 + 1], vb[3 * i2 + 2]};
            /* scene/primitive/triangle.hpp:20-30 */
            t->v0 = v0;
            t->normal = v3_normalized(v3_cross(v3_sub(v1, v0), v3_sub(v2, v0)));
            t->e1 = v3_sub(v1, v0);
            t->e2 = v3_sub(v2, v0);
            box_init(t->bmin, t->bmax);
            box_expand(t->bmin, t->bmax, v0); box_expand(t->bmin, t->bmax, v1); box_expand(t->bmin, t->bmax, v2);
            t->vi[0] = (uint32_t)(vert_base + i0); t->vi[1] = (uint32_t)(vert_base + i1); t->vi[2] = (uint32_t)(vert_base + i2);
            t->mesh = m;
            if (nuv) {                                                                                  /* io/json/loader.hpp:203-209 */
                t->uv[0] = ub[2 * i0]; t->uv[1] = ub[2 * i0 + 1]; t->uv[2] = ub[2 * i1]; t->uv[3] = ub[2 * i1 + 1];
                t->uv[4] = ub[2 * i2]; t->uv[5] = ub[2 * i2 + 1];
            }
            /* scene/object/mesh.hpp:27-39: mesh box + un-weighted vertex-normal accumulation, triangle order */
            box_expand(mmin, mmax, v0); box_expand(mmin, mmax, v1); box_expand(mmin, mmax, v2);
            const v3 tn = v3_normalized(v3_cross(v3_sub(v1, v0), v3_sub(v2, v0)));
            s->vnormals[vert_base + i0] = v3_add(s->vnormals[vert_base + i0], tn);
            s->vnormals[vert_base + i1] = v3_add(s->vnormals[vert_base + i1], tn);
            s->vnormals[vert_base + i2] = v3_add(s->vnormals[vert_base + i2], tn);
        }
        for (uint32_t i = 0; i < nv; ++i)                                                               /* mesh.hpp:41-43 */
            s->vnormals[vert_base + i] = v3_normalized(s->vnormals[vert_base + i]);
        /* kd_tree_simd.hpp:103-104: root_box.unite(mesh.box) - aabb3.hpp:34-41 */
        for (int a = 0; a < 3; ++a) { root_min[a] = fminf_std(root_min[a], mmin[a]); root_max[a] = fmaxf_std(root_max[a], mmax[a]); }
        tri_base += nt; vert_base += nv;
        free(vb); free(ub); free(tb);
    }
    free(heads);
    s->n_texels = rd_u32(&r);
    s->texels = (uint8_t*)malloc(s->n_texels ? s->n_texels : 1);
    rd_get(&r, s->texels, s->n_texels);
    if (r.bad || r.off != nbytes) { ro_scene_free(s); return NULL; }

    /* kd_tree_simd.hpp:100-115 */
    uint32_t* all = (uint32_t*)malloc((s->n_tris ? s->n_tris : 1) * sizeof(uint32_t));
    for (uint64_t i = 0; i < s->n_tris; ++i) all[i] = (uint32_t)i;
    push_node(s, RO_EMPTY, root_min, root_max);
    build_tree(s, 0, 0, all, s->n_tris);
    free(all);
    return s;
}

ro_scene* ro_scene_load_rtsc(const char* path, uint32_t kd_max_depth, uint32_t kd_max_leaf) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void* buf = malloc(n > 0 ? (size_t)n : 1);
    const size_t got = fread(buf, 1, (size_t)n, f);
    fclose(f);
    ro_scene* s = got == (size_t)n ? ro_scene_from_rtsc_bytes(buf, (uint64_t)n, kd_max_depth, kd_max_leaf) : NULL;
    free(buf);
    return s;
}

void ro_scene_free(ro_scene* s) {
    if (!s) return;
    free(s->lights); free(s->tex); free(s->mat); free(s->mesh_mat); free(s->texels); free(s->tris);
    free(s->vnormals); free(s->nodes); free(s->refs); free(s);
}

void ro_scene_info(const ro_scene* s, uint64_t* info) {
    info[0] = s->width; info[1] = s->height; info[2] = s->n_tris; info[3] = s->n_nodes; info[4] = s->n_refs;
    info[5] = s->n_leaves; info[6] = s->max_leaf_refs; info[7] = s->tree_depth;
}

void ro_tree(const ro_scene* s, uint64_t* node5, float* boxes, uint32_t* refs) {
    for (uint64_t i = 0; i < s->n_nodes; ++i) {
        const node_t* n = &s->nodes[i];
        if (node5) { node5[5 * i] = n->parent; node5[5 * i + 1] = n->child0; node5[5 * i + 2] = n->child1; node5[5 * i + 3] = n->first_ref; node5[5 * i + 4] = n->ref_count; }
        if (boxes) { memcpy(boxes + 6 * i, n->bmin, 12); memcpy(boxes + 6 * i + 3, n->bmax, 12); }
    }
    if (refs) memcpy(refs, s->refs, s->n_refs * sizeof(uint32_t));
}

void ro_geometry(const ro_scene* s, float* tri9, float* face_normals, float* vertex_normals, uint32_t* tri_vidx,
                 uint32_t* tri_mesh) {
    for (uint64_t i = 0; i < s->n_tris; ++i) {
        const tri_t* t = &s->tris[i];
        if (tri9) { memcpy(tri9 + 9 * i, &t->v0, 12); memcpy(tri9 + 9 * i + 3, &t->e1, 12); memcpy(tri9 + 9 * i + 6, &t->e2, 12); }
        if (face_normals) memcpy(face_normals + 3 * i, &t->normal, 12);
        if (tri_vidx) memcpy(tri_vidx + 3 * i, t->vi, 12);
        if (tri_mesh) tri_mesh[i] = t->mesh;
    }
    if (vertex_normals) memcpy(vertex_normals, s->vnormals, s->n_verts * sizeof(v3));
}

/* ------------------------------------------------------------------------------------------------------------
 * the hot path
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { int hit; float t, u, v; uint32_t tri; } cand_t;

/* core/math/aabb3.hpp:74-90 - slab test; only t_min is consumed by the caller (kd_tree_simd.hpp:203).
 * std::minmax(a,b) = (b<a) ? (b,a) : (a,b); std::max(a,b) = (a<b) ? b : a; std::min(a,b) = (b<a) ? b : a, so a
 * NaN (0*inf) leaves the running value alone.  Restated by value: the reference's :79 binds references to dead
 * temporaries (UB). */
static inline int slab_test(const node_t* n, const ray* r, float* t_min_out) {
    float t_min = 0.0f, t_max = FLT_MAX;
    for (int axis = 0; axis < 3; ++axis) {
        const float o = v3_get(r->o, axis), inv = v3_get(r->inv, axis);
        const float a = (n->bmin[axis] - o) * inv;
        const float b = (n->bmax[axis] - o) * inv;
        float t1, t2;
        if (b < a) { t1 = b; t2 = a; } else { t1 = a; t2 = b; }
        t_min = fmaxf_std(t_min, t1);
        t_max = fminf_std(t_max, t2);
        if (t_max < t_min) return 0;
    }
    *t_min_out = t_min;
    return 1;
}

/* kd_tree_simd.hpp:25-60 (triangle_packet::intersect, one lane) folded into :266-302 (intersect_leaf).
 * Lane-by-lane, in list order, with a strict `<` update: identical to the packet version because (i) inside a
 * pack the lowest lane among equal minima wins (:288-290), (ii) a later pack needs a strictly smaller t (:284),
 * (iii) padding lanes repeat the leaf's last triangle (:123) and therefore only ever tie with it. */
static inline void leaf_intersect(const ro_scene* s, const node_t* leaf, const ray* r, int cull, float eps,
                                  cand_t* best, uint64_t* n_tests) {
    const float dx = r->d.x, dy = r->d.y, dz = r->d.z;
    cand_t c; c.hit = 0; c.t = FLT_MAX; c.u = c.v = 0; c.tri = 0;
    for (uint64_t k = 0; k < leaf->ref_count; ++k) {
        const uint32_t id = s->refs[leaf->first_ref + k];
        const tri_t* T = &s->tris[id];
        const float pvx = dy * T->e2.z - dz * T->e2.y;                                                  /* :27 */
        const float pvy = dz * T->e2.x - dx * T->e2.z;                                                  /* :28 */
        const float pvz = dx * T->e2.y - dy * T->e2.x;                                                  /* :29 */
        const float det = T->e1.x * pvx + T->e1.y * pvy + T->e1.z * pvz;                                /* :31 L->R */
        int mask = cull ? (eps <= det) : (eps <= fabsf(det));                                           /* :33-38 */
        const float inv_det = 1.0f / det;                                                               /* :40 */
        const float tx = r->o.x - T->v0.x, ty = r->o.y - T->v0.y, tz = r->o.z - T->v0.z;                /* :42-44 */
        const float u = (tx * pvx + ty * pvy + tz * pvz) * inv_det;                                     /* :46 */
        mask &= (0.0f <= u) & (u <= 1.0f);                                                              /* :47 */
        const float qx = ty * T->e1.z - tz * T->e1.y;                                                   /* :49 */
        const float qy = tz * T->e1.x - tx * T->e1.z;                                                   /* :50 */
        const float qz = tx * T->e1.y - ty * T->e1.x;                                                   /* :51 */
        const float v = (dx * qx + dy * qy + dz * qz) * inv_det;                                        /* :53 */
        mask &= (0.0f <= v) & (u + v <= 1.0f);                                                          /* :54 */
        const float t = (T->e2.x * qx + T->e2.y * qy + T->e2.z * qz) * inv_det;                         /* :56 */
        mask &= (eps < t);                                                                              /* :57 */
        if (mask && t < c.t) { c.hit = 1; c.t = t; c.u = u; c.v = v; c.tri = id; }                      /* :276-298 */
    }
    *n_tests += leaf->ref_count;
    /* kd_tree_simd.hpp:216-226: a later leaf must be strictly closer */
    if (c.hit) {
        const float best_t = best->hit ? best->t : FLT_MAX;
        if (c.t < best_t) *best = c;
    }
}

/* kd_tree_simd.hpp:187-229 - traversal: LIFO stack, child0 pushed before child1 (so child1 is visited first),
 * prune when best_t < box.t_min (strict). */
static cand_t closest_hit(const ro_scene* s, const ray* r, int cull, float eps, uint64_t* counts) {
    cand_t best; best.hit = 0; best.t = FLT_MAX; best.u = best.v = 0; best.tri = 0;
    uint64_t stack[128];
    int sp = 0;
    uint64_t n_nodes = 0, n_tests = 0;
    stack[sp++] = 0;
    while (sp) {
        const node_t* n = &s->nodes[stack[--sp]];
        const float best_t = best.hit ? best.t : FLT_MAX;
        float t_min;
        ++n_nodes;
        if (!slab_test(n, r, &t_min) || best_t < t_min) continue;
        if (n->first_ref == RO_EMPTY) {
            if (n->child0 != RO_EMPTY) stack[sp++] = n->child0;
            if (n->child1 != RO_EMPTY) stack[sp++] = n->child1;
        } else {
            leaf_intersect(s, n, r, cull, eps, &best, &n_tests);
        }
    }
    if (counts) { counts[6] += n_nodes; counts[7] += n_tests; }
    return best;
}

/* what the reference's hit<F> carries (render/hit.hpp:9-21), assembled as kd_tree_simd.hpp:234-263 */
typedef struct {
    ray in;
    v3 position, hit_normal, face_normal;
    const float* uvs;
    float distance, u, v, w;
    uint32_t mesh, tri;
} hit_t;

static inline void assemble_hit(const ro_scene* s, const ray* r, const cand_t* c, hit_t* h) {
    const tri_t* T = &s->tris[c->tri];
    const float u = c->u, v = c->v;
    const float w = 1.0f - u - v;                                                                       /* :238 */
    const v3 n0 = s->vnormals[T->vi[0]], n1 = s->vnormals[T->vi[1]], n2 = s->vnormals[T->vi[2]];
    h->in = *r;
    h->hit_normal = v3_normalized(v3_add(v3_add(v3_scale(u, n1), v3_scale(v, n2)), v3_scale(w, n0)));   /* :250 */
    h->position = v3_add(r->o, v3_scale(c->t, r->d));                                                   /* :254 */
    h->face_normal = T->normal;
    h->uvs = T->uv;
    h->distance = c->t; h->u = u; h->v = v; h->w = w;
    h->mesh = T->mesh; h->tri = c->tri;
}

/* ------------------------------------------------------------------------------------------------------------
 * render loops (the callers)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    const ro_scene* s;
    const ro_params* p;
    rng_t rng;
    uint64_t counts[RO_N_COUNTS];
    ro_record* log; uint64_t log_n, log_cap;
} ctx_t;

static inline cand_t query(ctx_t* c, const ray* r, int cull, uint32_t kind) {
    const uint64_t nodes_before = c->counts[6], tests_before = c->counts[7];
    const cand_t h = closest_hit(c->s, r, cull, c->p->eps, c->counts);
    const int slot = kind == RO_KIND_PRIMARY ? 0 : (kind == RO_KIND_SHADOW ? 2 : 4);
    if (slot) { c->counts[6 + slot] += c->counts[6] - nodes_before; c->counts[7 + slot] += c->counts[7] - tests_before; }
    c->counts[slot]++;
    if (h.hit) c->counts[slot + 1]++;
    if (c->log && c->log_n < c->log_cap) {
        ro_record* q = &c->log[c->log_n++];
        q->o[0] = r->o.x; q->o[1] = r->o.y; q->o[2] = r->o.z; q->d[0] = r->d.x; q->d[1] = r->d.y; q->d[2] = r->d.z;
        q->t = h.hit ? h.t : 0; q->u = h.hit ? h.u : 0; q->v = h.hit ? h.v : 0;
        q->tri = h.hit ? (int32_t)h.tri : -1; q->cull = (uint32_t)cull; q->kind = kind;
    }
    return h;
}

/* render/render.hpp:110-131 */
static int is_occluded(ctx_t* c, ray r, float max_t) {
    const ro_scene* s = c->s;
    while (0.0f < max_t) {
        const cand_t h = query(c, &r, 0, RO_KIND_SHADOW);
        if (!h.hit || max_t < h.t) return 0;                                                            /* :117-119 */
        const mat_t* m = &s->mat[s->mesh_mat[s->tris[h.tri].mesh]];
        if (m->kind != 2) return 1;                                                                     /* :121-124, scene/material/queries.hpp:27-30 */
        const v3 pos = v3_add(r.o, v3_scale(h.t, r.d));                                                 /* hit.position, kd_tree_simd.hpp:254 */
        r.o = v3_add(pos, v3_scale(c->p->shadow_bias, r.d));                                            /* :126 (inv_direction unchanged) */
        max_t -= h.t;                                                                                   /* :127 */
    }
    return 0;
}

/* scene/texture/{albedo,edge,checker,bitmap}.hpp */
static col sample_texture(const ro_scene* s, const tex_t* t, const hit_t* h) {
    col r;
    const float hu = h->u, hv = h->v;
    const float hw = (float)(1. - (double)hu - (double)hv);            /* `1. - hit_u - hit_v` in double, narrowed */
    switch (t->kind) {
        case 0:                                                                                         /* albedo.hpp:11-13 */
            r.r = t->c0[0]; r.g = t->c0[1]; r.b = t->c0[2]; return r;
        case 1:                                                                                         /* edge.hpp:13-22 */
            if (hu < t->scalar || hv < t->scalar || hw < t->scalar) { r.r = t->c0[0]; r.g = t->c0[1]; r.b = t->c0[2]; }
            else { r.r = t->c1[0]; r.g = t->c1[1]; r.b = t->c1[2]; }
            return r;
        default: break;
    }
    /* final_uv = hit_w * uvs.x + hit_u * uvs.y + hit_v * uvs.z   (vec2 algebra, L->R) */
    const float fx = (hw * h->uvs[0] + hu * h->uvs[2]) + hv * h->uvs[4];
    const float fy = (hw * h->uvs[1] + hu * h->uvs[3]) + hv * h->uvs[5];
    if (t->kind == 2) {                                                                                 /* checker.hpp:12-26 */
        const int32_t u2 = (int32_t)(fx / t->scalar);
        const int32_t v2 = (int32_t)(fy / t->scalar);
        if ((u2 + v2) % 2 == 0) { r.r = t->c0[0]; r.g = t->c0[1]; r.b = t->c0[2]; }
        else { r.r = t->c1[0]; r.g = t->c1[1]; r.b = t->c1[2]; }
        return r;
    }
    /* bitmap.hpp:46-60: row in double, column in float; size_t conversion then clamp to [0, dim-1] */
    uint64_t row = (uint64_t)(int64_t)((1. - (double)fy) * (double)t->h);
    uint64_t colm = (uint64_t)(int64_t)(fx * (float)t->w);
    if (row > (uint64_t)t->h - 1) row = (uint64_t)t->h - 1;
    if (colm > (uint64_t)t->w - 1) colm = (uint64_t)t->w - 1;
    const uint8_t* px = s->texels + t->off + (row * t->w + colm) * 3;
    const float scale = (float)(1.0 / 255.0);                                                           /* bitmap.hpp:19 */
    r.r = (float)px[0] * scale; r.g = (float)px[1] * scale; r.b = (float)px[2] * scale;                 /* :27-29 */
    return r;
}

static const float RO_PI = 3.14159265358979323846f;  /* std::numbers::pi_v<float> */

/* direct lighting loop shared by the diffuse (:184-206) and texture (:213-236) branches */
static void add_direct(ctx_t* c, const hit_t* h, int smooth, col albedo, col* acc) {
    const ro_scene* s = c->s;
    for (uint32_t li = 0; li < s->n_lights; ++li) {
        const light_t* L = &s->lights[li];
        const v3 lp = {L->pos[0], L->pos[1], L->pos[2]};
        v3 ld = v3_sub(lp, h->position);
        const float radius = v3_len(ld);
        const float area = 4.0f * RO_PI * radius * radius;                                              /* L->R */
        ld = v3_normalized(ld);
        const float cosine = fmaxf_std(0.0f, v3_dot(ld, smooth ? h->hit_normal : h->face_normal));
        const ray sr = make_ray(v3_add(h->position, v3_scale(c->p->shadow_bias, ld)), ld);
        if (is_occluded(c, sr, radius)) continue;
        const float k = (L->intensity / area) * cosine;
        acc->r += k * albedo.r; acc->g += k * albedo.g; acc->b += k * albedo.b;
    }
}

/* render/render.hpp:133-308.  `key` is the Philox path-node key (unused with the reference's minstd sequence). */
static col color_hit(ctx_t* c, const hit_t* h, uint32_t depth, const uint32_t key[2]) {
    const ro_scene* s = c->s;
    const ro_params* p = c->p;
    col out = {0, 0, 0};
    if (depth == p->max_ray_depth) { out.r = s->bg[0]; out.g = s->bg[1]; out.b = s->bg[2]; return out; } /* :138-139 */
    const mat_t* m = &s->mat[s->mesh_mat[h->mesh]];
    const col albedo = {m->albedo[0], m->albedo[1], m->albedo[2]};
    const v3 I = h->in.d;

    switch (m->kind) {
    case 0: {                                                                                           /* diffuse :149-210 */
        for (uint32_t i = 0; i < p->gi_rays; ++i) {                                                     /* :151-182 */
            const v3 right = v3_normalized(v3_cross(I, h->hit_normal));
            const v3 up = h->hit_normal;
            const v3 fwd = v3_cross(right, up);
            float u1, u2;
            uint32_t ckey[2] = {0, 0};
            if (c->rng.mode == RO_RNG_MINSTD) {
                u1 = minstd_urand01(&c->rng);
                u2 = minstd_urand01(&c->rng);
            } else {
                const uint32_t ctr[4] = {i, 0, 0, RO_TAG_GI};
                uint32_t o[4];
                ro_philox4x32_10(ctr, key, o);
                u1 = u01_from_u32(o[0]); u2 = u01_from_u32(o[1]);
                const uint32_t cc[4] = {RO_SLOT_GI0 + i, 0, 0, RO_TAG_CHILD};
                ro_philox4x32_10(cc, key, o);
                ckey[0] = o[0]; ckey[1] = o[1];
            }
            const float a1 = RO_PI * u1;                                                                /* :160 */
            const v3 rv = {cosf(a1), sinf(a1), 0.0f};                                                   /* :161 */
            const float a2 = RO_PI * u2 * 2.0f;                                                         /* :163 */
            const float ca = cosf(a2), sa = sinf(a2);
            /* rotate_y_mat * rand_xy_vec, core/math/mat3.hpp:53-60 with the literal matrix of :164-168 */
            const v3 rot = {ca * rv.x + 0.0f * rv.y + (-sa) * rv.z,
                            0.0f * rv.x + 1.0f * rv.y + 0.0f * rv.z,
                            sa * rv.x + 0.0f * rv.y + ca * rv.z};
            const v3 org = v3_add(h->position, v3_scale(p->reflection_bias, h->hit_normal));            /* :172 */
            /* local_hit_mat rows = right, up, fwd (mat3.hpp:13-17) times rot */
            const v3 dir = {right.x * rot.x + right.y * rot.y + right.z * rot.z,
                            up.x * rot.x + up.y * rot.y + up.z * rot.z,
                            fwd.x * rot.x + fwd.y * rot.y + fwd.z * rot.z};
            const ray gr = make_ray(org, dir);
            const cand_t gh = query(c, &gr, 0, RO_KIND_GI);
            if (!gh.hit) continue;
            hit_t hh; assemble_hit(s, &gr, &gh, &hh);
            const col cc2 = color_hit(c, &hh, depth + 1, ckey);
            out.r += cc2.r; out.g += cc2.g; out.b += cc2.b;
        }
        add_direct(c, h, (int)m->smooth, albedo, &out);
        const float div = (float)(p->gi_rays + 1);                                                      /* :208 */
        out.r /= div; out.g /= div; out.b /= div;
        return out;
    }
    case 4: {                                                                                           /* texture :211-238 */
        hit_t tmp = *h;
        const col a = sample_texture(s, &s->tex[m->texture], &tmp);
        /* the reference samples inside the light loop (:234-235); the sample does not depend on the light */
        add_direct(c, h, (int)m->smooth, a, &out);
        return out;
    }
    case 1: {                                                                                           /* reflective :239-250 */
        const float k = 2.0f * v3_dot(I, h->hit_normal);
        const v3 rd = v3_sub(I, v3_scale(k, h->hit_normal));
        const ray rr = make_ray(v3_add(h->position, v3_scale(p->reflection_bias, rd)), rd);
        const cand_t rh = query(c, &rr, 0, RO_KIND_REFLECT);
        if (!rh.hit) { out.r = s->bg[0]; out.g = s->bg[1]; out.b = s->bg[2]; return out; }
        uint32_t ckey[2] = {0, 0};
        if (c->rng.mode == RO_RNG_PHILOX) {
            const uint32_t cc[4] = {RO_SLOT_REFLECT, 0, 0, RO_TAG_CHILD}; uint32_t o[4];
            ro_philox4x32_10(cc, key, o); ckey[0] = o[0]; ckey[1] = o[1];
        }
        hit_t hh; assemble_hit(s, &rr, &rh, &hh);
        return color_hit(c, &hh, depth + 1, ckey);
    }
    case 2: {                                                                                           /* refractive :251-301 */
        v3 n = v3_normalized(m->smooth ? h->hit_normal : h->face_normal);
        const v3 i = v3_normalized(I);
        float eta_i = 1.0f, eta_r = m->ior;
        if (0.0f < v3_dot(i, n)) { const float t = eta_i; eta_i = eta_r; eta_r = t; n = v3_neg(n); }    /* :258-261 */
        const float cos_i = -v3_dot(i, n);
        const float sin_i = sqrtf(1.0f - cos_i * cos_i);
        uint32_t key_refl[2] = {0, 0}, key_refr[2] = {0, 0};
        if (c->rng.mode == RO_RNG_PHILOX) {
            uint32_t o[4];
            const uint32_t c0[4] = {RO_SLOT_REFRACT, 0, 0, RO_TAG_CHILD};
            ro_philox4x32_10(c0, key, o); key_refr[0] = o[0]; key_refr[1] = o[1];
            const uint32_t c1[4] = {RO_SLOT_REFLECT, 0, 0, RO_TAG_CHILD};
            ro_philox4x32_10(c1, key, o); key_refl[0] = o[0]; key_refl[1] = o[1];
        }
        const float k2 = 2.0f * v3_dot(i, n);
        const v3 refl_d = v3_sub(i, v3_scale(k2, n));                                                   /* :267, :291 */
        if (eta_r / eta_i < sin_i) {                                                                    /* :266 total internal reflection */
            const ray rr = make_ray(v3_add(h->position, v3_scale(p->reflection_bias, refl_d)), refl_d);
            const cand_t rh = query(c, &rr, 0, RO_KIND_REFLECT);
            if (!rh.hit) return out;                                                                    /* black, :271-273 */
            hit_t hh; assemble_hit(s, &rr, &rh, &hh);
            return color_hit(c, &hh, depth + 1, key_refl);
        }
        const float sin_r = (sin_i * eta_i) / eta_r;                                                    /* :278 */
        const float cos_r = sqrtf(1.0f - sin_r * sin_r);
        const v3 tang = v3_normalized(v3_add(i, v3_scale(cos_i, n)));
        const v3 rdir = v3_add(v3_scale(cos_r, v3_neg(n)), v3_scale(sin_r, tang));                      /* :281 */
        col refr = {0, 0, 0}, refl = {0, 0, 0};
        {
            const ray rr = make_ray(v3_add(h->position, v3_scale(p->refraction_bias, rdir)), rdir);     /* :283 */
            const cand_t rh = query(c, &rr, 0, RO_KIND_REFRACT);
            if (rh.hit) { hit_t hh; assemble_hit(s, &rr, &rh, &hh); refr = color_hit(c, &hh, depth + 1, key_refr); }
        }
        {
            const ray rr = make_ray(v3_add(h->position, v3_scale(p->reflection_bias, refl_d)), refl_d); /* :292 */
            const cand_t rh = query(c, &rr, 0, RO_KIND_REFLECT);
            if (rh.hit) { hit_t hh; assemble_hit(s, &rr, &rh, &hh); refl = color_hit(c, &hh, depth + 1, key_refl); }
        }
        /* :300 - std::pow(float, int) promotes to double; 0.5 is a double literal; result narrowed to F */
        const float fresnel = (float)(0.5 * pow((double)(1.0f + v3_dot(i, n)), 5.0));
        const float omf = 1.0f - fresnel;
        out.r = fresnel * refl.r + omf * refr.r;                                                        /* :301 */
        out.g = fresnel * refl.g + omf * refr.g;
        out.b = fresnel * refl.b + omf * refr.b;
        return out;
    }
    case 3:                                                                                             /* constant :302-303 */
        return albedo;
    default:
        return out;
    }
}

/* render/render.hpp:35-62 - camera ray for raster position (rx, ry) */
static ray camera_ray(const ro_scene* s, const ro_params* p, float raster_x, float raster_y) {
    const float aspect = (float)s->width / (float)s->height;                                            /* :27 */
    const float ndc_x = raster_x / (float)s->width;                                                     /* :47 */
    const float ndc_y = raster_y / (float)s->height;                                                    /* :48 */
    float sx = (2.0f * ndc_x) - 1.0f;                                                                   /* :50 */
    float sy = 1.0f - (2.0f * ndc_y);                                                                   /* :51 */
    sx *= aspect;                                                                                       /* :53 */
    const double fov_rad = p->fov_degrees * (3.14159265358979323846 / 180.);                            /* utils/convert.hpp:3-6, F = double */
    const double th = tan(fov_rad / (double)2.0f);                                                      /* :56-57 */
    sx = (float)((double)sx * th);
    sy = (float)((double)sy * th);
    const float* m = s->cam_m;
    /* transpose(camera.matrix) * (sx, sy, -1): mat3.hpp:25-31, :53-60 */
    const v3 d = {m[0] * sx + m[3] * sy + m[6] * -1.0f,
                  m[1] * sx + m[4] * sy + m[7] * -1.0f,
                  m[2] * sx + m[5] * sy + m[8] * -1.0f};
    const v3 o = {s->cam_pos[0], s->cam_pos[1], s->cam_pos[2]};
    return make_ray(o, v3_normalized(d));
}

static uint32_t spp_total_of(const ro_params* p) { return p->spp_total ? p->spp_total : p->spp; }

/* one pixel: render/render.hpp:33-76 */
static void render_pixel(ctx_t* c, uint32_t x, uint32_t y, float* rgb) {
    const ro_scene* s = c->s;
    const ro_params* p = c->p;
    const uint32_t spp_total = spp_total_of(p);
    col sum = {0, 0, 0};
    for (uint32_t si = 0; si < p->spp; ++si) {
        float rx = (float)x, ry = (float)y;
        uint32_t key[2] = {0, 0};
        if (c->rng.mode == RO_RNG_PHILOX) {
            const uint32_t ctr[4] = {y * s->width + x, p->sample_offset + si, 0, RO_TAG_ROOT};
            const uint32_t k0[2] = {p->seed, 0};
            uint32_t o[4];
            ro_philox4x32_10(ctr, k0, o);
            key[0] = o[0]; key[1] = o[1];
            if (spp_total == 1) { rx += 0.5f; ry += 0.5f; }
            else { rx += u01_from_u32(o[2]); ry += u01_from_u32(o[3]); }
        } else {
            if (spp_total == 1) { rx += 0.5f; ry += 0.5f; }                                             /* :39-42 */
            else { rx += minstd_urand01(&c->rng); ry += minstd_urand01(&c->rng); }                      /* :43-44 */
        }
        const ray r = camera_ray(s, p, rx, ry);
        const cand_t h = query(c, &r, 1, RO_KIND_PRIMARY);                                              /* :64 culling ON */
        if (h.hit) {
            hit_t hh; assemble_hit(s, &r, &h, &hh);
            const col cc = color_hit(c, &hh, 0, key);
            sum.r += cc.r; sum.g += cc.g; sum.b += cc.b;                                                /* :66 */
        } else {
            sum.r += s->bg[0]; sum.g += s->bg[1]; sum.b += s->bg[2];                                    /* :68 */
        }
    }
    if (!p->raw_sum) {
        const float div = (float)spp_total;                                                             /* :72 */
        sum.r /= div; sum.g /= div; sum.b /= div;
    }
    float* o = rgb + ((uint64_t)y * s->width + x) * 3;
    o[0] = sum.r; o[1] = sum.g; o[2] = sum.b;
}

void ro_default_params(ro_params* p) {
    memset(p, 0, sizeof *p);
    p->fov_degrees = 90.;
    p->eps = (float)1e-6;
    p->shadow_bias = (float)1e-4; p->reflection_bias = (float)1e-4; p->refraction_bias = (float)1e-4;
    p->spp = 1; p->max_ray_depth = 5; p->gi_rays = 0; p->seed = 42; p->rng = RO_RNG_PHILOX;
}

typedef struct {
    ctx_t c;
    uint32_t x0, y0, x1, y1;
    float* rgb;
    volatile uint32_t* next_row;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (;;) {
        const uint32_t y = __atomic_fetch_add(j->next_row, 1, __ATOMIC_RELAXED);
        if (y >= j->y1) break;
        for (uint32_t x = j->x0; x < j->x1; ++x) render_pixel(&j->c, x, y, j->rgb);
    }
    return NULL;
}

void ro_render(const ro_scene* s, const ro_params* p, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
               float* rgb, int n_threads, uint64_t* counts) {
    if (x1 > s->width) x1 = s->width;
    if (y1 > s->height) y1 = s->height;
    if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (p->rng == RO_RNG_MINSTD || n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    volatile uint32_t next_row = y0;
    job_t* jobs = (job_t*)calloc((size_t)n_threads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 0; i < n_threads; ++i) {
        jobs[i].c.s = s; jobs[i].c.p = p; jobs[i].c.rng.mode = p->rng; jobs[i].c.rng.minstd = p->seed;
        jobs[i].x0 = x0; jobs[i].y0 = y0; jobs[i].x1 = x1; jobs[i].y1 = y1; jobs[i].rgb = rgb; jobs[i].next_row = &next_row;
    }
    if (n_threads == 1) worker(&jobs[0]);
    else {
        for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, worker, &jobs[i]);
        for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    }
    if (counts) {
        memset(counts, 0, sizeof(uint64_t) * RO_N_COUNTS);
        for (int i = 0; i < n_threads; ++i) for (int k = 0; k < RO_N_COUNTS; ++k) counts[k] += jobs[i].c.counts[k];
    }
    free(jobs); free(th);
}

uint64_t ro_record_frame(const ro_scene* s, const ro_params* p, ro_record* out, uint64_t cap, float* rgb) {
    ctx_t c; memset(&c, 0, sizeof c);
    c.s = s; c.p = p; c.rng.mode = p->rng; c.rng.minstd = p->seed; c.log = out; c.log_cap = cap;
    for (uint32_t y = 0; y < s->height; ++y)
        for (uint32_t x = 0; x < s->width; ++x) render_pixel(&c, x, y, rgb);
    return c.log_n;
}

void ro_trace(const ro_scene* s, float eps, const float* rays6, uint64_t n, int cull, float* tuv, int32_t* tri,
              uint64_t* counts) {
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = rays6 + 6 * i;
        const v3 o = {q[0], q[1], q[2]}, d = {q[3], q[4], q[5]};
        const ray r = make_ray(o, d);
        const cand_t h = closest_hit(s, &r, cull, eps, counts);
        tuv[3 * i] = h.hit ? h.t : 0; tuv[3 * i + 1] = h.hit ? h.u : 0; tuv[3 * i + 2] = h.hit ? h.v : 0;
        tri[i] = h.hit ? (int32_t)h.tri : -1;
    }
}

void ro_occluded(const ro_scene* s, const ro_params* p, const float* rays6, const float* max_t, uint64_t n,
                 uint8_t* out, uint64_t* counts) {
    ctx_t c; memset(&c, 0, sizeof c);
    c.s = s; c.p = p;
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = rays6 + 6 * i;
        const v3 o = {q[0], q[1], q[2]}, d = {q[3], q[4], q[5]};
        out[i] = (uint8_t)is_occluded(&c, make_ray(o, d), max_t[i]);
    }
    if (counts) memcpy(counts, c.counts, sizeof c.counts);
}

void ro_primary_rays(const ro_scene* s, const ro_params* p, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                     float* rays6) {
    uint64_t k = 0;
    for (uint32_t y = y0; y < y1; ++y)
        for (uint32_t x = x0; x < x1; ++x) {
            const ray r = camera_ray(s, p, (float)x + 0.5f, (float)y + 0.5f);
            float* q = rays6 + 6 * k++;
            q[0] = r.o.x; q[1] = r.o.y; q[2] = r.o.z; q[3] = r.d.x; q[4] = r.d.y; q[5] = r.d.z;
        }
}
