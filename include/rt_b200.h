/* rt_b200.h - C ABI of the B200-native backend for simd-raytracer's kd_tree_simd_accel hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch types.  The reference has no FFI of its
 * own - its plugin point is the C++20 concept `accelerator<A,F>` (/root/reference/include/raytracer/render/accel/
 * accel.hpp:8-12) plus the implicit `scene_ptr` member (render/render.hpp:21,113,136) - so each entry point below
 * names the reference interface it stands in for.  include/b200_accel.hpp is the header-only C++ adapter that
 * satisfies that concept on top of this ABI; INTEGRATION.md shows the reference-side binding.
 *
 * Every function returns an rt_status (0 = ok) and never throws across the boundary.  The reference's query is
 * `noexcept` and reports a miss as std::nullopt (kd_tree_simd.hpp:188,230-232); here a miss is tri == -1.
 *
 * There is no CPU fallback: compute entry points fail with RT_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 3

#if defined(__GNUC__)
#define RT_API __attribute__((visibility("default")))
#else
#define RT_API
#endif

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_BAD_ARG = 1,
    RT_ERR_NO_DEVICE = 2,      /* no CUDA device / not an sm_100 part / host-only scene used for compute */
    RT_ERR_CUDA = 3,           /* a CUDA runtime call failed; see rt_last_error() */
    RT_ERR_IO = 4,             /* scene / texture file could not be read */
    RT_ERR_PARSE = 5,          /* malformed .crtscene / RTSC; unknown material or texture type */
    RT_ERR_OOM = 6,
    RT_ERR_UNSUPPORTED = 7,    /* e.g. a bitmap format the loader cannot decode */
    RT_FRAME_RERENDERED = 8,   /* rt_frame_wait on an rt_render_frame_device_begin ticket: not an error, see there */
    RT_ERR_TIMEOUT = 9         /* multi-GPU combine: a peer did not signal a frame within RT_B200_PEER_TIMEOUT_MS (default 30 s;
                                  0 = wait for ever).  The waiting kernels give up instead of spinning on a dead peer, the stream
                                  drains, and every later rt_peer_* call of the group reports this status */
} rt_status;

/* ---- scene description: what io/json/loader.hpp:235-265 produces, flattened -------------------------------- */

enum { RT_TEX_ALBEDO = 0, RT_TEX_EDGES = 1, RT_TEX_CHECKER = 2, RT_TEX_BITMAP = 3 };      /* scene/texture/texture.hpp:11 */
enum { RT_MAT_DIFFUSE = 0, RT_MAT_REFLECTIVE = 1, RT_MAT_REFRACTIVE = 2, RT_MAT_CONSTANT = 3,
       RT_MAT_TEXTURE = 4 };                                                              /* scene/material/material.hpp:11-12 */

typedef struct rt_light_desc { float position[3]; float intensity; } rt_light_desc;       /* scene/light.hpp:5-9 */

typedef struct rt_texture_desc {     /* scene/texture/{albedo,edge,checker,bitmap}.hpp */
    uint32_t kind;
    float c0[3];                     /* albedo | edge_color  | color_A */
    float c1[3];                     /*        | inner_color | color_B */
    float scalar;                    /*        | edge_width  | square_size */
    uint32_t bmp_w, bmp_h, bmp_off;  /* bitmap: size and byte offset of its RGB8 texels in rt_scene_desc.texels */
} rt_texture_desc;

typedef struct rt_material_desc {    /* scene/material/ headers */
    uint32_t kind;
    float albedo[3];
    float ior;
    uint32_t smooth_shading;
    int32_t texture;                 /* RT_MAT_TEXTURE: index into textures[] (the reference keys by name) */
} rt_material_desc;

typedef struct rt_mesh_desc {        /* scene/object/mesh.hpp:14-21 as loaded by loader.hpp:149-233 */
    uint32_t material;
    uint32_t n_vertices, n_uvs, n_triangles;
    const float* vertices;           /* 3 * n_vertices */
    const float* uvs;                /* 2 * n_uvs (may be null when n_uvs == 0) */
    const uint32_t* triangles;       /* 3 * n_triangles vertex indices */
} rt_mesh_desc;

typedef struct rt_scene_desc {       /* scene/scene.hpp:14-22 */
    float background[3];
    uint32_t width, height, bucket_size;   /* scene/settings.hpp:7-13 */
    float camera_position[3];
    float camera_matrix[9];                /* row major, used transposed (render/render.hpp:60) */
    uint32_t n_lights;    const rt_light_desc* lights;
    uint32_t n_textures;  const rt_texture_desc* textures;
    uint32_t n_materials; const rt_material_desc* materials;
    uint32_t n_meshes;    const rt_mesh_desc* meshes;
    uint64_t n_texel_bytes; const uint8_t* texels;
} rt_scene_desc;

/* The accel's template arguments (kd_tree_simd.hpp:63-67) become run-time options. */
typedef struct rt_build_opts {
    uint32_t kd_max_depth;           /* default 8  (kd_tree_simd.hpp:65) */
    uint32_t kd_max_leaf_size;       /* default 64 (kd_tree_simd.hpp:66) */
    int32_t device;                  /* CUDA ordinal; RT_DEVICE_HOST_ONLY builds tree + layout without a GPU */
    /* width of the bounding-volume hierarchy the accelerated mode walks: 2 = 64-byte two-child nodes (csrc/rt_bvh.cuh), 4 = their
     * four-wide collapse, 128-byte nodes (csrc/rt_bvh4.cuh; the reference author's own TODO, README.md:118-124); 0 = the default
     * (RT_DEFAULT_ACCEL_WIDTH).  Frames, hits and ray counts do not depend on it. */
    uint32_t accel_width;
    /* where the backend's own hierarchy is built: RT_ACCEL_BUILD_HOST (0, the default) = binned SAH on the host threads
     * (host/bvh_build.cpp; seconds at 10 M triangles, the cheapest tree to walk); RT_ACCEL_BUILD_DEVICE (1) = a linear BVH built by
     * CUDA kernels (csrc/rt_lbvh.cuh; tens of milliseconds at 10 M triangles, more node visits per ray), concurrently with the host
     * build of the reference's kd-tree.  Needs a device and more than 16 triangles; a scene whose linear hierarchy comes out
     * deeper than the traversal stacks allow is built on the host instead (rt_scene_info.accel_build tells which ran).  Frames,
     * hits and ray counts do not depend on it. */
    uint32_t accel_build;
} rt_build_opts;
#define RT_DEFAULT_ACCEL_WIDTH 4
#define RT_ACCEL_BUILD_HOST   0u
#define RT_ACCEL_BUILD_DEVICE 1u
#define RT_DEVICE_HOST_ONLY (-1)

/* config.hpp:6-17 as run-time parameters, plus the tile / sample slice used for multi-GPU sharding. */
typedef struct rt_params {
    double fov_degrees;              /* config.hpp:6 */
    float epsilon;                   /* config.hpp:8 (narrowed to float as src/main.cpp:37); must be >= FLT_MIN (RT_ERR_BAD_ARG otherwise) */
    float shadow_bias;               /* config.hpp:9  */
    float reflection_bias;           /* config.hpp:10 */
    float refraction_bias;           /* config.hpp:11 */
    uint32_t samples_per_pixel;      /* config.hpp:13 - samples rendered by THIS call */
    uint32_t max_ray_depth;          /* config.hpp:14 */
    uint32_t diffuse_reflection_ray_count; /* config.hpp:15 */
    uint32_t seed;                   /* config.hpp:17 fixed_rng_seed -> Philox key */
    uint32_t sample_offset;          /* first global sample index of this slice */
    uint32_t spp_total;              /* samples of the whole frame (0 = samples_per_pixel); ==1 -> pixel centres */
    uint32_t x0, y0, x1, y1;         /* tile rectangle; x1 == 0 / y1 == 0 mean full width / height */
    uint32_t flags;                  /* RT_FLAG_* */
    /* row bands - the reference's bucket decomposition (render/tile/bucket.hpp:7-21) for tile-sharded multi-GPU frames, in ONE
     * call per rank: with band_rows != 0 the call renders the rows y of the frame with (y / band_rows) % band_period ==
     * band_phase (rank r of N: band_period = N, band_phase = r) and leaves every other row of the device frame untouched.
     * band_rows must be a multiple of 4; x0..y1 must be 0; only the device-frame entry points take bands (rt_render_frame_device,
     * rt_render_frame_device_begin) - RT_ERR_BAD_ARG otherwise.  Pixels are the same bits as in a whole-frame render. */
    uint32_t band_rows, band_period, band_phase;
} rt_params;

#define RT_FLAG_RAW_SUM       0x1u   /* write the slice's sample sum; the caller divides after combining ranks */
#define RT_FLAG_FAST_MATH     0x2u   /* FMA-contracted traversal + intersection (not bit-exact) */
#define RT_FLAG_ORDERED       0x4u   /* accelerated query: near-child-first traversal of the backend's OWN bounding-volume hierarchy
                                        (binned SAH, <= 4 triangles per leaf; two-wide 64-byte or four-wide 128-byte nodes, see
                                        rt_build_opts.accel_width) instead of the reference's kd-tree in the reference's visit
                                        order.  Same hits bit for bit: every triangle test is the reference's arithmetic, and a
                                        query with an exact-t tie between two triangles is re-run in reference order.  This is a
                                        stated deviation from "8-byte kd nodes, short stack in registers" (DESIGN.md section 0):
                                        the reference's own tree is still flattened to 8-byte nodes (rt_scene_get_device_layout)
                                        and walked by flags = 0 */

typedef struct rt_hit {              /* the part of hit<F> (render/hit.hpp:9-21) that cannot be recomputed */
    float t, u, v;
    int32_t tri;                     /* global triangle index (mesh-order concatenation, kd_tree_simd.hpp:103-111); -1 = miss */
} rt_hit;

typedef struct rt_scene_info {
    uint32_t width, height;
    uint64_t n_triangles, n_vertices, n_nodes, n_leaves, n_leaf_refs, n_packets;
    uint64_t max_leaf_refs, tree_depth;
    uint64_t device_bytes;           /* resident scene bytes in HBM */
    double build_seconds, flatten_seconds, upload_seconds;
    int32_t device;
    uint32_t bvh_leaf_size;          /* most triangles per leaf of the backend's hierarchy: 4, or 1 for scenes far beyond L2 */
    uint64_t bvh_n_nodes, bvh_n_refs, bvh_n_leaves, bvh_depth;   /* the bounding-volume hierarchy (64-byte two-child nodes) */
    uint32_t accel_width, accel_build;                           /* 2 or 4: what the accelerated mode walks; RT_ACCEL_BUILD_* that built it */
    uint64_t bvh4_n_nodes, bvh4_stack_need;                      /* four-wide collapse: nodes, worst-case traversal stack entries */
    double accel_build_seconds;                                  /* the backend's hierarchy alone: build + flatten + collapse (device build: + its uploads) */
} rt_scene_info;

typedef struct rt_counters {         /* of the last rt_render_frame* call */
    uint64_t primary, primary_hits;
    uint64_t shadow, shadow_hits;            /* closest-hit queries issued by the is_occluded loops */
    uint64_t secondary, secondary_hits;      /* reflection + refraction + GI */
    uint64_t nodes_pool, shadow_pool;        /* wavefront pool high-water marks */
    uint32_t kernel_launches;
    uint32_t passes;
    float ms_total;                  /* device time of the frame, CUDA events */
    float ms_primary, ms_secondary, ms_shadow, ms_shade, ms_resolve;   /* by kernel class */
} rt_counters;

typedef struct rt_scene rt_scene;

RT_API int rt_abi_version(void);
RT_API const char* rt_status_string(int status);
RT_API const char* rt_last_error(void);                         /* thread-local detail for the last non-zero status */

RT_API void rt_default_build_opts(rt_build_opts* o);
RT_API void rt_default_params(rt_params* p);                    /* the values of config.hpp:6-17 */

/* kd_tree_simd_accel ctor (kd_tree_simd.hpp:100-115): host kd-tree build, flatten, upload. */
RT_API int rt_scene_create(const rt_scene_desc* desc, const rt_build_opts* opts, rt_scene** out);
/* parse_scene_file (io/json/loader.hpp:235-265) + the ctor above.  asset_root resolves relative bitmap paths
 * (the reference resolves them against the CWD, README.md:32-35); may be null. */
RT_API int rt_scene_create_from_crtscene(const char* path, const char* asset_root, const rt_build_opts* opts, rt_scene** out);
/* same, from the flat RTSC container used by the test fixtures (tests/helpers/crtscene.py documents the layout) */
RT_API int rt_scene_create_from_rtsc(const void* bytes, uint64_t n_bytes, const rt_build_opts* opts, rt_scene** out);
/* the scene as loaded (io/json/loader.hpp:235-265 semantics: every number float(double), 3-component uvs cut to 2, bitmaps decoded
 * to RGB8) in the flat RTSC container; *n_bytes = its size; buf may be null to ask for the size */
RT_API int rt_scene_export_rtsc(const rt_scene* s, void* buf, uint64_t cap, uint64_t* n_bytes);
RT_API void rt_scene_destroy(rt_scene* s);
RT_API int rt_scene_get_info(const rt_scene* s, rt_scene_info* info);

/* Host-side structures for builder / flattener parity checks (no GPU needed).  Null pointers are skipped.
 *   node5  : 5 x u64 per node = parent, child0, child1, first_ref, ref_count (UINT64_MAX = none), as
 *            kd_tree_simd_accel::node (kd_tree_simd.hpp:75-84) with packs counted in triangle refs
 *   boxes  : 6 floats per node (min xyz, max xyz)
 *   refs   : n_leaf_refs triangle indices (leaf lists, in order)                                              */
RT_API int rt_scene_get_tree(const rt_scene* s, uint64_t* node5, float* boxes, uint32_t* refs);
/*   nodes8 : the 8-byte device nodes (2 x u32 per node); packets: 40 x u32 per 4-triangle SoA packet          */
RT_API int rt_scene_get_device_layout(const rt_scene* s, uint32_t* nodes8, uint32_t* packets);
/* the bounding-volume hierarchy (rt_scene_info.bvh_*): nodes16 = 16 x u32 per inner node
 * { c0.min.xyz, c0.max.xyz, c1.min.xyz, c1.max.xyz, ref0, ref1, cnt0, cnt1 }; tris12 as above, one record per triangle   */
RT_API int rt_scene_get_bvh_layout(const rt_scene* s, uint32_t* nodes16, uint32_t* tris12, float* root6);
/*   tri9 = v0,e1,e2 per triangle; face normals; vertex normals (mesh-concatenated vertex order)               */
RT_API int rt_scene_get_geometry(const rt_scene* s, float* tri9, float* face_normals, float* vertex_normals);

/* accel.intersect<cull>(ray) for a batch (kd_tree_simd.hpp:187-264; callers render.hpp:64,116,175,244,269,284,293).
 * rays: 6 floats per ray (origin, direction); inv_direction is derived as ray3's ctor does (ray3.hpp:11-14).
 * Host buffers: copies in, traces, copies out.  flags: RT_FLAG_FAST_MATH / RT_FLAG_ORDERED or 0.              */
RT_API int rt_trace_closest(rt_scene* s, const float* rays, uint64_t n, int backface_culling, float epsilon,
                     uint32_t flags, rt_hit* hits);
/* is_occluded(accel, ray, max_t) for a batch (render/render.hpp:110-131), incl. refractive pass-through.       */
RT_API int rt_trace_occluded(rt_scene* s, const float* rays, const float* max_t, uint64_t n, float epsilon,
                      float shadow_bias, uint32_t flags, uint8_t* occluded);
/* device-pointer variants (no copies; `stream` is a cudaStream_t or null for the scene's own stream)           */
RT_API int rt_trace_closest_device(rt_scene* s, const float* d_rays, uint64_t n, int backface_culling, float epsilon,
                            uint32_t flags, rt_hit* d_hits, void* stream);
RT_API int rt_trace_occluded_device(rt_scene* s, const float* d_rays, const float* d_max_t, uint64_t n, float epsilon,
                             float shadow_bias, uint32_t flags, uint8_t* d_occluded, void* stream);

/* render_frame<A,F>(accel, schedule) (render/render.hpp:18-108): ray generation, traversal, shading, shadows,
 * reflection / refraction / GI bounces, all on device.  rgb: height*width*3 floats, row major (image<F>);
 * only the tile rectangle is written.                                                                         */
RT_API int rt_render_frame(rt_scene* s, const rt_params* p, float* rgb);
/* same + the PPM writer's quantisation uint8(255.999*clamp(c,0,1)) (io/image/ppm.hpp:17-19) fused on device.   */
RT_API int rt_render_frame_rgb8(rt_scene* s, const rt_params* p, uint8_t* rgb8);
/* Frame sequences - what a caller of render_frame in a loop does (the reference's animation outputs, the mp4 under outputs/: one
 * render_frame per camera pose / parameter set, src/main.cpp:13-25).  rt_render_frame_begin renders the frame into one of two
 * device frames and queues its download into `rgb` on a copy stream, so frame i's PCIe transfer overlaps frame i+1's render;
 * rt_frame_wait blocks until `rgb` of that ticket is complete.  `rgb` should be pinned for the overlap to take place and must
 * stay valid until waited for.  At most two frames are in flight (a third rt_*_begin waits for the older one); they render on
 * two internal streams into two sets of wavefront pools, so one frame's last, thinning launches run under the next frame's
 * first ones - a frame of a sequence costs less than a frame alone - and nothing orders their completion
 * but rt_frame_wait.                                                                                              */
RT_API int rt_render_frame_begin(rt_scene* s, const rt_params* p, float* rgb, uint64_t* ticket);
RT_API int rt_frame_wait(rt_scene* s, uint64_t ticket);
/* The same sequence with the PPM writer's quantisation (io/image/ppm.hpp:17-19) fused on the device: the frame that crosses PCIe
 * is height*width*3 BYTES - what write_ppm consumes (src/main.cpp:23) - a quarter of the float frame.                          */
RT_API int rt_render_frame_rgb8_begin(rt_scene* s, const rt_params* p, uint8_t* rgb8, uint64_t* ticket);
/* page-locked host memory for the frames of a sequence, for hosts that do not link the CUDA runtime themselves */
RT_API void* rt_alloc_pinned(uint64_t bytes);           /* null on failure */
RT_API void rt_free_pinned(void* p);
/* The same queued render into the CALLER's device frame (no download), asynchronous on `stream`.  All frames of a scene that are
 * in flight must have been queued on the same stream (RT_ERR_BAD_ARG otherwise): frames on a caller's stream are ordered by it
 * and share one set of pools; null = the scene's own streams, as above.  Work queued behind the frame on `stream` (a peer
 * combine, a copy) runs without a host round trip.  rt_frame_wait then returns RT_OK, or RT_FRAME_RERENDERED when the queued attempt outgrew the wavefront pools (first
 * frames of a scene): d_rgb has been rendered again and is correct now, but whatever consumed it before must be redone.      */
RT_API int rt_render_frame_device_begin(rt_scene* s, const rt_params* p, float* d_rgb, void* stream, uint64_t* ticket);
/* device framebuffer (for the multi-GPU combine): d_rgb = height*width*3 floats in HBM; asynchronous on stream */
RT_API int rt_render_frame_device(rt_scene* s, const rt_params* p, float* d_rgb, void* stream);
/* primary rays only: ray generation + intersect<true> (render.hpp:35-64); hits: one per pixel of the tile,
 * row major over the tile.  Host buffer.                                                                      */
RT_API int rt_trace_primary(rt_scene* s, const rt_params* p, rt_hit* hits);

RT_API int rt_get_counters(rt_scene* s, rt_counters* c);        /* blocks until the last frame has finished */

/* fused post-combine step for spp-sliced multi-GPU frames: d_rgb8 = quantise(d_sum / spp_total)               */
RT_API int rt_resolve_sum_device(rt_scene* s, const float* d_sum, uint32_t spp_total, float* d_rgb_out, uint8_t* d_rgb8_out,
                          void* stream);

/* ---- multi-GPU combine over NVLink peer memory ------------------------------------------------
 * The reference renders one image on one host; its decomposition into independent tiles / samples
 * (render/tile/bucket.hpp:7-21, render/render.hpp:31-35,66-74) is what shards.  One process per GPU: every rank renders its
 * sample slice with RT_FLAG_RAW_SUM into rt_peer_framebuffer(), then rt_peer_combine() runs ONE kernel per rank that waits
 * for all peers on device flags, adds its 1/world slice of every rank's framebuffer through NVLink (rank order = the
 * reference's sample order), divides by spp_total, quantises (io/image/ppm.hpp:17-19) and stores into rank 0's result
 * buffers.  Handles are cudaIpcMemHandle_t bytes; exchange them with any host-side all-gather.                        */
#define RT_PEER_HANDLE_BYTES 64
#define RT_PEER_MAX_RANKS 16
#define RT_PEER_OUT_RGB 1u          /* float result in rank 0's rt_peer_result_rgb()  */
#define RT_PEER_OUT_RGB8 2u         /* 8-bit result in rank 0's rt_peer_result_rgb8() */
#define RT_PEER_OUT_HOST_RGB 4u     /* float result in the shared HOST frame (rt_peer_host_result_attach) */
typedef struct rt_peer_group rt_peer_group;
RT_API int rt_peer_group_create(uint32_t world, uint32_t rank, int device, uint32_t width, uint32_t height, rt_peer_group** out,
                                uint8_t* handle /* RT_PEER_HANDLE_BYTES, may be null */);
RT_API int rt_peer_group_connect(rt_peer_group* g, const uint8_t* handles /* world x RT_PEER_HANDLE_BYTES, rank order */);
/* all ranks inside one process (one GPU emulating several ranks, or several peer GPUs driven by one host thread).  When the
 * ranks share ONE device and stream, signal every rank (rt_peer_signal_ready) before any rank reduces: a reduce waits on the
 * device for its peers' flags, and a kernel queued behind it on the same stream cannot set them (the wait then runs out:
 * RT_ERR_TIMEOUT). */
RT_API int rt_peer_group_connect_local(rt_peer_group* const* groups, uint32_t world);
/* Every rank owns TWO frame slots: frame e+1 can be rendered while frame e is being combined.  rt_peer_framebuffer is where
 * the NEXT frame (the one that will be signalled next) is rendered; it flips with every rt_peer_signal_ready.  The result
 * pointers are those of the frame signalled LAST (rank 0's copy holds the combined frame once every rank is done).        */
RT_API float* rt_peer_framebuffer(rt_peer_group* g);      /* device pointers into this rank's block */
RT_API float* rt_peer_result_rgb(rt_peer_group* g);
RT_API uint8_t* rt_peer_result_rgb8(rt_peer_group* g);
/* signal + reduce/resolve + wait, asynchronous on `stream` (the stream the frame was rendered on)             */
RT_API int rt_peer_combine(rt_peer_group* g, uint32_t spp_total, uint32_t outputs, void* stream);
/* the three steps separately (single-process emulation must signal every rank before any rank reduces)        */
RT_API int rt_peer_signal_ready(rt_peer_group* g, void* stream);
RT_API int rt_peer_reduce_resolve(rt_peer_group* g, uint32_t spp_total, uint32_t outputs, void* stream);
RT_API int rt_peer_wait_done(rt_peer_group* g, void* stream);
/* Pipelined use: render frame e+1 on one stream while rt_peer_reduce_resolve + rt_peer_wait_done of frame e run on another
 * (after rt_peer_signal_ready(e) on the render stream); the render stream must wait for that rt_peer_wait_done before frame
 * e+2 is rendered into the slot frame e used.                                                                              */
/* rank 0: copy the combined frame of the frame signalled last to host buffers (either may be null).
 * rt_peer_download_result only queues the copies on `stream`; rt_peer_read_result also synchronises `stream`.              */
RT_API int rt_peer_download_result(rt_peer_group* g, float* rgb, uint8_t* rgb8, void* stream);
RT_API int rt_peer_read_result(rt_peer_group* g, float* rgb, uint8_t* rgb8, void* stream);
/* Shared host result (one process per GPU): every rank maps and pins the SAME host frame - the POSIX shared-memory object
 * `shm_name` ("/name"; rank 0 passes create = 1 first, the others attach after a barrier, then rank 0 may shm_unlink it) - and
 * rt_peer_reduce_resolve(..., RT_PEER_OUT_HOST_RGB) makes every rank copy ITS slice of the combined float frame into it with
 * its own copy engine over its own PCIe link: N links carry the frame instead of rank 0's one.  *host_rgb = this process's
 * mapping: two frames of height*width*3 floats; frame e (the e-th rt_peer_signal_ready) is in slot (e - 1) & 1 and is complete,
 * on any rank, once that rank's rt_peer_wait_done for the frame has completed (render/render.hpp:18-108 returns this image). */
RT_API int rt_peer_host_result_attach(rt_peer_group* g, const char* shm_name, int create, float** host_rgb);
RT_API void rt_peer_group_destroy(rt_peer_group* g);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
