// csrc/rt_wavefront.cuh - the wavefront renderer: generate -> trace -> shade -> (compact) -> ... -> resolve.
//
// The reference evaluates color_hit (render/render.hpp:133-308) as a recursion: a hit's colour is an expression of
// its children's colours (refractive: fresnel*refl + (1-fresnel)*refr; diffuse: (sum GI + sum lights)/(N+1)).  The
// wavefront keeps that expression tree explicitly, so that the frame is the SAME floating-point expression as the
// recursion and not a re-associated sum of weighted contributions:
//
//   level d of a pass = every closest-hit query issued at recursion depth d (level 0 = primary rays).
//   pool entry        = one query: its ray (32 B), its hit (16 B) and its shading record (32 B).
//   shade(level d)    turns hits into records, and appends the children (next level's rays, compacted with
//                     __ballot_sync/__popc + one atomicAdd per warp) and the shadow jobs of the level.
//   shadow            is_occluded (render/render.hpp:110-131) over all shadow jobs of the pass.
//   resolve(level d)  for d = max..0 folds children and light terms into the record's colour, in the reference's
//                     operation order.
//   accumulate        adds the level-0 colours of the pass into the framebuffer in sample order (render.hpp:66-74).
//
// All kernels are persistent: the grid is sized to the machine, warps claim 32-entry chunks from a device-side
// counter, and the level boundaries live in device memory, so a pass needs no host round trip.
#pragma once

#include "rt_device.cuh"

namespace rtb {

constexpr int MAX_LEVELS = 66;          // max_ray_depth is capped at 64 by the host
constexpr int N_WORK = 4 * MAX_LEVELS;  // one dynamic-fetch counter per launch of a pass

struct Ray { float ox, oy, oz, dx, dy, dz; uint32_t k0, k1; };                       // 32 B
struct Rec { uint32_t kind, first_child, first_shadow; float fresnel, r, g, b; uint32_t pad; };  // 32 B
struct ShadowJob { float ox, oy, oz, dx, dy, dz, max_t, k; };                        // 32 B; max_t < 0 after tracing = occluded

enum : uint32_t { REC_DONE = 0, REC_MISS = 1, REC_REFLECT = 2, REC_TIR = 3, REC_REFRACT = 4, REC_DIFFUSE = 5, REC_TEXTURE = 6 };
constexpr int TRI_INACTIVE = -2;        // level-0 padding lanes outside the tile rectangle

struct FrameParams {
    double tan_half_fov;                // tan(fov_rad/2) in double, evaluated on the host (render.hpp:55-57)
    float eps, shadow_bias, reflection_bias, refraction_bias;
    uint32_t max_ray_depth, gi_rays, seed, spp_total;
    uint32_t sample_first;              // global index of the first sample of this pass
    uint32_t n_samples;                 // samples of this pass
    uint32_t x0, y0, tw, th;            // tile rectangle
    uint32_t tiles_x, plane;            // 8x4 pixel tiles per row; entries per sample plane (= tiles * 32)
    uint32_t pool_cap, shadow_cap;
    uint32_t sparse0, reserved;         // sparse0: only the HITS of the camera rays are level-0 entries (k_stream_primary_sparse)
    // row bands (rt_params.band_rows): the call renders the frame's rows y with (y / band_rows) % band_period == band_phase;
    // th counts those rows and local row ly is the ly-th of them (frame_row).  band_rows == 0: the rectangle x0, y0, tw, th.
    uint32_t band_rows, band_period, band_phase;
};

// local row of the call's rectangle (or of its row bands, in order) -> row of the frame.  band_rows is a multiple of four, so
// the four rows of an 8x4 tile stay adjacent
__host__ __device__ __forceinline__ uint32_t frame_row(const FrameParams& fp, uint32_t ly) {
    if (!fp.band_rows) return fp.y0 + ly;
    const uint32_t b = ly / fp.band_rows;
    return (b * fp.band_period + fp.band_phase) * fp.band_rows + (ly - b * fp.band_rows);
}

struct FrameCounters {                  // device memory, reset at the start of every frame
    unsigned long long primary, primary_hits, shadow, shadow_hits, secondary, secondary_hits;
    unsigned long long stack_overflows; // queries of the four-wide stream kernels that outgrew the shared-memory stack (rt_stream.cuh)
};

struct PassState {                      // device memory, reset at the start of every pass (k_pass_init)
    uint32_t pool_count, shadow_count, overflow;
    uint32_t n_tiles0;                  // sparse level 0: tiles k_tile_cull left for the primary stream kernel
    // Multi-pass frames are queued without a host round trip between passes.  When a pass outgrows its pools it is discarded
    // on the device as before (overflow != 0) and k_pass_commit sets `carry`, the one word k_pass_init keeps: every LATER pass
    // of the frame starts with skipped != 0 and does nothing - the framebuffer holds exactly the passes in front of the failed
    // one, and the host resumes from it after growing the pools.  The first pass of a frame clears `carry`.
    uint32_t skipped, carry;
    uint32_t lv[MAX_LEVELS + 2];        // level d = pool entries [lv[d], lv[d+1])
    uint32_t work[N_WORK];
    FrameCounters pc;                   // this pass's ray counts; folded into the frame's when the pass is kept
};

// ---- work distribution ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t claim32(uint32_t* counter) {
    uint32_t base = 0;
    if ((threadIdx.x & 31u) == 0) base = atomicAdd(counter, 32u);
    return __shfl_sync(0xFFFFFFFFu, base, 0);
}
// Streaming kernels (shade, resolve) cost about the same per entry, so their warps take 32-entry chunks in a fixed
// grid-stride order: no atomic round trip in front of every chunk's loads (the dependent atomic + load chain made these
// kernels latency-bound at a third of the HBM rate).  The trace kernels keep the dynamic counter: their cost per ray varies.
__device__ __forceinline__ uint32_t first_chunk() { return (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u; }
__device__ __forceinline__ uint32_t chunk_stride() { return gridDim.x * (blockDim.x >> 5) * 32u; }
// The 32-entry chunks of a level for one warp of a streaming kernel.  Dense: grid-stride order, every entry live.  Sparse
// level 0 (mask = one word per chunk, bit = the entry exists): the warp reads the words of its next 32 chunks with ONE load
// and visits the non-empty ones; a dependent load in front of every chunk, most of them empty, kept the level-0 kernels
// waiting for a third of their time.  Every lane of the warp has to call next() (it shuffles).
struct ChunkIter {
    const uint32_t* mask;
    uint32_t n_chunks, c0, cbase, nz, words, nw, lane;
    __device__ __forceinline__ ChunkIter(uint32_t n_entries, const uint32_t* mask_or_null)
        : mask(mask_or_null), n_chunks((n_entries + 31u) >> 5), c0(first_chunk() >> 5), cbase(0), nz(0), words(0),
          nw(chunk_stride() >> 5), lane(threadIdx.x & 31u) {}
    // base: first entry of the chunk (relative to the level); word: which of its 32 entries exist
    __device__ __forceinline__ bool next(uint32_t& base, uint32_t& word) {
        __syncwarp();
        if (!mask) {
            if (c0 >= n_chunks) return false;
            base = c0 << 5; word = 0xFFFFFFFFu; c0 += nw;
            return true;
        }
        while (!nz) {
            if (c0 >= n_chunks) return false;
            const uint32_t c = c0 + lane * nw;
            words = c < n_chunks ? mask[c] : 0u;
            nz = __ballot_sync(0xFFFFFFFFu, words != 0u);
            cbase = c0; c0 += nw << 5;
        }
        const int l = __ffs(nz) - 1;
        nz &= nz - 1u;
        word = __shfl_sync(0xFFFFFFFFu, words, l);
        base = (cbase + uint32_t(l) * nw) << 5;
        return true;
    }
};
// warp-aggregated append: every lane asks for n slots, returns its first slot
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter, uint32_t n) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= uint32_t(o)) incl += v;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(counter, total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31);
    return base + incl - n;
}
__device__ __forceinline__ void warp_count(unsigned long long* counter, bool pred) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, pred);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(counter, (unsigned long long)__popc(m));
}

__device__ __forceinline__ void store_ray(Ray* __restrict__ dst, V3 o, V3 d, uint2 key) {
    float4* p = reinterpret_cast<float4*>(dst);
    p[0] = make_float4(o.x, o.y, o.z, d.x);
    p[1] = make_float4(d.y, d.z, __uint_as_float(key.x), __uint_as_float(key.y));
}
__device__ __forceinline__ void load_ray(const Ray* __restrict__ src, V3& o, V3& d, uint2& key) {
    const float4* p = reinterpret_cast<const float4*>(src);
    const float4 a = p[0], b = p[1];
    o = mk(a.x, a.y, a.z); d = mk(a.w, b.x, b.y);
    key = make_uint2(__float_as_uint(b.z), __float_as_uint(b.w));
}
__device__ __forceinline__ void store_hit(Hit* __restrict__ dst, const Hit& h) {
    *reinterpret_cast<float4*>(dst) = make_float4(h.t, h.u, h.v, __int_as_float(h.tri));
}
__device__ __forceinline__ Hit load_hit(const Hit* __restrict__ src) {
    const float4 a = *reinterpret_cast<const float4*>(src);
    Hit h; h.t = a.x; h.u = a.y; h.v = a.z; h.tri = __float_as_int(a.w);
    return h;
}
__device__ __forceinline__ void store_rec(Rec* __restrict__ dst, uint32_t kind, uint32_t first_child, uint32_t first_shadow,
                                          float fresnel, float r, float g, float b) {
    float4* p = reinterpret_cast<float4*>(dst);
    p[0] = make_float4(__uint_as_float(kind), __uint_as_float(first_child), __uint_as_float(first_shadow), fresnel);
    p[1] = make_float4(r, g, b, 0.0f);
}

// ---- camera rays: render_frame, render/render.hpp:35-62 ----------------------------------------------------------
__device__ __forceinline__ void camera_ray(const DScene& sc, double tan_half_fov, float raster_x, float raster_y, V3& o, V3& d) {
    const float aspect = __fdiv_rn(float(sc.width), float(sc.height));                                   // :27
    const float ndc_x = __fdiv_rn(raster_x, float(sc.width));                                            // :47
    const float ndc_y = __fdiv_rn(raster_y, float(sc.height));                                           // :48
    float sx = (2.0f * ndc_x) - 1.0f;                                                                    // :50
    float sy = 1.0f - (2.0f * ndc_y);                                                                    // :51
    sx *= aspect;                                                                                        // :53
    sx = float(__dmul_rn(double(sx), tan_half_fov));                                                     // :56  (double product, narrowed)
    sy = float(__dmul_rn(double(sy), tan_half_fov));                                                     // :57
    const float* m = sc.cam_m;                                                                           // transpose(M) * (sx, sy, -1), :59-60
    const V3 v = mk(m[0] * sx + m[3] * sy + m[6] * -1.0f, m[1] * sx + m[4] * sy + m[7] * -1.0f, m[2] * sx + m[5] * sy + m[8] * -1.0f);
    o = mk(sc.cam_pos[0], sc.cam_pos[1], sc.cam_pos[2]);
    d = normalized(v);
}

// pixel + sample -> raster position and path key.  spp_total == 1: pixel centre (render.hpp:39-42); otherwise the
// jitter comes from the root Philox block (the reference draws it from urand01(), render.hpp:43-44).
__device__ __forceinline__ void primary_sample(const DScene& sc, const FrameParams& fp, uint32_t x, uint32_t y, uint32_t sample,
                                               float& rx, float& ry, uint2& key) {
    rx = float(x); ry = float(y);
    key = make_uint2(0u, 0u);
    if (fp.spp_total == 1u && fp.gi_rays == 0u) { rx += 0.5f; ry += 0.5f; return; }
    const uint4 r = philox4x32_10(make_uint4(y * sc.width + x, sample, 0u, TAG_ROOT), make_uint2(fp.seed, 0u));
    key = make_uint2(r.x, r.y);
    if (fp.spp_total == 1u) { rx += 0.5f; ry += 0.5f; }
    else { rx += u01(r.z); ry += u01(r.w); }
}

__device__ __forceinline__ bool level0_pixel(const FrameParams& fp, uint32_t j, uint32_t& x, uint32_t& y) {
    const uint32_t tile = j >> 5, l = j & 31u;
    const uint32_t lx = (tile % fp.tiles_x) * 8u + (l & 7u), ly = (tile / fp.tiles_x) * 4u + (l >> 3);
    x = fp.x0 + lx; y = frame_row(fp, ly);
    return lx < fp.tw && ly < fp.th;
}

// ---- kernels: trace -------------------------------------------------------------------------------------------------
template <bool FAST, bool ORDERED>
__global__ void __launch_bounds__(256) k_primary(DScene sc, FrameParams fp, Ray* __restrict__ rays, Hit* __restrict__ hits,
                                                 PassState* __restrict__ ps, int work_slot) {
    pdl_wait();
    if (ps->skipped) return;
    FrameCounters* fc = &ps->pc;
    const uint32_t n0 = fp.plane * fp.n_samples;
    unsigned long long n_rays = 0, n_hits = 0;
    for (;;) {
        const uint32_t base = claim32(&ps->work[work_slot]);
        if (base >= n0) break;
        const uint32_t i = base + (threadIdx.x & 31u);
        const uint32_t s = i / fp.plane, j = i - s * fp.plane;
        uint32_t x, y;
        const bool valid = level0_pixel(fp, j, x, y);
        uint2 key = make_uint2(0u, 0u);
        V3 o = mk(0, 0, 0), d = mk(0, 0, 0);
        if (valid) {
            float rx, ry;
            primary_sample(sc, fp, x, y, fp.sample_first + s, rx, ry, key);
            camera_ray(sc, fp.tan_half_fov, rx, ry, o, d);
        }
        Hit h = trace_any<true, FAST, ORDERED>(sc, valid, o, d, fp.eps);                                // render.hpp:64, culling ON
        if (valid) {
            store_ray(rays + i, o, d, key);
            ++n_rays; n_hits += (h.tri >= 0);
        } else h.tri = TRI_INACTIVE;
        store_hit(hits + i, h);
    }
    // one atomic per warp
#pragma unroll
    for (int o = 16; o; o >>= 1) { n_rays += __shfl_xor_sync(0xFFFFFFFFu, n_rays, o); n_hits += __shfl_xor_sync(0xFFFFFFFFu, n_hits, o); }
    if ((threadIdx.x & 31u) == 0 && n_rays) { atomicAdd(&fc->primary, n_rays); atomicAdd(&fc->primary_hits, n_hits); }
}

// level d >= 1: closest hit without culling (render.hpp:175,244,269,284,293) over [lv[d], pool_count)
template <bool FAST, bool ORDERED>
__global__ void __launch_bounds__(256) k_trace_level(DScene sc, FrameParams fp, const Ray* __restrict__ rays, Hit* __restrict__ hits,
                                                     PassState* __restrict__ ps, int level, int work_slot) {
    pdl_wait();
    FrameCounters* fc = &ps->pc;
    const uint32_t begin = ps->lv[level];
    const uint32_t end = min(ps->pool_count, fp.pool_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) ps->lv[level + 1] = end;
    unsigned long long n_rays = 0, n_hits = 0;
    for (;;) {
        const uint32_t base = claim32(&ps->work[work_slot]);
        if (base >= end - begin) break;
        const uint32_t i = begin + base + (threadIdx.x & 31u);
        const bool valid = i < end;
        V3 o = mk(0, 0, 0), d = mk(0, 0, 0); uint2 key;
        if (valid) load_ray(rays + i, o, d, key);
        const Hit h = trace_any<false, FAST, ORDERED>(sc, valid, o, d, fp.eps);
        if (valid) {
            store_hit(hits + i, h);
            ++n_rays; n_hits += (h.tri >= 0);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { n_rays += __shfl_xor_sync(0xFFFFFFFFu, n_rays, o); n_hits += __shfl_xor_sync(0xFFFFFFFFu, n_hits, o); }
    if ((threadIdx.x & 31u) == 0 && n_rays) { atomicAdd(&fc->secondary, n_rays); atomicAdd(&fc->secondary_hits, n_hits); }
}

// is_occluded, render/render.hpp:110-131.  Every loop iteration is one closest-hit query without culling.
// TRANSMISSIVE == false (no refractive material in the scene): the loop body runs at most once and its answer is
// "closest.t <= max_t"; the running closest only ever decreases, so the traversal may stop as soon as it holds a
// candidate with t <= max_t - same answer, same query count.
template <bool TRANSMISSIVE, bool FAST, bool ORDERED>
__device__ __forceinline__ bool occluded_query(const DScene& sc, bool active, V3 o, V3 d, float max_t, float eps, float shadow_bias,
                                               unsigned long long& n_q, unsigned long long& n_h) {
    // called by all 32 lanes (the closest-hit query is warp-cooperative); a lane leaves the loop by clearing `active`
    bool result = false;
    active = active && (0.0f < max_t);                                                                   // :115
    while (__ballot_sync(0xFFFFFFFFu, active)) {
        const Hit h = trace_any<false, FAST, ORDERED>(sc, active, o, d, eps, max_t, !TRANSMISSIVE);      // :116
        if (active) {
            ++n_q;
            if (h.tri < 0) active = false;                                                               // :117
            else {
                ++n_h;
                if (max_t < h.t) active = false;                                                         // :117-119
                else if (!TRANSMISSIVE) { result = true; active = false; }
                else {
                    const uint32_t mat = __ldg(&sc.tri_index[h.tri]).w;
                    if (sc.materials[mat].kind != 2u) { result = true; active = false; }                 // :121-124
                    else {
                        const V3 pos = o + h.t * d;                                                      // hit.position, kd_tree_simd.hpp:254
                        o = pos + shadow_bias * d;                                                       // :126 (direction, hence inv_direction, unchanged)
                        max_t -= h.t;                                                                    // :127
                        active = 0.0f < max_t;                                                           // :115
                    }
                }
            }
        }
    }
    return result;
}

template <bool TRANSMISSIVE, bool FAST, bool ORDERED>
__global__ void __launch_bounds__(256) k_shadow(DScene sc, FrameParams fp, ShadowJob* __restrict__ jobs, PassState* __restrict__ ps,
                                                int work_slot) {
    pdl_wait();
    FrameCounters* fc = &ps->pc;
    const uint32_t end = min(ps->shadow_count, fp.shadow_cap);
    unsigned long long n_q = 0, n_h = 0;
    for (;;) {
        const uint32_t base = claim32(&ps->work[work_slot]);
        if (base >= end) break;
        const uint32_t i = base + (threadIdx.x & 31u);
        const bool valid = i < end;
        float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
        if (valid) { const float4* p = reinterpret_cast<const float4*>(jobs + i); a = p[0]; b = p[1]; }
        const bool occ = occluded_query<TRANSMISSIVE, FAST, ORDERED>(sc, valid, mk(a.x, a.y, a.z), mk(a.w, b.x, b.y), b.z, fp.eps,
                                                                     fp.shadow_bias, n_q, n_h);
        if (occ) jobs[i].max_t = -1.0f;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { n_q += __shfl_xor_sync(0xFFFFFFFFu, n_q, o); n_h += __shfl_xor_sync(0xFFFFFFFFu, n_h, o); }
    if ((threadIdx.x & 31u) == 0 && n_q) { atomicAdd(&fc->shadow, n_q); atomicAdd(&fc->shadow_hits, n_h); }
}

// ---- textures: scene/texture/{albedo,edge,checker,bitmap}.hpp ------------------------------------------------------
__device__ __forceinline__ V3 sample_texture(const DScene& sc, const DTexture& t, float hu, float hv, const float4 uv01, const float4 uv2) {
    const float hw = float(__dsub_rn(__dsub_rn(1.0, double(hu)), double(hv)));      // `1. - u - v` in double, narrowed
    if (t.kind == 0u) return mk(t.c0[0], t.c0[1], t.c0[2]);                                              // albedo.hpp:11-13
    if (t.kind == 1u) {                                                                                  // edge.hpp:13-22
        if (hu < t.scalar || hv < t.scalar || hw < t.scalar) return mk(t.c0[0], t.c0[1], t.c0[2]);
        return mk(t.c1[0], t.c1[1], t.c1[2]);
    }
    const float fx = (hw * uv01.x + hu * uv01.z) + hv * uv2.x;                                           // w*uv0 + u*uv1 + v*uv2, L->R
    const float fy = (hw * uv01.y + hu * uv01.w) + hv * uv2.y;
    if (t.kind == 2u) {                                                                                  // checker.hpp:12-26
        const int u2 = int(__fdiv_rn(fx, t.scalar)), v2 = int(__fdiv_rn(fy, t.scalar));
        if ((u2 + v2) % 2 == 0) return mk(t.c0[0], t.c0[1], t.c0[2]);
        return mk(t.c1[0], t.c1[1], t.c1[2]);
    }
    // bitmap.hpp:46-60: row in double, column in float, size_t conversion, clamp to [0, dim-1]
    unsigned long long row = (unsigned long long)(long long)(__dmul_rn(__dsub_rn(1.0, double(fy)), double(t.h)));
    unsigned long long col = (unsigned long long)(long long)(fx * float(t.w));
    if (row > (unsigned long long)t.h - 1ull) row = (unsigned long long)t.h - 1ull;
    if (col > (unsigned long long)t.w - 1ull) col = (unsigned long long)t.w - 1ull;
    const uint8_t* px = sc.texels + t.off + (row * t.w + col) * 3ull;
    const float scale = float(1.0 / 255.0);                                                              // bitmap.hpp:19
    return mk(float(px[0]) * scale, float(px[1]) * scale, float(px[2]) * scale);                         // :27-29
}

// (float)cos((double)x): the double-precision routine, rounded once.  glibc's cosf/sinf (what the reference's GI
// sampling calls, render.hpp:160-168) evaluate a double polynomial and round once as well; the two agree except
// on rare rounding-boundary inputs, which is why GI parity is statistical while the deterministic paths are exact.
__device__ __forceinline__ float cos_f(float x) { return float(cos(double(x))); }
__device__ __forceinline__ float sin_f(float x) { return float(sin(double(x))); }

// ---- kernel: shade ---------------------------------------------------------------------------------------------------
// color_hit (render/render.hpp:133-308) up to the point where it needs a child's colour.
template <bool HAS_GI>
__global__ void __launch_bounds__(256) k_shade(DScene sc, FrameParams fp, Ray* __restrict__ rays, const Hit* __restrict__ hits,
                                               Rec* __restrict__ recs, ShadowJob* __restrict__ jobs, PassState* __restrict__ ps,
                                               int level, int work_slot, const uint32_t* __restrict__ mask0) {
    pdl_wait();
    const uint32_t begin = ps->lv[level], end = ps->lv[level + 1];
    const float PI = 3.14159265358979323846f;
    const bool sparse = level == 0 && fp.sparse0 != 0u;           // level 0 starts at entry 0: chunk c is tile c, mask0[c] its hits
    (void)work_slot;
    ChunkIter chunks(end - begin, sparse ? mask0 : nullptr);
    uint32_t base, word;
    while (chunks.next(base, word)) {
        const uint32_t i = begin + base + (threadIdx.x & 31u);
        const bool live = i < end && ((word >> (threadIdx.x & 31u)) & 1u);

        uint32_t kind = REC_DONE, n_child = 0, n_shadow = 0;
        float fresnel = 0.0f;
        V3 col = mk(0.0f, 0.0f, 0.0f);
        // state carried from the material switch to the emission step
        V3 o = mk(0, 0, 0), I = mk(0, 0, 0), P = mk(0, 0, 0), hn = mk(0, 0, 0), fnrm = mk(0, 0, 0);
        V3 c0o = mk(0, 0, 0), c0d = mk(0, 0, 0), c1o = mk(0, 0, 0), c1d = mk(0, 0, 0);
        uint2 key = make_uint2(0u, 0u);
        uint32_t smooth = 0;
        Hit h; h.t = 0; h.u = 0; h.v = 0; h.tri = TRI_INACTIVE;
        if (live) h = load_hit(hits + i);

        if (live && h.tri == -1) kind = REC_MISS;
        else if (live && h.tri >= 0) {
            if (uint32_t(level) == fp.max_ray_depth) {                                                   // :138-139
                col = mk(sc.bg[0], sc.bg[1], sc.bg[2]);
            } else {
                load_ray(rays + i, o, I, key);
                // hit assembly, kd_tree_simd.hpp:234-263
                const uint4 ti = __ldg(&sc.tri_index[h.tri]);
                const DMaterial m = sc.materials[ti.w];
                smooth = m.smooth;
                const float u = h.u, v = h.v;
                const float w = 1.0f - u - v;                                                            // :238
                const float4 n0 = __ldg(&sc.vnormals[ti.x]), n1 = __ldg(&sc.vnormals[ti.y]), n2 = __ldg(&sc.vnormals[ti.z]);
                hn = normalized((u * mk(n1.x, n1.y, n1.z) + v * mk(n2.x, n2.y, n2.z)) + w * mk(n0.x, n0.y, n0.z));   // :250
                P = o + h.t * I;                                                                         // :254
                const float4 fn4 = __ldg(&sc.tri_normal[h.tri]);
                fnrm = mk(fn4.x, fn4.y, fn4.z);
                switch (m.kind) {
                case 3u:                                                                                 // constant :302-303
                    col = mk(m.albedo[0], m.albedo[1], m.albedo[2]);
                    break;
                case 1u: {                                                                               // reflective :239-250
                    const float k = 2.0f * dot(I, hn);
                    c0d = I - k * hn;
                    c0o = P + fp.reflection_bias * c0d;
                    kind = REC_REFLECT; n_child = 1;
                    break;
                }
                case 2u: {                                                                               // refractive :251-301
                    V3 n = normalized(smooth ? hn : fnrm);
                    const V3 iv = normalized(I);
                    float eta_i = 1.0f, eta_r = m.ior;
                    if (0.0f < dot(iv, n)) { const float t = eta_i; eta_i = eta_r; eta_r = t; n = -n; } // :258-261
                    const float cos_i = -dot(iv, n);
                    const float sin_i = __fsqrt_rn(1.0f - cos_i * cos_i);
                    const float k2 = 2.0f * dot(iv, n);
                    const V3 refl_d = iv - k2 * n;                                                       // :267, :291
                    const V3 refl_o = P + fp.reflection_bias * refl_d;
                    if (__fdiv_rn(eta_r, eta_i) < sin_i) {                                               // :266 total internal reflection
                        c0o = refl_o; c0d = refl_d;
                        kind = REC_TIR; n_child = 1;
                    } else {
                        const float sin_r = __fdiv_rn(sin_i * eta_i, eta_r);                             // :278
                        const float cos_r = __fsqrt_rn(1.0f - sin_r * sin_r);
                        const V3 tang = normalized(iv + cos_i * n);
                        const V3 rdir = cos_r * (-n) + sin_r * tang;                                     // :281
                        c0o = P + fp.refraction_bias * rdir; c0d = rdir;                                 // :283
                        c1o = refl_o; c1d = refl_d;                                                      // :292
                        // :300 - 0.5 * pow(double(1 + i.n), 5): x^5 by multiplication in double, narrowed once
                        const double x = double(1.0f + dot(iv, n));
                        const double x2 = __dmul_rn(x, x);
                        fresnel = float(__dmul_rn(0.5, __dmul_rn(__dmul_rn(x2, x2), x)));
                        kind = REC_REFRACT; n_child = 2;
                    }
                    break;
                }
                case 0u:                                                                                 // diffuse :149-210
                    col = mk(m.albedo[0], m.albedo[1], m.albedo[2]);
                    kind = REC_DIFFUSE; n_child = HAS_GI ? fp.gi_rays : 0u; n_shadow = sc.n_lights;
                    break;
                default: {                                                                               // texture :211-238
                    const float4 uv01 = __ldg(&sc.tri_uv[2 * h.tri]), uv2 = __ldg(&sc.tri_uv[2 * h.tri + 1]);
                    col = sample_texture(sc, sc.textures[m.texture], u, v, uv01, uv2);
                    kind = REC_TEXTURE; n_shadow = sc.n_lights;
                    break;
                }
                }
            }
        }

        // ---- compaction: children go to the next level, shadow jobs to the pass's job list ----
        const uint32_t first_child = warp_append(&ps->pool_count, n_child);
        const uint32_t first_shadow = warp_append(&ps->shadow_count, n_shadow);
        if (n_child && first_child + n_child > fp.pool_cap) { atomicOr(&ps->overflow, 1u); n_child = 0; kind = REC_DONE; }
        if (n_shadow && first_shadow + n_shadow > fp.shadow_cap) { atomicOr(&ps->overflow, 2u); n_shadow = 0; kind = REC_DONE; }

        if (kind == REC_REFLECT || kind == REC_TIR) {
            store_ray(rays + first_child, c0o, c0d, HAS_GI ? child_key(key, SLOT_REFLECT) : key);
        } else if (kind == REC_REFRACT) {
            store_ray(rays + first_child, c0o, c0d, HAS_GI ? child_key(key, SLOT_REFRACT) : key);
            store_ray(rays + first_child + 1, c1o, c1d, HAS_GI ? child_key(key, SLOT_REFLECT) : key);
        } else if (HAS_GI && kind == REC_DIFFUSE) {
            // GI bounce spawning, render.hpp:151-182
            const V3 right = normalized(cross(I, hn)), up = hn, fwd = cross(right, up);
            const V3 org = P + fp.reflection_bias * hn;                                                  // :172
            for (uint32_t g = 0; g < n_child; ++g) {
                const uint4 r = philox4x32_10(make_uint4(g, 0u, 0u, TAG_GI), key);
                const float u1 = u01(r.x), u2 = u01(r.y);
                const float a1 = PI * u1;                                                                // :160
                const V3 rv = mk(cos_f(a1), sin_f(a1), 0.0f);                                            // :161
                const float a2 = PI * u2 * 2.0f;                                                         // :163
                const float ca = cos_f(a2), sa = sin_f(a2);
                const V3 rot = mk(ca * rv.x + 0.0f * rv.y + (-sa) * rv.z,                                // :164-168, mat3.hpp:53-60
                                  0.0f * rv.x + 1.0f * rv.y + 0.0f * rv.z,
                                  sa * rv.x + 0.0f * rv.y + ca * rv.z);
                const V3 dir = mk(right.x * rot.x + right.y * rot.y + right.z * rot.z,                   // local_hit_mat * rot, :170-171
                                  up.x * rot.x + up.y * rot.y + up.z * rot.z,
                                  fwd.x * rot.x + fwd.y * rot.y + fwd.z * rot.z);
                store_ray(rays + first_child + g, org, dir, child_key(key, SLOT_GI0 + g));
            }
        }
        if (n_shadow) {
            // direct lighting terms, render.hpp:184-206 / :213-236
            for (uint32_t li = 0; li < n_shadow; ++li) {
                const DLight L = sc.lights[li];
                V3 ld = mk(L.pos[0], L.pos[1], L.pos[2]) - P;
                const float radius = len(ld);
                const float area = 4.0f * PI * radius * radius;                                          // L->R
                ld = normalized(ld);
                const float cosine = max_std(0.0f, dot(ld, smooth ? hn : fnrm));
                const V3 so = P + fp.shadow_bias * ld;
                const float k = __fdiv_rn(L.intensity, area) * cosine;
                float4* p = reinterpret_cast<float4*>(jobs + first_shadow + li);
                p[0] = make_float4(so.x, so.y, so.z, ld.x);
                p[1] = make_float4(ld.y, ld.z, radius, k);
            }
        }
        if (live) store_rec(recs + i, kind, first_child, first_shadow, fresnel, col.x, col.y, col.z);
    }
}

// ---- kernel: resolve -------------------------------------------------------------------------------------------------
// the part of color_hit that consumes child colours, deepest level first
__device__ __forceinline__ V3 child_colour(const Rec* __restrict__ recs, uint32_t idx, V3 on_miss) {
    const float4* p = reinterpret_cast<const float4*>(recs + idx);
    const float4 a = p[0], b = p[1];
    return (__float_as_uint(a.x) == REC_MISS) ? on_miss : mk(b.x, b.y, b.z);
}

// ACC (level 0 of a one-sample pass): the colour goes straight into the framebuffer - render.hpp:66-74 for one sample, the
// same operations k_accumulate performs - instead of being written to the record and read back by another kernel.  The
// pass may still be discarded (pool overflow, or a skipped level that was needed, see k_pass_commit): that is known once
// the last shade kernel has run, so it is evaluated here and nothing is written then.
template <bool ACC>
__global__ void __launch_bounds__(256) k_resolve(DScene sc, FrameParams fp, Rec* __restrict__ recs, const ShadowJob* __restrict__ jobs,
                                                 PassState* __restrict__ ps, int level, int work_slot, float* __restrict__ fb,
                                                 int first_pass, int divide, uint32_t launched, uint32_t total,
                                                 uint32_t* __restrict__ mask0) {
    pdl_wait();
    // sparse level 0: the misses wrote their pixels in k_stream_primary_sparse, mask0[tile] holds the hits.  The word is cleared
    // here, by its last reader, so the mask is all zero again when the pass ends (kept or discarded)
    const bool sparse = level == 0 && fp.sparse0 != 0u;
    const uint32_t begin = ps->lv[level], end = ps->lv[level + 1];
    const V3 bg = mk(sc.bg[0], sc.bg[1], sc.bg[2]), black = mk(0.0f, 0.0f, 0.0f);
    (void)work_slot;
    bool discard = false;
    if (ACC) discard = ps->overflow != 0u || (launched < total && ps->pool_count > ps->lv[launched]);
    ChunkIter chunks(end - begin, sparse ? mask0 : nullptr);
    uint32_t base, word;
    while (chunks.next(base, word)) {
        const uint32_t i = begin + base + (threadIdx.x & 31u);
        if (sparse && ACC && (threadIdx.x & 31u) == 0u) mask0[base >> 5] = 0u;   // multi-sample pass: k_accumulate reads it last
        if (i >= end || !((word >> (threadIdx.x & 31u)) & 1u)) continue;
        float4* p = reinterpret_cast<float4*>(recs + i);
        const float4 a = p[0];
        const uint32_t kind = __float_as_uint(a.x), fc = __float_as_uint(a.y), fs = __float_as_uint(a.z);
        V3 out;
        if (kind <= REC_MISS) {
            if (!ACC) continue;
            if (kind == REC_MISS) out = bg;
            else { const float4 b = p[1]; out = mk(b.x, b.y, b.z); }
        } else if (kind == REC_REFLECT) out = child_colour(recs, fc, bg);                                // :245-249
        else if (kind == REC_TIR) out = child_colour(recs, fc, black);                                   // :271-275
        else if (kind == REC_REFRACT) {
            const V3 refr = child_colour(recs, fc, black), refl = child_colour(recs, fc + 1, black);
            const float f = a.w, omf = 1.0f - f;
            out = mk(f * refl.x + omf * refr.x, f * refl.y + omf * refr.y, f * refl.z + omf * refr.z);   // :301
        } else {
            const float4 b = p[1];
            const V3 albedo = mk(b.x, b.y, b.z);
            out = black;
            if (kind == REC_DIFFUSE)
                for (uint32_t g = 0; g < fp.gi_rays; ++g) {                                              // :175-181, a miss adds nothing
                    const float4* q = reinterpret_cast<const float4*>(recs + fc + g);
                    if (__float_as_uint(q[0].x) == REC_MISS) continue;
                    const float4 c = q[1];
                    out = out + mk(c.x, c.y, c.z);
                }
            for (uint32_t li = 0; li < sc.n_lights; ++li) {                                              // :200-205
                const float4 j = reinterpret_cast<const float4*>(jobs + fs + li)[1];
                if (j.z < 0.0f) continue;                                                                // occluded
                out = out + j.w * albedo;
            }
            if (kind == REC_DIFFUSE) {                                                                   // :208
                const float div = float(fp.gi_rays + 1u);
                out = mk(__fdiv_rn(out.x, div), __fdiv_rn(out.y, div), __fdiv_rn(out.z, div));
            }
        }
        if (!ACC) { p[1] = make_float4(out.x, out.y, out.z, 0.0f); continue; }
        // level 0, one sample: entry index == position in the sample plane
        uint32_t x, y;
        if (discard || !level0_pixel(fp, i - begin, x, y)) continue;
        float* px = fb + (size_t(y) * sc.width + x) * 3;
        V3 sum = first_pass ? mk(0.0f, 0.0f, 0.0f) : mk(px[0], px[1], px[2]);
        sum = sum + out;
        if (divide) {
            const float div = float(fp.spp_total);
            sum = mk(__fdiv_rn(sum.x, div), __fdiv_rn(sum.y, div), __fdiv_rn(sum.z, div));
        }
        px[0] = sum.x; px[1] = sum.y; px[2] = sum.z;
    }
}

// ---- kernel: accumulate ------------------------------------------------------------------------------------------------
// render.hpp:66-74: per pixel, samples are summed in order and divided by samples_per_pixel once
__global__ void __launch_bounds__(256) k_accumulate(DScene sc, FrameParams fp, const Rec* __restrict__ recs, float* __restrict__ fb,
                                                    const PassState* __restrict__ ps, int first_pass, int divide, uint32_t* __restrict__ mask0) {
    pdl_wait();
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;               // plane is a multiple of 32: a warp is one 8x4 tile
    if (j >= fp.plane || ps->skipped) return;
    const bool discard = ps->overflow != 0u;        // an overflowed pass is discarded and rendered again by the host
    uint32_t x, y;
    const bool valid = level0_pixel(fp, j, x, y);
    float* px = fb + (size_t(y) * sc.width + x) * 3;
    V3 sum = (first_pass || !valid || discard) ? mk(0.0f, 0.0f, 0.0f) : mk(px[0], px[1], px[2]);
    const V3 bg = mk(sc.bg[0], sc.bg[1], sc.bg[2]);
    if (fp.sparse0) {
        // sparse level 0: a sample without its bit in mask0 missed everything; the word is cleared here, by its last reader
        for (uint32_t s = 0; s < fp.n_samples; ++s) {
            const uint32_t e = s * fp.plane + j;
            const uint32_t word = mask0[e >> 5];
            __syncwarp();
            if ((threadIdx.x & 31u) == 0u && word) mask0[e >> 5] = 0u;
            sum = sum + (((word >> (j & 31u)) & 1u) ? child_colour(recs, e, bg) : bg);
        }
    } else
        for (uint32_t s = 0; s < fp.n_samples; ++s) sum = sum + child_colour(recs, s * fp.plane + j, bg);
    if (!valid || discard) return;
    if (divide) {
        const float div = float(fp.spp_total);
        sum = mk(__fdiv_rn(sum.x, div), __fdiv_rn(sum.y, div), __fdiv_rn(sum.z, div));
    }
    px[0] = sum.x; px[1] = sum.y; px[2] = sum.z;
}

// io/image/ppm.hpp:17-19: uint8(255.999 * clamp(c, 0, 1)), product in double
__device__ __forceinline__ uint8_t quantise(float c) {
    const float cl = fminf(fmaxf(c, 0.0f), 1.0f);
    return uint8_t(__dmul_rn(255.999, double(cl)));
}
__global__ void k_quantise(const float* __restrict__ rgb, uint8_t* __restrict__ out, size_t n) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = quantise(rgb[i]);
}
// post-combine step of an spp-sliced multi-GPU frame: colour = sum / spp_total, optionally quantised
__global__ void k_resolve_sum(const float* __restrict__ sum, float div, float* __restrict__ rgb, uint8_t* __restrict__ rgb8, size_t n) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float c = __fdiv_rn(sum[i], div);
    if (rgb) rgb[i] = c;
    if (rgb8) rgb8[i] = quantise(c);
}

// ---- batch queries (rt_trace_closest / rt_trace_occluded) ----------------------------------------------------------------
template <bool CULL, bool FAST, bool ORDERED>
__global__ void __launch_bounds__(256) k_trace_batch(DScene sc, const float* __restrict__ rays6, unsigned long long n, float eps,
                                                     Hit* __restrict__ hits) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounds = (n + stride - 1) / stride;                     // warp-uniform trip count
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned long long r = 0; r < rounds; ++r, i += stride) {
        const bool valid = i < n;
        V3 o = mk(0, 0, 0), d = mk(0, 0, 0);
        if (valid) { const float* q = rays6 + 6 * i; o = mk(q[0], q[1], q[2]); d = mk(q[3], q[4], q[5]); }
        const Hit h = trace_any<CULL, FAST, ORDERED>(sc, valid, o, d, eps);
        if (valid) store_hit(hits + i, h);
    }
}
template <bool TRANSMISSIVE, bool FAST, bool ORDERED>
__global__ void __launch_bounds__(256) k_occluded_batch(DScene sc, const float* __restrict__ rays6, const float* __restrict__ max_t,
                                                        unsigned long long n, float eps, float shadow_bias, uint8_t* __restrict__ out) {
    unsigned long long n_q = 0, n_h = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounds = (n + stride - 1) / stride;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned long long r = 0; r < rounds; ++r, i += stride) {
        const bool valid = i < n;
        V3 o = mk(0, 0, 0), d = mk(0, 0, 0);
        float mt = 0.0f;
        if (valid) { const float* q = rays6 + 6 * i; o = mk(q[0], q[1], q[2]); d = mk(q[3], q[4], q[5]); mt = max_t[i]; }
        const bool occ = occluded_query<TRANSMISSIVE, FAST, ORDERED>(sc, valid, o, d, mt, eps, shadow_bias, n_q, n_h);
        if (valid) out[i] = occ ? 1 : 0;
    }
}
// primary rays only, hits in row-major tile order (rt_trace_primary)
template <bool FAST, bool ORDERED>
__global__ void __launch_bounds__(256) k_primary_hits(DScene sc, FrameParams fp, Hit* __restrict__ hits) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = 0, y = 0;
    const bool valid = j < fp.plane && level0_pixel(fp, j, x, y);
    uint2 key; V3 o = mk(0, 0, 0), d = mk(0, 0, 0);
    if (valid) {
        float rx, ry;
        primary_sample(sc, fp, x, y, fp.sample_first, rx, ry, key);
        camera_ray(sc, fp.tan_half_fov, rx, ry, o, d);
    }
    const Hit h = trace_any<true, FAST, ORDERED>(sc, valid, o, d, fp.eps);
    if (valid) store_hit(hits + size_t(y - fp.y0) * fp.tw + (x - fp.x0), h);
}

}  // namespace rtb
