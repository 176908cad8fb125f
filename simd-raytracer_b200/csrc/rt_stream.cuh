// csrc/rt_stream.cuh - the trace kernels of the accelerated mode (RT_FLAG_ORDERED): persistent warps that keep every lane
// busy.  A lane owns one query at a time (the BvhState machine of rt_bvh.cuh); when it finishes - early for an occluded shadow ray,
// at once for a camera ray that misses the scene box - it takes the next query from the level's device-side counter
// instead of idling until the slowest ray of a 32-ray chunk is done.  Bursts of traversal steps alternate with a
// warp-uniform completion phase, which is where exact-t ties are re-run in reference order (trace_warp) and where the
// is_occluded loop (render/render.hpp:110-131) re-arms a lane that passed through a refractive surface.
#pragma once

#include "rt_wavefront.cuh"
#include "rt_tilecull.cuh"

namespace rtb {

// tuning knobs (overridable at compile time for parameter sweeps: simd-raytracer_b200/build.py --define ... --out ...)
#ifndef RT_STREAM_BURST
#define RT_STREAM_BURST 8
#endif
#ifndef RT_STREAM_NODE_STEPS
#define RT_STREAM_NODE_STEPS 4
#endif
#ifndef RT_STREAM_LEAF_MIN
#define RT_STREAM_LEAF_MIN 8
#endif
#ifndef RT_STREAM_REFILL_BELOW
#define RT_STREAM_REFILL_BELOW 22
#endif
#ifndef RT_STREAM_LEAF_FRAC
#define RT_STREAM_LEAF_FRAC 3
#endif
#ifndef RT_STREAM_THREADS
#define RT_STREAM_THREADS 256
#endif
#ifndef RT_STREAM_MIN_BLOCKS
#define RT_STREAM_MIN_BLOCKS 3
#endif
#ifndef RT_STREAM4_THREADS
#define RT_STREAM4_THREADS 128
#endif
#ifndef RT_STREAM4_MIN_BLOCKS
#define RT_STREAM4_MIN_BLOCKS 5
#endif
#ifndef RT_STREAM4_STACK_START
#define RT_STREAM4_STACK_START 16
#endif
// four-wide traversal: stack entries per lane a scene starts with (fewer when the hierarchy's worst case is smaller; doubled by
// the host while queries overflow, see StreamStack<4>)
constexpr uint32_t STREAM4_STACK_START = RT_STREAM4_STACK_START;
// threads per block and resident blocks per SM of the stream kernels, by hierarchy width (registers per thread follow from them:
// two-wide 256 x 3 = 80 registers; four-wide 128 x 5 = 96 registers - its node step holds seven rows of a node at once)
template <int WIDE> struct StreamCfg { static constexpr int THREADS = RT_STREAM_THREADS, MIN_BLOCKS = RT_STREAM_MIN_BLOCKS; };
template <> struct StreamCfg<4> { static constexpr int THREADS = RT_STREAM4_THREADS, MIN_BLOCKS = RT_STREAM4_MIN_BLOCKS; };
constexpr int STREAM_BURST = RT_STREAM_BURST;                // rounds (node steps + one leaf phase) between completion phases
constexpr int STREAM_NODE_STEPS = RT_STREAM_NODE_STEPS;      // single-node steps per round; a lane that reaches a leaf parks until the leaf phase
constexpr int STREAM_LEAF_MIN = RT_STREAM_LEAF_MIN;          // run the leaf phase when this many lanes are parked ...
constexpr int STREAM_LEAF_FRAC = RT_STREAM_LEAF_FRAC;        // ... or when parked * this >= walking (0: only when nobody can walk on)
constexpr int STREAM_REFILL_BELOW = RT_STREAM_REFILL_BELOW;  // go and fetch new queries when fewer lanes than this are traversing

// RT_STREAM_STATS (developer builds only, scripts/gpu_stream_stats.py): where the lanes of a warp are while it executes node
// steps and leaf phases.  [0] node-step slots, [1] lanes walking in them, [2] lanes parked at a leaf, [3] lanes finished and
// waiting for the completion phase, [4] lanes without a query, [5] leaf phases, [6] lanes in them, [7] refill rounds,
// [8] completion phases, [9] bursts
#ifdef RT_STREAM_STATS
__device__ unsigned long long g_stream_stats[16];
// per-warp log of the LAST stream kernel that ran: { start ns, end ns, queries taken, node-step slots } (scripts/gpu_stream_log.py)
constexpr int STREAM_LOG_WARPS = 8192;
__device__ unsigned long long g_stream_log[STREAM_LOG_WARPS][4];
__device__ __forceinline__ unsigned long long stream_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define STREAM_STAT(i, v) (stats[i] += (v))
#else
#define STREAM_STAT(i, v) ((void)0)
#endif

// The reference-order re-run of a tied query is rare; keeping it out of line keeps its register needs out of the hot loop.
template <bool CULL, bool FAST>
__device__ __noinline__ void exact_rerun(const DScene& sc, bool active, float ox, float oy, float oz, float dx, float dy, float dz, float eps,
                                         Hit* out) {
    const Hit e = trace_warp<CULL, FAST>(sc, active, mk(ox, oy, oz), mk(dx, dy, dz), eps);
    if (active) *out = e;
}

// ---- the traversal stack of a lane, by hierarchy width -------------------------------------------------------------------------
// Two-wide (rt_bvh.cuh): 16-byte entries in thread-local memory, as in round 1.
// Four-wide (rt_bvh4.cuh): 8-byte entries in SHARED memory, one column per thread (consecutive lanes are consecutive 8-byte
// words, so a warp's access is two conflict-free wavefronts whatever the lanes' depths): no stack traffic reaches L1, which the
// node fetches need, and the pushes of a visit are predicated STS.64 with no branch between them.  The block's dynamic shared
// memory holds `rows` entries per lane.  The worst case of a hierarchy (host/bvh4_collapse.hpp) is 24-43 entries on the tested
// scenes, but every KB of shared memory is a KB less L1 for the node and triangle fetches and no recorded query of configs 1-4
// goes deeper than 14 (scripts/bvh_ray_lengths_cpu.py), so a scene starts with min(worst case, STREAM4_STACK_START) rows.  A
// query that would need more ends with the KD_OVERFLOW mark (rt_bvh4.cuh), is answered exactly by the reference-order traversal
// in the completion phase - and is counted: when more than one query in 2^14 of a frame overflowed, the host doubles the rows
// for the next frame, up to the worst case, where overflow is impossible (rays through the dense random mesh of config 5 go
// deep: that scene renders its second frame with the full stack).  Measured against a thread-local tail for the deep entries
// (one branch per visit): 0.386 vs 0.400 ms on config 2, 3.40 vs 3.52 on config 3, 29.1 vs 30.3 on config 5 at 1 M triangles.
template <int WIDE> struct StreamStack;
template <> struct StreamStack<2> {
    AccelStackEntry e[ACCEL_STACK];
    __device__ __forceinline__ void bind(uint32_t) {}
};
extern __shared__ uint2 stream_shared_stack[];
template <> struct StreamStack<4> {
    static constexpr uint32_t ROW = uint32_t(StreamCfg<4>::THREADS) * 8u;
    uint32_t col;                                // shared-memory address of this thread's column
    int rows;                                    // entries per lane
    __device__ __forceinline__ void bind(uint32_t n_rows) { col = uint32_t(__cvta_generic_to_shared(stream_shared_stack + threadIdx.x)); rows = int(n_rows); }
    __device__ __forceinline__ bool overflows(int entries) const { return entries > rows; }
    __device__ __forceinline__ void put_if(bool on, int pos, float t0, uint32_t child) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.shared.v2.b32 [%0], {%1, %2};\n\t}"
                     :: "r"(col + uint32_t(pos) * ROW), "r"(__float_as_uint(t0)), "r"(child), "r"(uint32_t(on)) : "memory");
    }
    __device__ __forceinline__ void get(int pos, float& t0, uint32_t& child) const {
        uint2 v;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(col + uint32_t(pos) * ROW) : "memory");
        t0 = __uint_as_float(v.x); child = v.y;
    }
};
template <int WIDE>
__device__ __forceinline__ void stream_node_step(AccelState& st, StreamStack<WIDE>& stack, const DScene& sc) {
    if constexpr (WIDE == 4) bvh4_node_step(st, stack, sc.w_nodes);
    else accel_node_step(st, stack.e, sc);
}
template <bool CULL, bool FAST, int WIDE>
__device__ __forceinline__ void stream_leaf_step(AccelState& st, StreamStack<WIDE>& stack, const DScene& sc, float eps) {
    if constexpr (WIDE == 4) bvh4_leaf_step<CULL, FAST>(st, stack, sc.b_tris, eps);
    else accel_leaf_step<CULL, FAST>(st, stack.e, sc, eps);
}

// Policy concept:
//   bool load(const DScene&, uint32_t& idx, V3& o, V3& d, float& t_far, bool& any_hit)  false: entry needs no query; may remap idx
//   void entered(uint32_t idx, V3 o, V3 d)                                                  the query touches the scene box and will be traced
//   bool finish(const DScene&, uint32_t idx, const Hit& h, AccelState& st)                  true: lane re-armed (st re-initialised)
template <bool CULL, bool FAST, int WIDE, class Policy>
__device__ __forceinline__ void stream_loop(const DScene& sc, Policy& p, uint32_t* __restrict__ counter, uint32_t end, float eps,
                                            unsigned long long* __restrict__ overflow_count) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t FULL = 0xFFFFFFFFu;
    AccelState st;
    StreamStack<WIDE> stack;
    stack.bind(sc.w_stack_rows);
    st.sp = 0; st.phase = KD8_DONE; st.any_hit = false; st.best.tri = -1; st.best.t = FLT_MAX; st.best.tie_t = -1.0f; st.t_far = FLT_MAX;
    st.ox = st.oy = st.oz = st.dx = st.dy = st.dz = 0.0f;
    bool busy = false, exhausted = false;
    uint32_t idx = 0, n_overflow = 0;
#ifdef RT_STREAM_STATS
    unsigned long long stats[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned long long log_t0 = stream_now_ns();
    unsigned long long log_taken = 0;
#endif
    for (;;) {
        // ---- refill idle lanes ----
#pragma unroll 1
        for (int round = 0; round < 8 && !exhausted; ++round) {
            const uint32_t idle = __ballot_sync(FULL, !busy);
            if (32 - __popc(idle) >= STREAM_REFILL_BELOW) break;
            STREAM_STAT(7, 1);
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == uint32_t(leader)) base = atomicAdd(counter, uint32_t(__popc(idle)));
            base = __shfl_sync(FULL, base, leader);
            if (base + uint32_t(__popc(idle)) >= end) exhausted = true;
            if (!busy) {
                const uint32_t mine = base + uint32_t(__popc(idle & ((1u << lane) - 1u)));
                if (mine < end) {
#ifdef RT_STREAM_STATS
                    ++log_taken;
#endif
                    idx = mine;
                    V3 o, d; float t_far; bool any_hit;
                    if (p.load(sc, idx, o, d, t_far, any_hit)) {
                        if (accel_init(st, sc, o.x, o.y, o.z, d.x, d.y, d.z, t_far, any_hit)) { busy = true; p.entered(idx, o, d); }
                        else {
                            Hit miss; miss.t = FLT_MAX; miss.u = 0.0f; miss.v = 0.0f; miss.tri = -1;
                            busy = p.finish(sc, idx, miss, st);
                        }
                    }
                }
            }
        }
        if (!__ballot_sync(FULL, busy)) {
            if (exhausted) break;
            continue;
        }
        STREAM_STAT(9, 1);
        // ---- burst: lanes walk inner nodes in lock step (one node per step, the same code for every lane); a lane that
        // reaches a leaf parks, and parked lanes test their leaves together - neither phase runs with a handful of lanes ----
#pragma unroll 1
        for (int it = 0; it < STREAM_BURST; ++it) {
#pragma unroll 1
            for (int k = 0; k < STREAM_NODE_STEPS; ++k) {
#ifdef RT_STREAM_STATS
                const int nw = __popc(__ballot_sync(FULL, busy && st.phase == KD8_WALK));
                if (nw) {
                    stats[0] += 1; stats[1] += nw;
                    stats[2] += __popc(__ballot_sync(FULL, busy && st.phase == KD8_LEAF));
                    stats[3] += __popc(__ballot_sync(FULL, busy && st.phase == KD8_DONE));
                    stats[4] += __popc(__ballot_sync(FULL, !busy));
                }
#endif
                if (busy && st.phase == KD8_WALK) stream_node_step<WIDE>(st, stack, sc);
            }
            const uint32_t parked = __ballot_sync(FULL, busy && st.phase == KD8_LEAF);
            const uint32_t walking = __ballot_sync(FULL, busy && st.phase == KD8_WALK);
            // the leaf phase runs when enough lanes are parked - in absolute terms, or relative to the lanes still walking: in the
            // tail of a launch a warp holds a handful of queries, and a parked one must not wait for eight
            if (parked && (__popc(parked) >= STREAM_LEAF_MIN || __popc(parked) * STREAM_LEAF_FRAC >= __popc(walking))) {
                STREAM_STAT(5, 1); STREAM_STAT(6, __popc(parked));
                if (busy && st.phase == KD8_LEAF) stream_leaf_step<CULL, FAST, WIDE>(st, stack, sc, eps);
            }
            const int running = __popc(__ballot_sync(FULL, busy && st.phase != KD8_DONE));
            if (running == 0 || (!exhausted && running < STREAM_REFILL_BELOW)) break;
        }
        // ---- completion (warp-uniform) ----
        STREAM_STAT(8, 1);
        const bool fin = busy && st.phase == KD8_DONE;
        Hit h; h.t = st.best.t; h.u = st.best.u; h.v = st.best.v; h.tri = st.best.tri;
        const bool tie = fin && (st.best.tri == KD_RERUN || st.best.tri == KD_OVERFLOW || (!st.any_hit && st.best.tri >= 0 && st.best.tie_t == st.best.t));
        n_overflow += (fin && st.best.tri == KD_OVERFLOW);
        if (__ballot_sync(FULL, tie)) {
            // two different triangles at exactly the winner's t: the reference's leaf order decides, so ask it
            exact_rerun<CULL, FAST>(sc, tie, st.ox, st.oy, st.oz, st.dx, st.dy, st.dz, eps, &h);
        }
        if (fin) busy = p.finish(sc, idx, h, st);
    }
    if (n_overflow) atomicAdd(overflow_count, (unsigned long long)n_overflow);      // rare: queries that outgrew the shared stack
#ifdef RT_STREAM_STATS
    for (int o = 16; o; o >>= 1) log_taken += __shfl_xor_sync(FULL, log_taken, o);
    if (lane == 0) {
        for (int i = 0; i < 10; ++i) atomicAdd(&g_stream_stats[i], stats[i]);
        const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (w < uint32_t(STREAM_LOG_WARPS)) { g_stream_log[w][0] = log_t0; g_stream_log[w][1] = stream_now_ns(); g_stream_log[w][2] = log_taken; g_stream_log[w][3] = stats[0]; }
    }
#endif
}

template <class T>
__device__ __forceinline__ void warp_sum_to(unsigned long long* a, unsigned long long* b, T na, T nb) {
    unsigned long long x = na, y = nb;
#pragma unroll
    for (int o = 16; o; o >>= 1) { x += __shfl_xor_sync(0xFFFFFFFFu, x, o); y += __shfl_xor_sync(0xFFFFFFFFu, y, o); }
    if ((threadIdx.x & 31u) == 0 && (x | y)) { atomicAdd(a, x); atomicAdd(b, y); }
}

// ---- primary rays ---------------------------------------------------------------------------------------------------------
struct PrimaryPolicy {
    const FrameParams* fp; Ray* rays; Hit* hits;
    uint32_t n_rays = 0, n_hits = 0;
    __device__ __forceinline__ bool load(const DScene& sc, uint32_t& i, V3& o, V3& d, float& t_far, bool& any_hit) {
        const uint32_t s = i / fp->plane, j = i - s * fp->plane;
        uint32_t x, y;
        if (!level0_pixel(*fp, j, x, y)) {
            Hit h; h.t = 0.0f; h.u = 0.0f; h.v = 0.0f; h.tri = TRI_INACTIVE;
            store_hit(hits + i, h);
            return false;
        }
        float rx, ry;
        primary_sample(sc, *fp, x, y, fp->sample_first + s, rx, ry, key);
        camera_ray(sc, fp->tan_half_fov, rx, ry, o, d);
        t_far = FLT_MAX; any_hit = false;
        ++n_rays;
        return true;
    }
    // the ray is only ever read back by the shading of a HIT (k_shade), so a camera ray that misses the scene box - most of
    // them on the dragon scenes - is never written to memory
    uint2 key;
    __device__ __forceinline__ void entered(uint32_t i, V3 o, V3 d) { store_ray(rays + i, o, d, key); }
    __device__ __forceinline__ bool finish(const DScene&, uint32_t i, const Hit& h, AccelState&) {
        store_hit(hits + i, h);
        n_hits += (h.tri >= 0);
        return false;
    }
};

template <bool FAST, int WIDE>
__global__ void __launch_bounds__(StreamCfg<WIDE>::THREADS, StreamCfg<WIDE>::MIN_BLOCKS) k_stream_primary(DScene sc, FrameParams fp, Ray* __restrict__ rays, Hit* __restrict__ hits,
                                                        PassState* __restrict__ ps, int work_slot) {
    pdl_wait();
    if (ps->skipped) return;
    PrimaryPolicy p; p.fp = &fp; p.rays = rays; p.hits = hits;
    stream_loop<true, FAST, WIDE>(sc, p, &ps->work[work_slot], fp.plane * fp.n_samples, fp.eps, &ps->pc.stack_overflows);                // render.hpp:64, culling ON
    warp_sum_to(&ps->pc.primary, &ps->pc.primary_hits, p.n_rays, p.n_hits);
}

// ---- primary rays, sparse level 0 ------------------------------------------------------------------------------------------
// A camera ray that hits nothing - 86 % of them on the dragon scenes - is finished the moment its query is.  In a one-sample
// pass that overwrites the framebuffer (first_pass) the lane writes the miss colour of render.hpp:66-74 ((0 + background) / spp)
// into the pixel and nothing else; in a multi-sample pass it writes nothing at all and k_accumulate adds the background for
// every sample whose bit is not set.  Only HITS become level-0 entries: ray and hit are stored at the entry's place in the
// sample plane as before, and the entry's bit is set in `mask0` (one word per 8x4 pixel tile), so the level-0 shade and resolve
// kernels skip every tile without a hit and never touch the entries of the misses.  Entries keep their pixel order (the
// shadow jobs and child rays they spawn stay coherent).  The ray is the traversal state's own (st.o*, st.d*).
struct SparsePrimaryPolicy {
    const FrameParams* fp; Ray* rays; Hit* hits; uint32_t* mask0;
    float* fb;                          // one-sample pass: misses write their pixel here; null in a multi-sample pass (k_accumulate adds them)
    const uint32_t* tile_list;          // tiles k_tile_cull kept, or null: every tile
    uint32_t per_sample;                // work items per sample = 32 * (tiles kept, or all tiles)
    V3 miss_rgb;
    uint32_t n_rays = 0, n_hits = 0;
    // work item -> level-0 entry: sample s, pixel (w & 31) of the (w >> 5)-th listed tile; entry = s * plane + tile * 32 + pixel
    __device__ __forceinline__ bool load(const DScene& sc, uint32_t& i, V3& o, V3& d, float& t_far, bool& any_hit) {
        const uint32_t s = i / per_sample, w = i - s * per_sample;
        const uint32_t j = tile_list ? __ldg(tile_list + (w >> 5)) * 32u + (w & 31u) : w;
        i = s * fp->plane + j;
        uint32_t x, y;
        if (!level0_pixel(*fp, j, x, y)) return false;                                                   // padding of the 8x4 tiles
        float rx, ry; uint2 key;
        primary_sample(sc, *fp, x, y, fp->sample_first + s, rx, ry, key);
        camera_ray(sc, fp->tan_half_fov, rx, ry, o, d);
        t_far = FLT_MAX; any_hit = false;
        ++n_rays;
        return true;
    }
    __device__ __forceinline__ void entered(uint32_t, V3, V3) {}
    __device__ __forceinline__ bool finish(const DScene& sc, uint32_t i, const Hit& h, AccelState& st) {
        const uint32_t s = i / fp->plane;
        uint32_t x, y;
        level0_pixel(*fp, i - s * fp->plane, x, y);
        if (h.tri >= 0) {
            float rx, ry; uint2 key;
            primary_sample(sc, *fp, x, y, fp->sample_first + s, rx, ry, key);                            // the path key of the sample
            store_ray(rays + i, mk(st.ox, st.oy, st.oz), mk(st.dx, st.dy, st.dz), key);
            store_hit(hits + i, h);
            atomicOr(mask0 + (i >> 5), 1u << (i & 31u));
            ++n_hits;
        } else if (fb) {
            float* px = fb + (size_t(y) * sc.width + x) * 3;
            px[0] = miss_rgb.x; px[1] = miss_rgb.y; px[2] = miss_rgb.z;
        }
        return false;
    }
};

// the operations k_resolve<true> performs for a miss of the first pass: (0 + background), divided once when the frame ends here
__device__ __forceinline__ V3 first_pass_miss_colour(const DScene& sc, const FrameParams& fp, int divide) {
    V3 m = mk(0.0f, 0.0f, 0.0f) + mk(sc.bg[0], sc.bg[1], sc.bg[2]);
    if (divide) { const float div = float(fp.spp_total); m = mk(__fdiv_rn(m.x, div), __fdiv_rn(m.y, div), __fdiv_rn(m.z, div)); }
    return m;
}

// ---- tile culling in front of the sparse primary kernel -----------------------------------------------------------------------
// rt_tilecull.cuh: all camera rays of an 8x4 pixel tile lie in the pyramid through the tile's four corners; if the scene's root
// box is outside one of its side planes (by a margin far above any rounding), every ray of the tile is a miss in the reference
// too - its first act is the slab test of the root box (kd_tree_simd.hpp:200-203).  Culled tiles get their miss colour here (or
// nothing, in a pass that k_accumulate finishes); the tiles that remain are listed for the stream kernel.
__global__ void __launch_bounds__(256) k_tile_cull(DScene sc, FrameParams fp, float* __restrict__ fb, int divide, PassState* __restrict__ ps,
                                                   uint32_t* __restrict__ tile_list) {
    pdl_wait();
    if (ps->skipped) return;
    const uint32_t n_tiles = fp.plane >> 5, lane = threadIdx.x & 31u, FULL = 0xFFFFFFFFu;
    const V3 miss = first_pass_miss_colour(sc, fp, divide);
    TileCamera cam;
    for (int k = 0; k < 9; ++k) cam.m[k] = sc.cam_m[k];
    for (int k = 0; k < 3; ++k) cam.pos[k] = sc.cam_pos[k];
    cam.width = float(sc.width); cam.height = float(sc.height); cam.tan_half_fov = float(fp.tan_half_fov);
    uint32_t n_culled = 0;
    for (uint32_t base = first_chunk(); base < n_tiles; base += chunk_stride()) {                        // 32 tiles per warp and round
        const uint32_t t = base + lane;
        uint32_t lx0 = 0, ly0 = 0, w = 0, h = 0;
        bool cull = false;
        if (t < n_tiles) {
            lx0 = (t % fp.tiles_x) * 8u; ly0 = (t / fp.tiles_x) * 4u;
            w = min(8u, fp.tw - lx0); h = min(4u, fp.th - ly0);
            const uint32_t fy0 = frame_row(fp, ly0);
            cull = tile_misses_box(cam, sc.root_min, sc.root_max, float(fp.x0 + lx0), float(fy0), float(fp.x0 + lx0 + w), float(fy0 + h));
        }
        const uint32_t keep = __ballot_sync(FULL, t < n_tiles && !cull);
        uint32_t slot = 0;
        if (lane == 0 && keep) slot = atomicAdd(&ps->n_tiles0, uint32_t(__popc(keep)));
        slot = __shfl_sync(FULL, slot, 0);
        if (t < n_tiles && !cull) tile_list[slot + uint32_t(__popc(keep & ((1u << lane) - 1u)))] = t;
        // culled tiles, one at a time, the warp's lanes being the tile's 32 pixels
        uint32_t todo = __ballot_sync(FULL, cull);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t px = __shfl_sync(FULL, lx0, src) + (lane & 7u), py = __shfl_sync(FULL, ly0, src) + (lane >> 3);
            const uint32_t tw = __shfl_sync(FULL, w, src), tht = __shfl_sync(FULL, h, src);
            if ((lane & 7u) < tw && (lane >> 3) < tht) {
                if (fb) {                                                           // null in a multi-sample pass: k_accumulate adds the misses
                    float* q = fb + (size_t(frame_row(fp, py)) * sc.width + (fp.x0 + px)) * 3;
                    q[0] = miss.x; q[1] = miss.y; q[2] = miss.z;
                }
                n_culled += fp.n_samples;
            }
        }
    }
    warp_sum_to(&ps->pc.primary, &ps->pc.primary_hits, n_culled, 0u);
}

template <bool FAST, int WIDE>
__global__ void __launch_bounds__(StreamCfg<WIDE>::THREADS, StreamCfg<WIDE>::MIN_BLOCKS) k_stream_primary_sparse(DScene sc, FrameParams fp, Ray* __restrict__ rays, Hit* __restrict__ hits,
                                                               uint32_t* __restrict__ mask0, float* __restrict__ fb, int divide,
                                                               PassState* __restrict__ ps, int work_slot, const uint32_t* __restrict__ tile_list) {
    pdl_wait();
    if (ps->skipped) return;
    SparsePrimaryPolicy p; p.fp = &fp; p.rays = rays; p.hits = hits; p.mask0 = mask0; p.fb = fb; p.tile_list = tile_list;
    p.per_sample = tile_list ? ps->n_tiles0 * 32u : fp.plane;
    p.miss_rgb = first_pass_miss_colour(sc, fp, divide);
    stream_loop<true, FAST, WIDE>(sc, p, &ps->work[work_slot], p.per_sample * fp.n_samples, fp.eps, &ps->pc.stack_overflows);             // render.hpp:64, culling ON
    warp_sum_to(&ps->pc.primary, &ps->pc.primary_hits, p.n_rays, p.n_hits);
}

// ---- level d >= 1 ------------------------------------------------------------------------------------------------------------
struct LevelPolicy {
    const Ray* rays; Hit* hits; uint32_t begin;
    uint32_t n_rays = 0, n_hits = 0;
    __device__ __forceinline__ bool load(const DScene&, uint32_t& i, V3& o, V3& d, float& t_far, bool& any_hit) {
        uint2 key;
        load_ray(rays + begin + i, o, d, key);
        t_far = FLT_MAX; any_hit = false;
        ++n_rays;
        return true;
    }
    __device__ __forceinline__ void entered(uint32_t, V3, V3) {}
    __device__ __forceinline__ bool finish(const DScene&, uint32_t i, const Hit& h, AccelState&) {
        store_hit(hits + begin + i, h);
        n_hits += (h.tri >= 0);
        return false;
    }
};

template <bool FAST, int WIDE>
__global__ void __launch_bounds__(StreamCfg<WIDE>::THREADS, StreamCfg<WIDE>::MIN_BLOCKS) k_stream_level(DScene sc, FrameParams fp, const Ray* __restrict__ rays, Hit* __restrict__ hits,
                                                      PassState* __restrict__ ps, int level, int work_slot) {
    pdl_wait();
    const uint32_t begin = ps->lv[level];
    const uint32_t end = min(ps->pool_count, fp.pool_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) ps->lv[level + 1] = end;
    LevelPolicy p; p.rays = rays; p.hits = hits; p.begin = begin;
    stream_loop<false, FAST, WIDE>(sc, p, &ps->work[work_slot], end - begin, fp.eps, &ps->pc.stack_overflows);
    warp_sum_to(&ps->pc.secondary, &ps->pc.secondary_hits, p.n_rays, p.n_hits);
}

// ---- shadow jobs: is_occluded, render/render.hpp:110-131 ---------------------------------------------------------------------------
template <bool TRANSMISSIVE>
struct ShadowPolicy {
    ShadowJob* jobs; float shadow_bias;
    uint32_t n_q = 0, n_h = 0;
    // Jobs are stored hit-major: the n_lights jobs of one shading point sit next to each other (k_shade), so 32 consecutive
    // jobs would send a warp towards n_lights different lights.  Work index -> job index transposes every full block of
    // 32 * n_lights jobs, so that consecutive work items are 32 neighbouring shading points and ONE light: coherent rays.
    uint32_t n_lights, n_jobs;
    __device__ __forceinline__ uint32_t job_of(uint32_t i) const {
        const uint32_t span = 32u * n_lights, block = i / span;
        if (n_lights <= 1u || (block + 1u) * span > n_jobs) return i;
        const uint32_t within = i - block * span;
        return block * span + (within & 31u) * n_lights + (within >> 5);
    }
    __device__ __forceinline__ bool load(const DScene&, uint32_t& i, V3& ro, V3& rd, float& t_far, bool& any_hit) {
        i = job_of(i);
        const float4* q = reinterpret_cast<const float4*>(jobs + i);
        const float4 a = q[0], b = q[1];
        if (!(0.0f < b.z)) return false;                                                                 // :115 - not occluded, no query
        ro = mk(a.x, a.y, a.z); rd = mk(a.w, b.x, b.y); t_far = b.z; any_hit = !TRANSMISSIVE;
        return true;
    }
    __device__ __forceinline__ void entered(uint32_t, V3, V3) {}
    // the query's origin, direction and remaining max_t are the traversal state's own (st.o*, st.d*, st.t_far)
    __device__ __forceinline__ bool finish(const DScene& sc, uint32_t i, const Hit& h, AccelState& st) {
        ++n_q;                                                                                           // :116 one closest-hit query
        if (h.tri < 0) return false;                                                                     // :117
        ++n_h;
        if (st.t_far < h.t) return false;                                                                // :117-119
        if (TRANSMISSIVE) {
            const uint32_t mat = __ldg(&sc.tri_index[h.tri]).w;
            if (sc.materials[mat].kind == 2u) {                                                          // :121-124 refractive: pass through
                const V3 d = mk(st.dx, st.dy, st.dz);
                const V3 pos = mk(st.ox, st.oy, st.oz) + h.t * d;
                const V3 o = pos + shadow_bias * d;                                                      // :126
                const float max_t = st.t_far - h.t;                                                      // :127
                if (!(0.0f < max_t)) return false;                                                       // :115
                if (accel_init(st, sc, o.x, o.y, o.z, d.x, d.y, d.z, max_t, false)) return true;
                ++n_q;                                                                                   // the next query misses the scene box
                return false;
            }
        }
        jobs[i].max_t = -1.0f;
        return false;
    }
};

template <bool TRANSMISSIVE, bool FAST, int WIDE>
__global__ void __launch_bounds__(StreamCfg<WIDE>::THREADS, StreamCfg<WIDE>::MIN_BLOCKS) k_stream_shadow(DScene sc, FrameParams fp, ShadowJob* __restrict__ jobs, PassState* __restrict__ ps,
                                                       int work_slot) {
    pdl_wait();
    const uint32_t end = min(ps->shadow_count, fp.shadow_cap);
    ShadowPolicy<TRANSMISSIVE> p; p.jobs = jobs; p.shadow_bias = fp.shadow_bias;
    p.n_lights = sc.n_lights; p.n_jobs = end;
    stream_loop<false, FAST, WIDE>(sc, p, &ps->work[work_slot], end, fp.eps, &ps->pc.stack_overflows);
    warp_sum_to(&ps->pc.shadow, &ps->pc.shadow_hits, p.n_q, p.n_h);
}

}  // namespace rtb
