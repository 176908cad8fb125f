// csrc/rt_api.cu - the extern "C" layer of include/rt_b200.h: scene upload, ray batches, frames, counters.
//
// Build (see simd-raytracer_b200/build.py): nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo.
// -fmad=false is part of the numerical contract (rt_device.cuh); nothing in this file may be compiled with
// --use_fast_math.  There is no CPU fallback: without a usable sm_100 device every compute entry point returns
// RT_ERR_NO_DEVICE.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <cerrno>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include "../../include/rt_b200.h"
#include "../host/kd_build.hpp"
#include "../host/scene.hpp"
#include "../host/bvh4_collapse.hpp"
#include "rt_stream.cuh"
#include "rt_peer.cuh"
#include "rt_lbvh.cuh"
#include "../host/kd_parallel.hpp"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thread>

using namespace rtb;

namespace {

thread_local std::string g_last_error;

int fail(int status, const std::string& what) {
    g_last_error = what;
    return status;
}

struct cuda_error { cudaError_t e; const char* what; };
#define CK(call)                                                   \
    do {                                                           \
        const cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) throw cuda_error{e_, #call};        \
    } while (0)

template <class F>
int guarded(F&& f) {
    try {
        return f();
    } catch (const rt_error& e) {
        return fail(e.status, e.what());
    } catch (const cuda_error& e) {
        const int st = (e.e == cudaErrorMemoryAllocation) ? RT_ERR_OOM
                       : (e.e == cudaErrorNoDevice || e.e == cudaErrorInsufficientDriver) ? RT_ERR_NO_DEVICE : RT_ERR_CUDA;
        cudaGetLastError();
        return fail(st, std::string(e.what) + ": " + cudaGetErrorString(e.e));
    } catch (const std::bad_alloc&) {
        return fail(RT_ERR_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(RT_ERR_BAD_ARG, e.what());
    }
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class T>
struct DBuf {                       // grow-only device buffer
    T* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr; cap = 0;
        CK(cudaMalloc(&p, n * sizeof(T)));
        cap = n;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
T* upload(const void* src, size_t n_elems, uint64_t& total) {
    T* d = nullptr;
    const size_t bytes = std::max<size_t>(n_elems, 1) * sizeof(T);
    CK(cudaMalloc(&d, bytes));
    if (n_elems) CK(cudaMemcpy(d, src, n_elems * sizeof(T), cudaMemcpyHostToDevice));
    total += bytes;
    return d;
}

enum TimeClass { TC_PRIMARY = 0, TC_SECONDARY, TC_SHADOW, TC_SHADE, TC_RESOLVE, TC_N };

}  // namespace

struct rt_scene {
    HostScene host;
    Geometry geom;
    KdTree tree;
    DeviceLayout layout;
    BvhLayout bvh_layout;            // the bounding-volume hierarchy (RT_FLAG_ORDERED, csrc/rt_bvh.cuh)
    std::vector<uint32_t> bvh4_nodes; // its four-wide collapse (csrc/rt_bvh4.cuh, host/bvh4_collapse.hpp); empty when accel_width == 2
    // the same structures built ON THE DEVICE (rt_build_opts.accel_build = 1, csrc/rt_lbvh.cuh): device arrays, owned by the scene
    struct DeviceBvh {
        uint32_t *nodes16 = nullptr, *tris12 = nullptr, *nodes32 = nullptr;
        uint64_t n_nodes2 = 0, n_nodes4 = 0, n_tris = 0, bytes = 0;
        uint32_t stack_need = 0, depth2 = 0;
        float root_min[3] = {0, 0, 0}, root_max[3] = {0, 0, 0};
        double seconds = 0;
        bool built = false;
    } dev_bvh;
    bool wide = false;               // the accelerated mode walks the four-wide nodes
    rt_scene_info info{};

    int device = RT_DEVICE_HOST_ONLY;
    int n_sm = 0;
    cudaStream_t stream = nullptr;
    DScene d{};
    std::vector<void*> owned;        // device allocations of the resident scene

    // wavefront pools (grow-only, reused across frames)
    // Everything a frame in flight writes on the device besides its framebuffer.  Two sets: the two frames of a sequence that are
    // in flight (rt_render_frame_begin) render on two streams into their own set, so that one frame's thinning launches - a
    // persistent grid ends when its longest queries end, profiles/r2_stream_tails.txt - run under the other frame's full ones.
    // `w` is the set the launch helpers below work on; every synchronous call uses set 0.
    struct Pools {
        DBuf<Ray> rays; DBuf<Hit> hits; DBuf<Rec> recs; DBuf<ShadowJob> jobs;
        DBuf<uint32_t> tiles0;       // sparse level 0: the tiles k_tile_cull kept
        DBuf<uint32_t> mask0;        // sparse level 0: one word per 8x4 pixel tile, bit = the camera ray hit something; zero between passes
        PassState* ps = nullptr;
        FrameCounters* fc = nullptr;
    } pools[2];
    Pools* w = &pools[0];
    cudaStream_t slot_stream[2] = {nullptr, nullptr};   // render streams of the two sequence slots (frames queued on the scene's own stream)
    uint32_t* h_flags = nullptr;     // pinned: 4 words per pass of the frame being rendered (k_pass_commit): pool_count, shadow_count, overflow, levels
    size_t h_flags_passes = 0;
    FrameCounters* h_fc = nullptr;   // pinned
    double pool_factor = 2.0, shadow_factor = 1.0;
    DBuf<float> fb; DBuf<uint8_t> fb8;
    DBuf<float> q_rays, q_maxt; DBuf<Hit> q_hits; DBuf<uint8_t> q_occ;
    // frame sequences (rt_render_frame_begin / rt_frame_wait): two frames in flight, each with its own device frame, its own
    // pinned completion words and counters; downloads run on their own stream
    struct SeqSlot {
        DBuf<float> fb;
        DBuf<uint8_t> fb8;               // rt_render_frame_rgb8_begin: the quantised frame the copy engine downloads
        uint8_t* host_rgb8 = nullptr;
        cudaEvent_t rendered = nullptr, copied = nullptr, a = nullptr, b = nullptr;
        uint32_t* h_flags = nullptr;     // pinned: what k_pass_commit published for every pass of this frame (4 words each)
        size_t h_flags_passes = 0;
        uint32_t n_passes = 0, per_pass = 0;
        FrameCounters* h_fc = nullptr;   // pinned
        bool in_flight = false, deferred = false, used = false;
        rt_params params{}, key{};
        float* host_rgb = nullptr;       // download target (frame sequences to the host), or null
        float* target = nullptr;         // device frame the slot renders into: its own `fb`, or the caller's (rt_render_frame_device_begin)
        cudaStream_t stream = nullptr;   // the stream the frame renders on
        cudaStream_t asked = nullptr;    // the stream the caller queued it on (the scene's own: `stream` is the slot's render stream)
        int pool = 0;                    // the pool set it uses
        bool rerendered = false;         // the frame overflowed its pools when first queued and was rendered again at wait time
        uint64_t n0 = 0;
        uint32_t launches = 0;
    } seq[2];
    cudaStream_t copy_stream = nullptr;
    uint64_t seq_issued = 0;         // tickets handed out so far (ticket t lives in slot t & 1)

    // counters of the last frame
    rt_counters counters{};
    struct Span { cudaEvent_t a, b; int cls; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    size_t events_used = 0;
    cudaEvent_t frame_a = nullptr, frame_b = nullptr, frame_c = nullptr;   // c: counters of the frame are in h_fc
    bool counters_pending = false;
    uint64_t pool_hwm = 0, shadow_hwm = 0;
    uint32_t launches = 0, passes = 0;
    // levels that held rays in the last pass rendered with `hint_params` (0 = unknown): deeper, empty levels are not launched
    uint32_t levels_hint = 0;
    rt_params hint_params{};

    std::mutex mtx;
    int g_primary[4] = {0, 0, 0, 0}, g_trace[4] = {0, 0, 0, 0}, g_shadow[8] = {0, 0, 0, 0, 0, 0, 0, 0}, g_shade[2] = {0, 0}, g_resolve = 0;
    int gs_primary[2] = {0, 0}, gs_sparse[2] = {0, 0}, gs_level[2] = {0, 0}, gs_shadow[4] = {0, 0, 0, 0};   // stream kernels (accelerated mode, the scene's width)

    ~rt_scene() {
        if (device >= 0) {
            cudaSetDevice(device);
            if (stream) cudaStreamSynchronize(stream);
            for (cudaStream_t t : slot_stream) if (t) cudaStreamSynchronize(t);
            for (void* p : owned) cudaFree(p);
            if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
            for (SeqSlot& q : seq) {
                for (cudaEvent_t e : {q.rendered, q.copied, q.a, q.b}) if (e) cudaEventDestroy(e);
                if (q.h_flags) cudaFreeHost(q.h_flags);
                if (q.h_fc) cudaFreeHost(q.h_fc);
                q.fb.release(); q.fb8.release();
            }
            for (Pools& q : pools) {
                q.rays.release(); q.hits.release(); q.recs.release(); q.jobs.release(); q.mask0.release(); q.tiles0.release();
                if (q.ps) cudaFree(q.ps);
                if (q.fc) cudaFree(q.fc);
            }
            for (cudaStream_t t : slot_stream) if (t) cudaStreamDestroy(t);
            fb.release(); fb8.release();
            q_rays.release(); q_maxt.release(); q_hits.release(); q_occ.release();
            if (h_flags) cudaFreeHost(h_flags);
            if (h_fc) cudaFreeHost(h_fc);
            for (cudaEvent_t e : event_pool) cudaEventDestroy(e);
            if (frame_a) cudaEventDestroy(frame_a);
            if (frame_b) cudaEventDestroy(frame_b);
            if (frame_c) cudaEventDestroy(frame_c);
            if (stream) cudaStreamDestroy(stream);
        }
    }
    cudaEvent_t next_event() {
        if (events_used == event_pool.size()) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            event_pool.push_back(e);
        }
        return event_pool[events_used++];
    }
};

namespace {

void require_device(const rt_scene* s) {
    if (!s) throw rt_error(RT_ERR_BAD_ARG, "null scene");
    if (s->device < 0) throw rt_error(RT_ERR_NO_DEVICE, "scene was built host-only (RT_DEVICE_HOST_ONLY): no device to compute on");
}

// the device the scene is to live on exists and is an sm_100 part; makes it current for the calling thread
void select_device(rt_scene* s) {
    int count = 0;
    const cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw rt_error(RT_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    if (s->device >= count) throw rt_error(RT_ERR_NO_DEVICE, "device ordinal out of range");
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, s->device));
    if (prop.major != 10) throw rt_error(RT_ERR_NO_DEVICE, std::string("device is not an sm_100 part: ") + prop.name);
    s->n_sm = prop.multiProcessorCount;
    CK(cudaSetDevice(s->device));
}

// ---- the backend's hierarchy built on the device (csrc/rt_lbvh.cuh) -----------------------------------------------------------
// Runs on its own host thread beside the host build of the reference's kd-tree (finish_create); temporaries are freed before it
// returns, the three result arrays stay with the scene.  Throws RT_ERR_UNSUPPORTED when the tree comes out deeper than the
// traversal stacks allow (degenerate inputs: thousands of coincident triangles) - the caller then builds on the host.
struct DevTmp {
    std::vector<void*> p;
    template <class T> T* get(size_t n) { void* q = nullptr; CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T))); p.push_back(q); return static_cast<T*>(q); }
    ~DevTmp() { for (void* q : p) cudaFree(q); }
};

void device_build_bvh(rt_scene* s, bool wide, uint32_t leaf) {
    const double t0 = now_s();
    select_device(s);
    const auto& tris = s->geom.tris;
    const uint64_t n64 = tris.size();
    if (n64 <= 4 * LBVH_MAX_LEAF || n64 >= (1ull << 29)) throw rt_error(RT_ERR_UNSUPPORTED, "device build: triangle count outside (16, 2^29)");
    leaf = std::min(std::max(leaf, 1u), LBVH_MAX_LEAF);
    const uint32_t n = uint32_t(n64), n_inner = n - 1;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    DevTmp tmp;
    rt_scene::DeviceBvh out;
    auto fail_free = [&] { for (uint32_t* q : {out.nodes16, out.tris12, out.nodes32}) if (q) cudaFree(q); };
    try {
        // v0, e1, e2 of every triangle: the only geometry the builder (and every triangle test) reads
        std::vector<float> tri9(size_t(n) * 9);
        parallel_for(n, 1 << 16, [&](uint64_t b, uint64_t e) {
            for (uint64_t i = b; i < e; ++i) { std::memcpy(&tri9[9 * i], tris[i].v0, 12); std::memcpy(&tri9[9 * i + 3], tris[i].e1, 12); std::memcpy(&tri9[9 * i + 6], tris[i].e2, 12); }
        });
        float* d_tri9 = tmp.get<float>(size_t(n) * 9);
        CK(cudaMemcpyAsync(d_tri9, tri9.data(), size_t(n) * 36, cudaMemcpyHostToDevice, st));
        float root6[6];
        std::memcpy(root6, s->geom.root_min, 12); std::memcpy(root6 + 3, s->geom.root_max, 12);
        float* d_root = tmp.get<float>(6);
        CK(cudaMemcpyAsync(d_root, root6, 24, cudaMemcpyHostToDevice, st));
        uint64_t *keys = tmp.get<uint64_t>(n), *keys_alt = tmp.get<uint64_t>(n);
        uint32_t *ids = tmp.get<uint32_t>(n), *ids_alt = tmp.get<uint32_t>(n);
        const unsigned B = 256, G = (n + B - 1) / B;
        k_lbvh_morton<<<G, B, 0, st>>>(d_tri9, n, d_root, keys, ids);
        CK(cudaGetLastError());
        cub::DoubleBuffer<uint64_t> kb(keys, keys_alt);
        cub::DoubleBuffer<uint32_t> vb(ids, ids_alt);
        size_t sort_bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, kb, vb, int(n), 0, 63, st));
        void* sort_tmp = tmp.get<uint8_t>(sort_bytes);
        CK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, kb, vb, int(n), 0, 63, st));
        const uint64_t* skeys = kb.Current();
        const uint32_t* sids = vb.Current();
        uint32_t *left = tmp.get<uint32_t>(n), *right = tmp.get<uint32_t>(n), *first = tmp.get<uint32_t>(n), *last = tmp.get<uint32_t>(n);
        uint32_t *parent_inner = tmp.get<uint32_t>(n), *parent_leaf = tmp.get<uint32_t>(n), *arrived = tmp.get<uint32_t>(n);
        LbvhBox *leaf_box = tmp.get<LbvhBox>(n), *inner_box = tmp.get<LbvhBox>(n);
        k_lbvh_tree<<<G, B, 0, st>>>(skeys, int(n), left, right, first, last, parent_inner, parent_leaf);
        CK(cudaGetLastError());
        CK(cudaMemsetAsync(arrived, 0, size_t(n) * 4, st));
        k_lbvh_boxes<<<G, B, 0, st>>>(d_tri9, sids, int(n), left, right, parent_inner, parent_leaf, leaf_box, inner_box, arrived);
        CK(cudaGetLastError());
        uint32_t *kept = tmp.get<uint32_t>(n), *dense = tmp.get<uint32_t>(n);
        k_lbvh_mark<<<G, B, 0, st>>>(first, last, int(n_inner), leaf, kept);
        CK(cudaGetLastError());
        size_t scan_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, kept, dense, int(n_inner), st));
        void* scan_tmp = tmp.get<uint8_t>(scan_bytes);
        CK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, kept, dense, int(n_inner), st));
        uint32_t tail[2] = {0, 0};
        CK(cudaMemcpyAsync(&tail[0], dense + (n_inner - 1), 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&tail[1], kept + (n_inner - 1), 4, cudaMemcpyDeviceToHost, st));
        LbvhBox root_box;
        CK(cudaMemcpyAsync(&root_box, inner_box, sizeof root_box, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint32_t n2 = tail[0] + tail[1];
        if (n2 == 0) throw rt_error(RT_ERR_UNSUPPORTED, "device build: empty hierarchy");
        float scale = 0.0f;
        for (int c = 0; c < 3; ++c) scale = std::max(scale, std::max(std::fabs(root_box.lo[c]), std::fabs(root_box.hi[c])));
        const float pad = 2e-5f * scale;                       // host/bvh_build.cpp flatten_bvh explains the padding
        CK(cudaMalloc(&out.nodes16, size_t(n2) * 64));
        CK(cudaMalloc(&out.tris12, size_t(n) * 48));
        k_lbvh_emit_nodes<<<G, B, 0, st>>>(left, right, first, last, kept, dense, leaf_box, inner_box, int(n_inner), pad, out.nodes16);
        CK(cudaGetLastError());
        k_lbvh_emit_tris<<<G, B, 0, st>>>(d_tri9, sids, n, out.tris12);
        CK(cudaGetLastError());
        // four-wide collapse, level by level (also measures the two-wide depth and the worst-case stack need)
        uint32_t* nodes32 = tmp.get<uint32_t>(size_t(n2) * 32);
        LbvhFrontier *fa = tmp.get<LbvhFrontier>(n2), *fb = tmp.get<LbvhFrontier>(n2);
        LbvhCounters* ctr = tmp.get<LbvhCounters>(1);
        LbvhCounters hc{1u, 0u, 0u, 0u};
        const LbvhFrontier rootf{0u, 0u, 0u, 0u};
        CK(cudaMemcpyAsync(fa, &rootf, sizeof rootf, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctr, &hc, sizeof hc, cudaMemcpyHostToDevice, st));
        uint32_t n_in = 1;
        for (int level = 0; n_in; ++level) {
            if (level > 256) throw rt_error(RT_ERR_UNSUPPORTED, "device build: hierarchy too deep");
            k_lbvh_collapse_level<<<(n_in + 127) / 128, 128, 0, st>>>(out.nodes16, fa, n_in, fb, n2, ctr, nodes32, n2);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&hc, ctr, sizeof hc, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            n_in = hc.next_count;
            const uint32_t zero = 0;
            CK(cudaMemcpyAsync(&ctr->next_count, &zero, 4, cudaMemcpyHostToDevice, st));
            std::swap(fa, fb);
        }
        if (hc.depth2 > 44 || hc.stack_need > uint32_t(BVH4_STACK) || hc.n_nodes4 > n2)
            throw rt_error(RT_ERR_UNSUPPORTED, "device build: the linear hierarchy is deeper than the traversal stacks allow");
        if (wide) {
            CK(cudaMalloc(&out.nodes32, size_t(hc.n_nodes4) * 128));
            CK(cudaMemcpyAsync(out.nodes32, nodes32, size_t(hc.n_nodes4) * 128, cudaMemcpyDeviceToDevice, st));
        }
        CK(cudaStreamSynchronize(st));
        out.n_nodes2 = n2; out.n_nodes4 = hc.n_nodes4; out.n_tris = n; out.stack_need = hc.stack_need; out.depth2 = hc.depth2;
        out.bytes = size_t(n2) * 64 + size_t(n) * 48 + (wide ? size_t(hc.n_nodes4) * 128 : 0);
        for (int c = 0; c < 3; ++c) { out.root_min[c] = root_box.lo[c] - pad; out.root_max[c] = root_box.hi[c] + pad; }
        out.seconds = now_s() - t0;
        out.built = true;
        s->dev_bvh = out;
    } catch (...) {
        fail_free();
        throw;
    }
}

void upload_scene(rt_scene* s) {
    const double t0 = now_s();
    select_device(s);
    CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));

    const DeviceLayout& L = s->layout;
    const HostScene& H = s->host;
    uint64_t bytes = 0;
    DScene& d = s->d;
    auto keep = [&](auto* p) { s->owned.push_back((void*)p); return p; };
    if (s->dev_bvh.built) {
        d.b_nodes = reinterpret_cast<const float*>(keep(s->dev_bvh.nodes16));
        d.b_tris = reinterpret_cast<const float*>(keep(s->dev_bvh.tris12));
        d.w_nodes = s->wide ? reinterpret_cast<const float*>(keep(s->dev_bvh.nodes32)) : nullptr;
        bytes += s->dev_bvh.bytes;
        std::memcpy(d.b_root_min, s->dev_bvh.root_min, 12);
        std::memcpy(d.b_root_max, s->dev_bvh.root_max, 12);
    } else {
        d.b_nodes = reinterpret_cast<const float*>(keep(upload<float4>(s->bvh_layout.nodes.data(), s->bvh_layout.nodes.size() / 4, bytes)));
        d.b_tris = reinterpret_cast<const float*>(keep(upload<float4>(s->bvh_layout.tris.data(), s->bvh_layout.tris.size() / 4, bytes)));
        d.w_nodes = s->wide ? reinterpret_cast<const float*>(keep(upload<float4>(s->bvh4_nodes.data(), s->bvh4_nodes.size() / 4, bytes))) : nullptr;
        std::memcpy(d.b_root_min, s->bvh_layout.root_min, 12);
        std::memcpy(d.b_root_max, s->bvh_layout.root_max, 12);
    }
    d.nodes32 = keep(upload<float4>(L.nodes32.data(), L.nodes32.size() / 4, bytes));
    d.packets = keep(upload<float4>(L.packets.data(), L.packets.size() / 4, bytes));
    d.tri_index = keep(upload<uint4>(L.tri_index.data(), L.tri_index.size() / 4, bytes));
    d.tri_normal = keep(upload<float4>(L.tri_normal.data(), L.tri_normal.size() / 4, bytes));
    d.tri_uv = keep(upload<float4>(L.tri_uv.data(), L.tri_uv.size() / 4, bytes));
    d.vnormals = keep(upload<float4>(L.vnormals.data(), L.vnormals.size() / 4, bytes));
    static_assert(sizeof(DMaterial) == sizeof(rt_material_desc) && sizeof(DTexture) == sizeof(rt_texture_desc) &&
                  sizeof(DLight) == sizeof(rt_light_desc), "device mirrors of the ABI structs");
    d.materials = keep(upload<DMaterial>(H.materials.data(), H.materials.size(), bytes));
    d.textures = keep(upload<DTexture>(H.textures.data(), H.textures.size(), bytes));
    d.lights = keep(upload<DLight>(H.lights.data(), H.lights.size(), bytes));
    d.texels = keep(upload<uint8_t>(H.texels.data(), H.texels.size(), bytes));
    d.n_lights = uint32_t(H.lights.size());
    d.n_nodes = uint32_t(s->tree.nodes.size());
    uint32_t start_rows = STREAM4_STACK_START;                       // RT_B200_STACK_ROWS: tests force the overflow path with it
    if (const char* e = std::getenv("RT_B200_STACK_ROWS")) std::sscanf(e, "%u", &start_rows);
    d.w_stack_rows = s->wide ? uint32_t(std::min<uint64_t>(s->info.bvh4_stack_need + 1, std::max(start_rows, 1u))) : 0u;
    d.width = H.width; d.height = H.height;
    std::memcpy(d.bg, H.background, 12);
    std::memcpy(d.cam_pos, H.camera_position, 12);
    std::memcpy(d.cam_m, H.camera_matrix, 36);
    std::memcpy(d.root_min, s->geom.root_min, 12);
    std::memcpy(d.root_max, s->geom.root_max, 12);
    d.has_transmissive = L.has_transmissive ? 1 : 0;

    for (rt_scene::Pools& q : s->pools) {
        CK(cudaMalloc(&q.ps, sizeof(PassState)));
        CK(cudaMalloc(&q.fc, sizeof(FrameCounters)));
    }
    CK(cudaMallocHost(&s->h_flags, 4 * sizeof(uint32_t)));
    s->h_flags_passes = 1;
    CK(cudaMallocHost(&s->h_fc, sizeof(FrameCounters)));
    CK(cudaEventCreate(&s->frame_a));
    CK(cudaEventCreate(&s->frame_b));
    CK(cudaEventCreateWithFlags(&s->frame_c, cudaEventDisableTiming));
    CK(cudaDeviceSynchronize());
    s->info.device_bytes = bytes;
    s->info.upload_seconds = now_s() - t0;
}

int finish_create(rt_scene* s, const rt_build_opts* opts, rt_scene** out) {
    rt_build_opts o;
    rt_default_build_opts(&o);
    if (opts) o = *opts;
    if (o.kd_max_depth > 30) throw rt_error(RT_ERR_BAD_ARG, "kd_max_depth > 30");
    if (o.kd_max_leaf_size == 0) throw rt_error(RT_ERR_BAD_ARG, "kd_max_leaf_size == 0");
    // the accelerated mode's hierarchy: its width (rt_build_opts.accel_width; RT_B200_ACCEL_WIDTH overrides the default for sweeps)
    // and where it is built (rt_build_opts.accel_build; RT_B200_ACCEL_BUILD likewise)
    // Triangles per leaf of the backend's hierarchy.  A scene that stays in L1 / L2 is bound by instruction issue and wants few,
    // full leaves (4: a node visit costs more instructions than a triangle test); a scene far beyond L2 pays a 48-byte fetch
    // from L2 / HBM for every triangle it tests and wants one triangle per leaf (config 5: -5 % at 10 M triangles with either
    // builder, profiles/r2_sweeps.txt).  The switch is the L2 capacity against ~176 bytes of hierarchy per triangle (126 MB:
    // ~750 K triangles).  RT_B200_BVH_LEAF overrides (sweeps).
    uint32_t leaf = s->host.triangle_count() * 176ull > (126ull << 20) ? 1u : 4u;
    if (const char* e = std::getenv("RT_B200_BVH_LEAF")) std::sscanf(e, "%u", &leaf);
    if (leaf == 0) throw rt_error(RT_ERR_BAD_ARG, "RT_B200_BVH_LEAF must be >= 1");
    uint32_t width = o.accel_width;
    if (!width) { width = RT_DEFAULT_ACCEL_WIDTH; if (const char* e = std::getenv("RT_B200_ACCEL_WIDTH")) std::sscanf(e, "%u", &width); }
    if (width != 2 && width != 4) throw rt_error(RT_ERR_BAD_ARG, "accel_width must be 0 (default), 2 or 4");
    if (width == 4 && leaf > BVH4_MAX_LEAF) throw rt_error(RT_ERR_BAD_ARG, "accel_width 4 holds at most 7 triangles per leaf");
    uint32_t where = o.accel_build;
    if (!where) if (const char* e = std::getenv("RT_B200_ACCEL_BUILD")) std::sscanf(e, "%u", &where);
    if (where > RT_ACCEL_BUILD_DEVICE) throw rt_error(RT_ERR_BAD_ARG, "accel_build must be RT_ACCEL_BUILD_HOST or RT_ACCEL_BUILD_DEVICE");
    if (where == RT_ACCEL_BUILD_DEVICE && o.device < 0) throw rt_error(RT_ERR_BAD_ARG, "accel_build = device needs a device");
    s->wide = width == 4;
    s->info.accel_width = width;
    s->info.bvh_leaf_size = leaf;
    s->device = o.device;
    s->info.device = o.device;
    const bool verbose = std::getenv("RT_B200_VERBOSE") != nullptr;

    double t0 = now_s();
    s->geom = prepare_geometry(s->host);
    const double t_prepared = now_s();
    // the device build runs beside the host build of the reference's kd-tree
    struct Side {
        std::thread th;
        std::exception_ptr err;
        ~Side() { if (th.joinable()) th.join(); }
    } side;
    if (where == RT_ACCEL_BUILD_DEVICE && s->geom.tris.size() > 4 * LBVH_MAX_LEAF)
        side.th = std::thread([&] { try { device_build_bvh(s, s->wide, leaf); } catch (...) { side.err = std::current_exception(); } });
    s->tree = build_kd_tree(s->geom, o.kd_max_depth, o.kd_max_leaf_size);
    s->info.build_seconds = now_s() - t0;
    t0 = now_s();
    s->layout = flatten(s->host, s->geom, s->tree);
    if (verbose) std::fprintf(stderr, "[rt_b200] geometry %.3f s, reference kd-tree %.3f s, its flattening %.3f s\n", t_prepared - (t0 - s->info.build_seconds),
                              s->info.build_seconds - (t_prepared - (t0 - s->info.build_seconds)), now_s() - t0);
    if (side.th.joinable()) side.th.join();
    if (side.err) {
        try { std::rethrow_exception(side.err); }
        catch (const rt_error& e) {
            if (e.status != RT_ERR_UNSUPPORTED) throw;
            if (verbose) std::fprintf(stderr, "[rt_b200] %s; building on the host\n", e.what());
        }
    }
    if (s->dev_bvh.built) {
        const auto& b = s->dev_bvh;
        s->info.accel_build = RT_ACCEL_BUILD_DEVICE;
        s->info.accel_build_seconds = b.seconds;
        s->info.bvh_n_nodes = b.n_nodes2; s->info.bvh_n_refs = b.n_tris;
        s->info.bvh_n_leaves = b.n_nodes2 + 1; s->info.bvh_depth = b.depth2;       // every two-wide node has exactly two children
        if (s->wide) { s->info.bvh4_n_nodes = b.n_nodes4; s->info.bvh4_stack_need = b.stack_need; }
        if (verbose) std::fprintf(stderr, "[rt_b200] device bvh build %.3f s: %llu two-wide nodes (depth %u), %llu four-wide (stack need %u)\n", b.seconds,
                                  (unsigned long long)b.n_nodes2, b.depth2, (unsigned long long)b.n_nodes4, b.stack_need);
    } else {
        const double t3 = now_s();
        KdTree bvh = build_bvh(s->geom, leaf);
        const double t4 = now_s();
        s->bvh_layout = flatten_bvh(s->geom, bvh);
        if (verbose) std::fprintf(stderr, "[rt_b200] bvh build %.3f s, flatten %.3f s\n", t4 - t3, now_s() - t4);
        s->info.accel_build = RT_ACCEL_BUILD_HOST;
        s->info.bvh_n_nodes = s->bvh_layout.n_nodes; s->info.bvh_n_refs = s->bvh_layout.n_refs;
        s->info.bvh_n_leaves = bvh.n_leaves; s->info.bvh_depth = bvh.depth;
        if (s->wide) {
            uint32_t need = 0;
            try { s->bvh4_nodes = bvh4_collapse(s->bvh_layout.nodes.data(), s->bvh_layout.n_nodes, &need); }
            catch (const std::length_error& e) { throw rt_error(RT_ERR_UNSUPPORTED, e.what()); }
            s->info.bvh4_n_nodes = s->bvh4_nodes.size() / 32;
            s->info.bvh4_stack_need = need;
            if (s->info.bvh4_stack_need > uint64_t(BVH4_STACK)) throw rt_error(RT_ERR_UNSUPPORTED, "four-wide hierarchy needs a deeper traversal stack than BVH4_STACK");
        }
        s->info.accel_build_seconds = now_s() - t3;
        if (verbose) std::fprintf(stderr, "[rt_b200] bvh build + flatten %.3f s\n", now_s() - t3);
    }
    s->info.flatten_seconds = now_s() - t0;
    s->info.width = s->host.width; s->info.height = s->host.height;
    s->info.n_triangles = s->geom.tris.size();
    s->info.n_vertices = s->geom.vertex_normals.size() / 3;
    s->info.n_nodes = s->tree.nodes.size();
    s->info.n_leaves = s->tree.n_leaves;
    s->info.n_leaf_refs = s->tree.refs.size();
    s->info.n_packets = s->layout.n_packets;
    s->info.max_leaf_refs = s->tree.max_leaf_refs;
    s->info.tree_depth = s->tree.depth;
    if (o.device >= 0) upload_scene(s);
    *out = s;
    return RT_OK;
}

template <class Loader>
int create_with(Loader&& load, const rt_build_opts* opts, rt_scene** out) {
    if (!out) return fail(RT_ERR_BAD_ARG, "null output handle");
    *out = nullptr;
    rt_scene* s = nullptr;
    const int st = guarded([&] {
        s = new rt_scene();
        s->host = load();
        return finish_create(s, opts, out);
    });
    if (st != RT_OK) { delete s; *out = nullptr; }
    return st;
}

// ---- launch helpers -------------------------------------------------------------------------------------------------
struct Mode { bool fast, ordered; };
Mode mode_of(uint32_t flags) { return Mode{(flags & RT_FLAG_FAST_MATH) != 0, (flags & RT_FLAG_ORDERED) != 0}; }

template <class K>
int grid_for(rt_scene* s, K kernel, int threads = 256, size_t dyn_smem = 0) {
    if (dyn_smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn_smem)));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem));
    return s->n_sm * std::max(per_sm, 1);
}
// the four-wide stream kernels keep every lane's traversal stack in shared memory: d.w_stack_rows entries per lane (rt_stream.cuh)
size_t wide_stack_bytes(const rt_scene* s) { return s->wide ? size_t(s->d.w_stack_rows) * StreamCfg<4>::THREADS * sizeof(uint2) : 0; }

#define DISPATCH_MODE(m, CALL)                                      \
    do {                                                            \
        if (!(m).fast && !(m).ordered) { CALL(false, false); }      \
        else if ((m).fast && !(m).ordered) { CALL(true, false); }   \
        else if (!(m).fast && (m).ordered) { CALL(false, true); }   \
        else { CALL(true, true); }                                  \
    } while (0)

void launch_trace_batch(rt_scene* s, const float* d_rays, uint64_t n, bool cull, float eps, Mode m, Hit* d_hits, cudaStream_t st) {
    if (!n) return;
    const int blocks = int(std::min<uint64_t>((n + 255) / 256, uint64_t(s->n_sm) * 64));
#define CALL(F, O)                                                                                          \
    if (cull) k_trace_batch<true, F, O><<<blocks, 256, 0, st>>>(s->d, d_rays, n, eps, d_hits);              \
    else k_trace_batch<false, F, O><<<blocks, 256, 0, st>>>(s->d, d_rays, n, eps, d_hits)
    DISPATCH_MODE(m, CALL);
#undef CALL
    CK(cudaGetLastError());
}

void launch_occluded_batch(rt_scene* s, const float* d_rays, const float* d_maxt, uint64_t n, float eps, float bias, Mode m,
                           uint8_t* d_out, cudaStream_t st) {
    if (!n) return;
    const int blocks = int(std::min<uint64_t>((n + 255) / 256, uint64_t(s->n_sm) * 64));
    const bool tr = s->d.has_transmissive != 0;
#define CALL(F, O)                                                                                                   \
    if (tr) k_occluded_batch<true, F, O><<<blocks, 256, 0, st>>>(s->d, d_rays, d_maxt, n, eps, bias, d_out);         \
    else k_occluded_batch<false, F, O><<<blocks, 256, 0, st>>>(s->d, d_rays, d_maxt, n, eps, bias, d_out)
    DISPATCH_MODE(m, CALL);
#undef CALL
    CK(cudaGetLastError());
}

// first_of_frame: a new frame starts here; otherwise a pass behind a discarded pass of the same frame is skipped (PassState::carry)
__global__ void k_pass_init(PassState* ps, uint32_t n0, int first_of_frame) {
    pdl_wait();
    const uint32_t skip = first_of_frame ? 0u : ps->carry;
    __syncthreads();
    uint32_t* w = reinterpret_cast<uint32_t*>(ps);
    for (uint32_t i = threadIdx.x; i < sizeof(PassState) / 4; i += blockDim.x) w[i] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        ps->skipped = skip; ps->carry = skip;
        if (!skip) { ps->pool_count = n0; ps->lv[1] = n0; }          // a skipped pass has no entries at any level and no shadow jobs
    }
}
// end of a pass: publish the pool usage to pinned host memory and, if the pass is kept, fold its ray counts
// `launched` of the `total` levels the frame can reach were traced and shaded (the host skips levels the previous pass of the
// same frame parameters left empty); if the last launched level spawned children after all, the pass is marked truncated
// (overflow bit 4) and the host renders it again with every level.  out[3] = levels that held entries.
// A pass skipped behind a discarded one reports overflow bit 8 and nothing else.
__global__ void k_pass_commit(PassState* ps, FrameCounters* fc, uint32_t* out, uint32_t launched, uint32_t total) {
    pdl_wait();
    if (ps->skipped) { out[0] = 0u; out[1] = 0u; out[2] = 8u; out[3] = 0u; return; }
    uint32_t used = 0;
    for (uint32_t d = 0; d < launched; ++d)
        if (ps->lv[d + 1] > ps->lv[d]) used = d + 1;
    if (launched < total && ps->pool_count > ps->lv[launched]) ps->overflow |= 4u;
    out[0] = ps->pool_count; out[1] = ps->shadow_count; out[2] = ps->overflow; out[3] = used;
    if (ps->overflow) { ps->carry = 1u; return; }
    fc->primary += ps->pc.primary; fc->primary_hits += ps->pc.primary_hits;
    fc->shadow += ps->pc.shadow; fc->shadow_hits += ps->pc.shadow_hits;
    fc->secondary += ps->pc.secondary; fc->secondary_hits += ps->pc.secondary_hits;
    fc->stack_overflows += ps->pc.stack_overflows;
}

struct Rect { uint32_t x0, y0, x1, y1; };

// rows of the frame a call with row bands renders (rt_params.band_rows): bands phase, phase + period, ... of band_rows rows each
uint32_t band_row_count(uint32_t height, const rt_params& p) {
    const uint32_t n_bands = (height + p.band_rows - 1) / p.band_rows;
    uint64_t rows = 0;
    for (uint32_t b = p.band_phase; b < n_bands; b += p.band_period) rows += std::min(p.band_rows, height - b * p.band_rows);
    return uint32_t(rows);
}

// entry points that deliver a host image or a hit array take a rectangle only
void forbid_bands(const rt_params& p) {
    if (p.band_rows) throw rt_error(RT_ERR_BAD_ARG, "row bands are rendered into a device frame (rt_render_frame_device, rt_render_frame_device_begin)");
}

Rect rect_of(const rt_scene* s, const rt_params& p) {
    if (p.band_rows) {
        if (p.band_rows % 4u || !p.band_period || p.band_phase >= p.band_period) throw rt_error(RT_ERR_BAD_ARG, "band_rows must be a multiple of 4 and band_phase < band_period");
        if (p.x0 | p.y0 | p.x1 | p.y1) throw rt_error(RT_ERR_BAD_ARG, "row bands and a tile rectangle exclude each other");
        if (!band_row_count(s->host.height, p)) throw rt_error(RT_ERR_BAD_ARG, "no row band of the frame has this phase");
        return Rect{0, 0, s->host.width, s->host.height};
    }
    Rect r{p.x0, p.y0, p.x1 ? p.x1 : s->host.width, p.y1 ? p.y1 : s->host.height};
    r.x1 = std::min(r.x1, s->host.width); r.y1 = std::min(r.y1, s->host.height);
    if (r.x0 >= r.x1 || r.y0 >= r.y1) throw rt_error(RT_ERR_BAD_ARG, "empty tile rectangle");
    return r;
}

FrameParams frame_params(const rt_scene* s, const rt_params& p, const Rect& r) {
    FrameParams fp{};
    // utils/convert.hpp:3-6 with F = double, then tan(fov_rad / 2) (render.hpp:55-57): host libm, as the reference
    const double fov_rad = p.fov_degrees * (3.14159265358979323846 / 180.);
    fp.tan_half_fov = std::tan(fov_rad / double(2.0f));
    fp.eps = p.epsilon; fp.shadow_bias = p.shadow_bias; fp.reflection_bias = p.reflection_bias; fp.refraction_bias = p.refraction_bias;
    fp.max_ray_depth = p.max_ray_depth; fp.gi_rays = p.diffuse_reflection_ray_count; fp.seed = p.seed;
    fp.spp_total = p.spp_total ? p.spp_total : p.samples_per_pixel;
    fp.x0 = r.x0; fp.y0 = r.y0; fp.tw = r.x1 - r.x0; fp.th = r.y1 - r.y0;
    if (p.band_rows) { fp.band_rows = p.band_rows; fp.band_period = p.band_period; fp.band_phase = p.band_phase; fp.th = band_row_count(s->host.height, p); }
    fp.tiles_x = (fp.tw + 7) / 8;
    const uint64_t plane = uint64_t(fp.tiles_x) * ((fp.th + 3) / 4) * 32;
    if (plane >= (1ull << 31)) throw rt_error(RT_ERR_BAD_ARG, "tile too large");
    fp.plane = uint32_t(plane);
    (void)s;
    return fp;
}

// The exact-mode pre-filter of the triangle test estimates 1/det with rcp.approx.ftz, which flushes a subnormal determinant to
// zero; with epsilon >= FLT_MIN the determinant test has rejected such a triangle before (kd_tree_simd.hpp:33-38), so the
// bit-exact contract holds for every epsilon from the smallest normal float up (the reference's is 1e-6, config.hpp:8).
void check_epsilon(float eps) {
    if (!(eps >= FLT_MIN)) throw rt_error(RT_ERR_BAD_ARG, "epsilon must be >= FLT_MIN (1.17549435e-38)");
}

void check_params(const rt_params& p) {
    check_epsilon(p.epsilon);
    if (p.samples_per_pixel == 0) throw rt_error(RT_ERR_BAD_ARG, "samples_per_pixel == 0");
    if (p.max_ray_depth > 64) throw rt_error(RT_ERR_BAD_ARG, "max_ray_depth > 64");
    if (p.diffuse_reflection_ray_count > 1024) throw rt_error(RT_ERR_BAD_ARG, "diffuse_reflection_ray_count > 1024");
}

// levels a frame can reach: without a reflective / refractive material and without GI only level 0 has rays
uint32_t level_count(const rt_scene* s, const rt_params& p) {
    bool spawns = p.diffuse_reflection_ray_count > 0;
    for (const auto& m : s->host.materials) spawns = spawns || m.kind == RT_MAT_REFLECTIVE || m.kind == RT_MAT_REFRACTIVE;
    return spawns ? p.max_ray_depth + 1 : 1;
}

// level-0 entries per pass.  A pass costs 80 B per entry and level (ray, hit, record) plus 32 B per shadow job - 2^25 entries are
// 5-8 GB of the 180 GB - and the more samples a pass holds, the smaller the share of its kernels' tails.
// RT_B200_PASS_ENTRIES overrides (tests force small passes with it).
uint64_t primary_budget() {
    static const uint64_t v = [] {
        unsigned long long e = 0;
        if (const char* t = std::getenv("RT_B200_PASS_ENTRIES")) std::sscanf(t, "%llu", &e);
        return e ? uint64_t(e) : (1ull << 25);
    }();
    return v;
}

void reserve_flags(uint32_t** flags, size_t* have, size_t passes) {
    if (passes <= *have) return;
    if (*flags) CK(cudaFreeHost(*flags));
    *flags = nullptr; *have = 0;
    CK(cudaMallocHost(flags, passes * 4 * sizeof(uint32_t)));
    *have = passes;
}

// persistent grid sizes of the kernels a frame in this mode launches (occupancy queries, once per scene)
void ensure_grids(rt_scene* s, Mode m, bool has_gi) {
    const int mi = (m.fast ? 1 : 0) | (m.ordered ? 2 : 0), fi = m.fast ? 1 : 0;
    const bool tr = s->d.has_transmissive != 0;
    if (!s->g_resolve) s->g_resolve = grid_for(s, k_resolve<false>);
    if (!s->g_shade[has_gi]) s->g_shade[has_gi] = has_gi ? grid_for(s, k_shade<true>) : grid_for(s, k_shade<false>);
    if (m.ordered) {
        // the stream kernels exist per hierarchy width; a scene only ever launches those of its own
#define STREAM_GRID(K2, K4) (s->wide ? grid_for(s, K4, StreamCfg<4>::THREADS, wide_stack_bytes(s)) : grid_for(s, K2, StreamCfg<2>::THREADS))
        if (!s->gs_primary[fi]) s->gs_primary[fi] = m.fast ? STREAM_GRID((k_stream_primary<true, 2>), (k_stream_primary<true, 4>)) : STREAM_GRID((k_stream_primary<false, 2>), (k_stream_primary<false, 4>));
        if (!s->gs_sparse[fi]) s->gs_sparse[fi] = m.fast ? STREAM_GRID((k_stream_primary_sparse<true, 2>), (k_stream_primary_sparse<true, 4>)) : STREAM_GRID((k_stream_primary_sparse<false, 2>), (k_stream_primary_sparse<false, 4>));
        if (!s->gs_level[fi]) s->gs_level[fi] = m.fast ? STREAM_GRID((k_stream_level<true, 2>), (k_stream_level<true, 4>)) : STREAM_GRID((k_stream_level<false, 2>), (k_stream_level<false, 4>));
        if (!s->gs_shadow[fi * 2 + tr])
            s->gs_shadow[fi * 2 + tr] = tr ? (m.fast ? STREAM_GRID((k_stream_shadow<true, true, 2>), (k_stream_shadow<true, true, 4>)) : STREAM_GRID((k_stream_shadow<true, false, 2>), (k_stream_shadow<true, false, 4>)))
                                           : (m.fast ? STREAM_GRID((k_stream_shadow<false, true, 2>), (k_stream_shadow<false, true, 4>)) : STREAM_GRID((k_stream_shadow<false, false, 2>), (k_stream_shadow<false, false, 4>)));
#undef STREAM_GRID
    } else {
        if (!s->g_primary[mi]) s->g_primary[mi] = m.fast ? grid_for(s, k_primary<true, false>) : grid_for(s, k_primary<false, false>);
        if (!s->g_trace[mi]) s->g_trace[mi] = m.fast ? grid_for(s, k_trace_level<true, false>) : grid_for(s, k_trace_level<false, false>);
        if (!s->g_shadow[mi * 2 + tr])
            s->g_shadow[mi * 2 + tr] = tr ? (m.fast ? grid_for(s, k_shadow<true, true, false>) : grid_for(s, k_shadow<true, false, false>))
                                          : (m.fast ? grid_for(s, k_shadow<false, true, false>) : grid_for(s, k_shadow<false, false, false>));
    }
}

// One pass = `fp.n_samples` samples of every pixel of the tile through the whole wavefront.
struct PassLaunch {
    FrameParams fp;
    Mode m;
    uint32_t levels, launched;        // levels the frame can reach / levels traced and shaded in this pass
    bool has_gi;
    int first_pass, divide;           // framebuffer: overwrite instead of add / divide by spp_total after adding
    int first_of_frame;               // first pass queued for this frame (or for its resumption): clears PassState::carry
};

// Queues the kernels of one pass on `st` (nothing here waits for the device).  launch(class, f) issues f() - the synchronous
// path wraps every launch in a pair of timing events, the frame-sequence path does not.  k_pass_commit publishes pool usage,
// the overflow bits and the level count of the pass to `h_flags` (pinned host memory, 4 words).
// Every kernel of a pass is launched with programmatic stream serialisation (programmatic dependent launch): the next kernel
// of the stream is set up and its blocks take the SM slots the current kernel's blocks leave in its tail, and they wait in
// pdl_wait() (rt_device.cuh, griddepcontrol.wait) until the current kernel has completed and its writes are visible.  What
// this removes is the launch gap between the 7-25 dependent kernels of a frame.  RT_B200_NO_PDL=1 turns it off (A/B runs).
bool pdl_enabled() {
    static const bool on = [] { const char* e = std::getenv("RT_B200_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}
template <class... KArgs, class... Args>
void launch_ks(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t dyn_smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = dyn_smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}
template <class... KArgs, class... Args>
void launch_k(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args&&... args) {
    launch_ks(kernel, grid, block, 0, st, std::forward<Args>(args)...);
}

// a stream kernel of the scene's hierarchy width, exact or fast arithmetic
#define STREAM_LAUNCH(K, GRID, ...)                                                                       \
    do {                                                                                                  \
        if (s->wide) { if (m.fast) launch_ks(K<true, 4>, GRID, StreamCfg<4>::THREADS, wide_stack_bytes(s), st, __VA_ARGS__);   \
                       else launch_ks(K<false, 4>, GRID, StreamCfg<4>::THREADS, wide_stack_bytes(s), st, __VA_ARGS__); }       \
        else { if (m.fast) launch_k(K<true, 2>, GRID, StreamCfg<2>::THREADS, st, __VA_ARGS__);            \
               else launch_k(K<false, 2>, GRID, StreamCfg<2>::THREADS, st, __VA_ARGS__); }                \
    } while (0)
#define STREAM_LAUNCH_T(K, T, GRID, ...)                                                                  \
    do {                                                                                                  \
        if (s->wide) { if (m.fast) launch_ks(K<T, true, 4>, GRID, StreamCfg<4>::THREADS, wide_stack_bytes(s), st, __VA_ARGS__);   \
                       else launch_ks(K<T, false, 4>, GRID, StreamCfg<4>::THREADS, wide_stack_bytes(s), st, __VA_ARGS__); }       \
        else { if (m.fast) launch_k(K<T, true, 2>, GRID, StreamCfg<2>::THREADS, st, __VA_ARGS__);         \
               else launch_k(K<T, false, 2>, GRID, StreamCfg<2>::THREADS, st, __VA_ARGS__); }             \
    } while (0)

template <class Launch>
void enqueue_pass(rt_scene* s, const PassLaunch& P, float* d_rgb, uint32_t* h_flags, cudaStream_t st, Launch&& launch) {
    FrameParams fp = P.fp;
    const Mode m = P.m;
    // accelerated mode: only the HITS of the camera rays become level-0 entries (rt_stream.cuh).  In a one-sample pass that
    // overwrites the framebuffer the misses write their pixels at once and level 0 resolves straight into the framebuffer
    // (fuse_acc); in every other pass k_accumulate adds the samples of a pixel in order, misses included
    fp.sparse0 = m.ordered ? 1u : 0u;
    const bool fuse_acc = fp.n_samples == 1 && (P.first_pass || !m.ordered);
    float* const miss_fb = fuse_acc ? d_rgb : nullptr;
    const int mi = (m.fast ? 1 : 0) | (m.ordered ? 2 : 0), fi = m.fast ? 1 : 0;
    const bool tr = s->d.has_transmissive != 0, has_gi = P.has_gi;
    const uint32_t launched = P.launched, levels = P.levels;
    const uint32_t n0 = fp.plane * fp.n_samples;
    int slot = 0;
    launch_k(k_pass_init, 1, 256, st, s->w->ps, n0, P.first_of_frame);
    CK(cudaGetLastError());
    // a camera outside the scene's root box: tiles whose rays cannot reach the box are finished by k_tile_cull (rt_stream.cuh)
    bool cull_tiles = false;
    if (fp.sparse0)
        for (int k = 0; k < 3; ++k) cull_tiles = cull_tiles || !(s->d.root_min[k] <= s->d.cam_pos[k] && s->d.cam_pos[k] <= s->d.root_max[k]);
    if (cull_tiles)
        launch(TC_PRIMARY, [&] {
            launch_k(k_tile_cull, std::min<unsigned>((fp.plane / 32 + 255) / 256, unsigned(s->g_resolve)), 256, st, s->d, fp, miss_fb, P.divide, s->w->ps, s->w->tiles0.p);
        });
    launch(TC_PRIMARY, [&] {
        if (fp.sparse0) {
            const uint32_t* tiles = cull_tiles ? s->w->tiles0.p : nullptr;
            STREAM_LAUNCH(k_stream_primary_sparse, s->gs_sparse[fi], s->d, fp, s->w->rays.p, s->w->hits.p, s->w->mask0.p, miss_fb, P.divide, s->w->ps, slot, tiles);
        } else if (m.ordered) {
            STREAM_LAUNCH(k_stream_primary, s->gs_primary[fi], s->d, fp, s->w->rays.p, s->w->hits.p, s->w->ps, slot);
        } else if (m.fast) launch_k(k_primary<true, false>, s->g_primary[mi], 256, st, s->d, fp, s->w->rays.p, s->w->hits.p, s->w->ps, slot);
        else launch_k(k_primary<false, false>, s->g_primary[mi], 256, st, s->d, fp, s->w->rays.p, s->w->hits.p, s->w->ps, slot);
    });
    ++slot;
    for (uint32_t lvl = 0; lvl < launched; ++lvl) {
        if (lvl > 0) {
            launch(TC_SECONDARY, [&] {
                if (m.ordered) {
                    STREAM_LAUNCH(k_stream_level, s->gs_level[fi], s->d, fp, s->w->rays.p, s->w->hits.p, s->w->ps, int(lvl), slot);
                } else if (m.fast) launch_k(k_trace_level<true, false>, s->g_trace[mi], 256, st, s->d, fp, s->w->rays.p, s->w->hits.p, s->w->ps, int(lvl), slot);
                else launch_k(k_trace_level<false, false>, s->g_trace[mi], 256, st, s->d, fp, s->w->rays.p, s->w->hits.p, s->w->ps, int(lvl), slot);
            });
            ++slot;
        }
        launch(TC_SHADE, [&] {
            if (has_gi) launch_k(k_shade<true>, s->g_shade[1], 256, st, s->d, fp, s->w->rays.p, s->w->hits.p, s->w->recs.p, s->w->jobs.p, s->w->ps, int(lvl), slot, s->w->mask0.p);
            else launch_k(k_shade<false>, s->g_shade[0], 256, st, s->d, fp, s->w->rays.p, s->w->hits.p, s->w->recs.p, s->w->jobs.p, s->w->ps, int(lvl), slot, s->w->mask0.p);
        });
        ++slot;
    }
    if (!s->host.lights.empty()) {
        launch(TC_SHADOW, [&] {
            if (m.ordered) {
                const int g = s->gs_shadow[fi * 2 + tr];
                if (tr) { STREAM_LAUNCH_T(k_stream_shadow, true, g, s->d, fp, s->w->jobs.p, s->w->ps, slot); }
                else { STREAM_LAUNCH_T(k_stream_shadow, false, g, s->d, fp, s->w->jobs.p, s->w->ps, slot); }
            } else {
                const int g = s->g_shadow[mi * 2 + tr];
                if (tr) { if (m.fast) launch_k(k_shadow<true, true, false>, g, 256, st, s->d, fp, s->w->jobs.p, s->w->ps, slot);
                          else launch_k(k_shadow<true, false, false>, g, 256, st, s->d, fp, s->w->jobs.p, s->w->ps, slot); }
                else { if (m.fast) launch_k(k_shadow<false, true, false>, g, 256, st, s->d, fp, s->w->jobs.p, s->w->ps, slot);
                       else launch_k(k_shadow<false, false, false>, g, 256, st, s->d, fp, s->w->jobs.p, s->w->ps, slot); }
            }
        });
        ++slot;
    }
    for (int lvl = int(launched) - 1; lvl >= 0; --lvl) {
        launch(TC_RESOLVE, [&] {
            if (lvl == 0 && fuse_acc)
                launch_k(k_resolve<true>, s->g_resolve, 256, st, s->d, fp, s->w->recs.p, s->w->jobs.p, s->w->ps, lvl, slot, d_rgb, P.first_pass, P.divide,
                                                              launched, levels, s->w->mask0.p);
            else
                launch_k(k_resolve<false>, s->g_resolve, 256, st, s->d, fp, s->w->recs.p, s->w->jobs.p, s->w->ps, lvl, slot, nullptr, 0, 0, launched, levels, s->w->mask0.p);
        });
        ++slot;
    }
    launch_k(k_pass_commit, 1, 1, st, s->w->ps, s->w->fc, h_flags, launched, levels);
    CK(cudaGetLastError());
    // the accumulate kernel skips itself on the device when the pass overflowed its pools
    if (!fuse_acc)
        launch(TC_RESOLVE, [&] {
            launch_k(k_accumulate, (fp.plane + 255) / 256, 256, st, s->d, fp, s->w->recs.p, d_rgb, s->w->ps, P.first_pass, P.divide, s->w->mask0.p);
        });
}

// wavefront pools for a pass of n0 level-0 entries, sized from the factors learnt so far; fills fp.pool_cap / fp.shadow_cap
void reserve_pools(rt_scene* s, FrameParams& fp, uint64_t n0, uint32_t levels) {
    const uint64_t pool_cap = std::max<uint64_t>(uint64_t(double(n0) * (levels > 1 ? s->pool_factor : 1.0)) + 1024, n0);
    const uint64_t shadow_cap = uint64_t(double(n0) * s->shadow_factor * std::max<size_t>(s->host.lights.size(), 1)) + 1024;
    if (pool_cap >= (1ull << 32) || shadow_cap >= (1ull << 32)) throw rt_error(RT_ERR_OOM, "wavefront pool exceeds 2^32 entries; lower samples per pass");
    // a pool that has to grow is freed and allocated again: frames still in flight (a queued frame of a sequence) use the old
    // one, so wait for them explicitly instead of leaning on cudaFree's implicit device synchronisation
    if (pool_cap > s->w->rays.cap || shadow_cap > s->w->jobs.cap || n0 / 32 + 1 > s->w->tiles0.cap || n0 / 32 + 1 > s->w->mask0.cap) CK(cudaDeviceSynchronize());
    s->w->rays.reserve(pool_cap); s->w->hits.reserve(pool_cap); s->w->recs.reserve(pool_cap); s->w->jobs.reserve(shadow_cap);
    s->w->tiles0.reserve(n0 / 32 + 1);
    if (s->w->mask0.cap < n0 / 32 + 1) {
        s->w->mask0.reserve(n0 / 32 + 1);
        CK(cudaMemset(s->w->mask0.p, 0, s->w->mask0.cap * sizeof(uint32_t)));
        CK(cudaDeviceSynchronize());                      // the render streams do not synchronise with the null stream
    }
    fp.pool_cap = uint32_t(std::min<uint64_t>(s->w->rays.cap, 0xFFFFFFFFull));
    fp.shadow_cap = uint32_t(std::min<uint64_t>(s->w->jobs.cap, 0xFFFFFFFFull));
}

// a truncated pass reports lower bounds of what it needed: grow past them (the pass is deterministic and is rendered again)
void grow_pools_after_overflow(rt_scene* s, const uint32_t* flags, uint64_t n0) {
    const uint64_t used_pool = flags[0], used_shadow = flags[1];
    if (flags[2] & 4u) s->levels_hint = 0;              // a skipped level was needed: all levels from now on
    if (flags[2] & 1u) s->pool_factor = std::max(s->pool_factor * 1.5, double(used_pool) / double(n0) * 1.25);
    if (flags[2] & 2u)
        s->shadow_factor = std::max(s->shadow_factor * 1.5, double(used_shadow) / double(n0 * std::max<size_t>(s->host.lights.size(), 1)) * 1.25);
}

void drain_sequence(rt_scene* s);

// Queries that outgrew the shared-memory stack of the four-wide stream kernels were answered by the reference-order traversal:
// exact, but slow.  More than one in 2^14 of a frame's queries: the next frame gets twice the rows (up to the hierarchy's worst
// case, where no query can overflow).  The grids are sized again, because the rows are the kernels' dynamic shared memory.
void adapt_stack_rows(rt_scene* s, const FrameCounters& c) {
    if (!s->wide || !c.stack_overflows) return;
    const uint64_t queries = c.primary + c.shadow + c.secondary, worst = s->info.bvh4_stack_need + 1;
    if (c.stack_overflows * 16384ull <= queries || s->d.w_stack_rows >= worst) return;
    s->d.w_stack_rows = uint32_t(std::min<uint64_t>(worst, uint64_t(s->d.w_stack_rows) * 2));
    for (int* g : {s->gs_primary, s->gs_sparse, s->gs_level}) g[0] = g[1] = 0;
    for (int& g : s->gs_shadow) g = 0;
    if (std::getenv("RT_B200_VERBOSE")) std::fprintf(stderr, "[rt_b200] %llu of %llu queries outgrew the shared stack: %u rows from now on\n",
                                                     (unsigned long long)c.stack_overflows, (unsigned long long)queries, s->d.w_stack_rows);
}

// counters of the last synchronously rendered frame: waits for it, reads the per-class CUDA events
void fetch_counters(rt_scene* s) {
    if (!s->counters_pending) return;
    CK(cudaSetDevice(s->device));
    CK(cudaEventSynchronize(s->frame_c));      // not a device-wide sync: a frame download may be in flight on the copy stream
    adapt_stack_rows(s, *s->h_fc);
    rt_counters k{};
    k.primary = s->h_fc->primary; k.primary_hits = s->h_fc->primary_hits;
    k.shadow = s->h_fc->shadow; k.shadow_hits = s->h_fc->shadow_hits;
    k.secondary = s->h_fc->secondary; k.secondary_hits = s->h_fc->secondary_hits;
    k.nodes_pool = s->pool_hwm; k.shadow_pool = s->shadow_hwm;
    k.kernel_launches = s->launches; k.passes = s->passes;
    CK(cudaEventElapsedTime(&k.ms_total, s->frame_a, s->frame_b));
    float by_class[TC_N] = {0, 0, 0, 0, 0};
    for (const auto& sp : s->spans) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, sp.a, sp.b));
        by_class[sp.cls] += ms;
    }
    k.ms_primary = by_class[TC_PRIMARY]; k.ms_secondary = by_class[TC_SECONDARY]; k.ms_shadow = by_class[TC_SHADOW];
    k.ms_shade = by_class[TC_SHADE]; k.ms_resolve = by_class[TC_RESOLVE];
    s->counters = k;
    s->counters_pending = false;
}

// One frame (or tile / sample slice of one) into a device framebuffer of height*width*3 floats.
void render_device(rt_scene* s, const rt_params& p, float* d_rgb, cudaStream_t st, bool drain = true) {
    check_params(p);
    const Rect rect = rect_of(s, p);
    FrameParams fp = frame_params(s, p, rect);
    const Mode m = mode_of(p.flags);
    const bool has_gi = fp.gi_rays > 0;
    const uint32_t levels = level_count(s, p);
    const bool raw = (p.flags & RT_FLAG_RAW_SUM) != 0;

    if (drain) { drain_sequence(s); s->w = &s->pools[0]; }   // frames of a sequence still in flight use the pools
    // finish the bookkeeping of the previous frame before its events are reused
    fetch_counters(s);
    s->events_used = 0; s->spans.clear();
    s->launches = 0; s->passes = 0; s->pool_hwm = 0; s->shadow_hwm = 0;
    CK(cudaMemsetAsync(s->w->fc, 0, sizeof(FrameCounters), st));
    CK(cudaEventRecord(s->frame_a, st));

    const uint32_t spp = p.samples_per_pixel;
    const uint32_t per_pass = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(spp, primary_budget() / fp.plane)));
    const uint32_t n_passes = (spp + per_pass - 1) / per_pass;
    // per-launch timing events only where they can be read back as a per-class split of ONE pass structure; a long multi-pass
    // frame is queued bare (the events would also switch programmatic dependent launch off between its kernels)
    const bool with_spans = n_passes <= 64;
    auto timed = [&](int cls, auto&& launch) {
        if (!with_spans) { launch(); CK(cudaGetLastError()); ++s->launches; return; }
        rt_scene::Span sp{s->next_event(), s->next_event(), cls};
        CK(cudaEventRecord(sp.a, st));
        launch();
        CK(cudaGetLastError());
        CK(cudaEventRecord(sp.b, st));
        s->spans.push_back(sp);
        ++s->launches;
    };
    ensure_grids(s, m, has_gi);

    // same frame parameters as the last frame: launch only the levels that held rays then (checked on device, see k_pass_commit)
    rt_params key = p;
    key.sample_offset = 0; key.samples_per_pixel = 0;
    if (std::memcmp(&key, &s->hint_params, sizeof key) != 0) { s->levels_hint = 0; s->hint_params = key; }
    uint32_t levels_seen = 0;

    // All passes of the frame are queued back to back; the host waits ONCE, at the end, and reads what every pass's
    // k_pass_commit published.  A pass that outgrew its pools was discarded on the device and the passes behind it skipped
    // themselves (PassState::carry), so the framebuffer holds exactly the passes in front of it: grow the pools, resume there.
    reserve_flags(&s->h_flags, &s->h_flags_passes, n_passes);
    uint32_t first = 0;                               // first pass that is not in the framebuffer yet
    for (int attempt = 0; first < n_passes; ++attempt) {
        if (attempt > 24) throw rt_error(RT_ERR_OOM, "wavefront pools keep overflowing");
        fp.n_samples = std::min(per_pass, spp);
        reserve_pools(s, fp, uint64_t(fp.plane) * fp.n_samples, levels);           // the largest pass; may grow (and synchronise) here only
        for (uint32_t k = first; k < n_passes; ++k) {
            const uint32_t done = k * per_pass, ns = std::min(per_pass, spp - done);
            fp.n_samples = ns;
            fp.sample_first = p.sample_offset + done;
            PassLaunch P{fp, m, levels, s->levels_hint ? std::min(levels, s->levels_hint) : levels, has_gi,
                         done == 0 ? 1 : 0, (!raw && done + ns == spp) ? 1 : 0, k == first ? 1 : 0};
            enqueue_pass(s, P, d_rgb, s->h_flags + 4 * size_t(k), st, timed);
        }
        CK(cudaStreamSynchronize(st));
        uint32_t k = first;
        for (; k < n_passes && s->h_flags[4 * size_t(k) + 2] == 0; ++k) {
            const uint32_t* f = s->h_flags + 4 * size_t(k);
            s->pool_hwm = std::max<uint64_t>(s->pool_hwm, f[0]);
            s->shadow_hwm = std::max<uint64_t>(s->shadow_hwm, f[1]);
            levels_seen = std::max(levels_seen, f[3]);
            ++s->passes;
        }
        if (k < n_passes) {
            const uint32_t done = k * per_pass;
            grow_pools_after_overflow(s, s->h_flags + 4 * size_t(k), uint64_t(fp.plane) * std::min(per_pass, spp - done));
        }
        first = k;
    }
    if (s->levels_hint == 0 || levels_seen > s->levels_hint) s->levels_hint = std::max<uint32_t>(levels_seen, 1);
    CK(cudaEventRecord(s->frame_b, st));
    CK(cudaMemcpyAsync(s->h_fc, s->w->fc, sizeof(FrameCounters), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(s->frame_c, st));
    s->counters_pending = true;
}

// ---- frame sequences ------------------------------------------------------------------------------------------------------
// A frame of a sequence is queued without waiting for the device: kernels of frame i+1 are already in the stream while frame
// i runs, and the download of frame i (copy stream) overlaps frame i+1.  What the synchronous path learns at its end-of-pass
// sync - did the pools overflow, which levels held rays - is read when the frame is waited for; an overflowed frame is then
// rendered again through the synchronous path (deterministic, so the caller sees the same frame, later).

void copy_rect_to_host(rt_scene* s, const Rect& r, const float* d_rgb, float* rgb, cudaStream_t st) {
    const size_t at = (size_t(r.y0) * s->host.width + r.x0) * 3;
    if (r.x0 == 0 && r.x1 == s->host.width) {                  // whole rows: one contiguous transfer
        CK(cudaMemcpyAsync(rgb + at, d_rgb + at, size_t(s->host.width) * 12 * (r.y1 - r.y0), cudaMemcpyDeviceToHost, st));
        return;
    }
    CK(cudaMemcpy2DAsync(rgb + at, size_t(s->host.width) * 12, d_rgb + at, size_t(s->host.width) * 12, size_t(r.x1 - r.x0) * 12,
                         r.y1 - r.y0, cudaMemcpyDeviceToHost, st));
}

void copy_rect8_to_host(rt_scene* s, const Rect& r, const uint8_t* d_rgb8, uint8_t* rgb8, cudaStream_t st) {
    const size_t at = (size_t(r.y0) * s->host.width + r.x0) * 3;
    if (r.x0 == 0 && r.x1 == s->host.width) {
        CK(cudaMemcpyAsync(rgb8 + at, d_rgb8 + at, size_t(s->host.width) * 3 * (r.y1 - r.y0), cudaMemcpyDeviceToHost, st));
        return;
    }
    CK(cudaMemcpy2DAsync(rgb8 + at, size_t(s->host.width) * 3, d_rgb8 + at, size_t(s->host.width) * 3, size_t(r.x1 - r.x0) * 3,
                         r.y1 - r.y0, cudaMemcpyDeviceToHost, st));
}
// io/image/ppm.hpp:17-19 on the rows of the tile rectangle, on the render stream
void quantise_rows(rt_scene* s, const Rect& r, const float* d_rgb, uint8_t* d_rgb8, cudaStream_t st) {
    const size_t first = size_t(r.y0) * s->host.width * 3, count = size_t(r.y1 - r.y0) * s->host.width * 3;
    k_quantise<<<unsigned((count + 255) / 256), 256, 0, st>>>(d_rgb + first, d_rgb8 + first, count);
    CK(cudaGetLastError());
}

void finalize_slot(rt_scene* s, int k) {
    rt_scene::SeqSlot& q = s->seq[k];
    if (!q.in_flight) return;
    CK(cudaEventSynchronize(q.copied));
    q.in_flight = false;
    if (!q.deferred) return;                                   // rendered by the synchronous path: nothing left to check
    uint32_t bad = 0, levels_used = 0;
    uint64_t pool_hwm = 0, shadow_hwm = 0;
    for (; bad < q.n_passes && q.h_flags[4 * size_t(bad) + 2] == 0; ++bad) {
        pool_hwm = std::max<uint64_t>(pool_hwm, q.h_flags[4 * size_t(bad)]);
        shadow_hwm = std::max<uint64_t>(shadow_hwm, q.h_flags[4 * size_t(bad) + 1]);
        levels_used = std::max(levels_used, q.h_flags[4 * size_t(bad) + 3]);
    }
    if (bad < q.n_passes) {
        grow_pools_after_overflow(s, q.h_flags + 4 * size_t(bad), q.n0);
        CK(cudaStreamSynchronize(q.stream));                   // (frames that share this set: the pools are about to grow)
        s->w = &s->pools[q.pool];
        render_device(s, q.params, q.target, q.stream, false);
        if (q.host_rgb) copy_rect_to_host(s, rect_of(s, q.params), q.target, q.host_rgb, q.stream);
        if (q.host_rgb8) {
            quantise_rows(s, rect_of(s, q.params), q.target, q.fb8.p, q.stream);
            copy_rect8_to_host(s, rect_of(s, q.params), q.fb8.p, q.host_rgb8, q.stream);
        }
        CK(cudaStreamSynchronize(q.stream));
        q.rerendered = true;
        return;
    }
    adapt_stack_rows(s, *q.h_fc);
    s->pool_hwm = pool_hwm; s->shadow_hwm = shadow_hwm;
    if (std::memcmp(&q.key, &s->hint_params, sizeof q.key) == 0 && (s->levels_hint == 0 || levels_used > s->levels_hint))
        s->levels_hint = std::max<uint32_t>(levels_used, 1);
    rt_counters c{};
    c.primary = q.h_fc->primary; c.primary_hits = q.h_fc->primary_hits;
    c.shadow = q.h_fc->shadow; c.shadow_hits = q.h_fc->shadow_hits;
    c.secondary = q.h_fc->secondary; c.secondary_hits = q.h_fc->secondary_hits;
    c.nodes_pool = s->pool_hwm; c.shadow_pool = s->shadow_hwm;
    c.kernel_launches = q.launches; c.passes = q.n_passes;
    CK(cudaEventElapsedTime(&c.ms_total, q.a, q.b));           // no per-class events inside a queued frame
    s->counters = c;
    s->counters_pending = false;
}

void drain_sequence(rt_scene* s) {
    if (!s->seq[0].in_flight && !s->seq[1].in_flight) return;
    const int older = int(s->seq_issued & 1);                  // the slot the next ticket would take holds the older frame
    finalize_slot(s, older);
    finalize_slot(s, older ^ 1);
}

// rgb / rgb8: host frame to download into (frame sequences) or null; d_ext: the caller's device frame or null (then the slot's own)
// RT_B200_SEQ_OVERLAP=0: the two frames of a sequence share one stream and one pool set (A/B runs)
bool sequence_overlap() {
    static const bool on = [] { const char* e = std::getenv("RT_B200_SEQ_OVERLAP"); return !(e && std::atoi(e) == 0); }();
    return on;
}

void begin_frame(rt_scene* s, const rt_params& p, float* rgb, float* d_ext, cudaStream_t st, uint64_t* ticket, uint8_t* rgb8 = nullptr) {
    check_params(p);
    if (rgb || rgb8) forbid_bands(p);
    const Rect rect = rect_of(s, p);
    // the frames a caller queues on ITS stream are ordered by that stream and share pool set 0: two of them in flight on
    // different streams would race on it
    for (const rt_scene::SeqSlot& o : s->seq)
        if (o.in_flight && o.asked != st) throw rt_error(RT_ERR_BAD_ARG, "all queued frames of a scene must use the same stream");
    const cudaStream_t asked = st;
    const int slot = int(s->seq_issued & 1);
    // frames queued on the scene's own stream are independent of anything the caller does: the two slots render on two streams
    // into two pool sets, and a frame's tail overlaps the next frame's start
    const bool overlap = st == s->stream && sequence_overlap();
    if (overlap) {
        if (!s->slot_stream[slot]) CK(cudaStreamCreateWithFlags(&s->slot_stream[slot], cudaStreamNonBlocking));
        st = s->slot_stream[slot];
    }
    if (!s->copy_stream) {
        CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        for (rt_scene::SeqSlot& q : s->seq) {
            CK(cudaEventCreateWithFlags(&q.rendered, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&q.copied, cudaEventDisableTiming));
            CK(cudaEventCreate(&q.a));
            CK(cudaEventCreate(&q.b));
            CK(cudaMallocHost(&q.h_flags, 4 * sizeof(uint32_t)));
            q.h_flags_passes = 1;
            CK(cudaMallocHost(&q.h_fc, sizeof(FrameCounters)));
        }
    }
    const int k = int(s->seq_issued & 1);
    rt_scene::SeqSlot& q = s->seq[k];
    finalize_slot(s, k);                                       // the frame two tickets back (normally long complete)
    s->w = &s->pools[overlap ? slot : 0];                      // (after it: a frame that has to be rendered again uses its own set)
    if (!d_ext) q.fb.reserve(size_t(s->host.width) * s->host.height * 3);
    if (rgb8) q.fb8.reserve(size_t(s->host.width) * s->host.height * 3);
    float* const target = d_ext ? d_ext : q.fb.p;
    q.rerendered = false;

    FrameParams fp = frame_params(s, p, rect);
    const uint32_t spp = p.samples_per_pixel;
    {
        // every pass of the frame is queued (render_device explains the device-side overflow protocol); what the passes
        // published is read at wait time, and a frame with a discarded pass is rendered again there
        fetch_counters(s);
        const Mode m = mode_of(p.flags);
        const bool has_gi = fp.gi_rays > 0;
        const uint32_t levels = level_count(s, p);
        ensure_grids(s, m, has_gi);
        rt_params key = p;
        key.sample_offset = 0; key.samples_per_pixel = 0;
        if (std::memcmp(&key, &s->hint_params, sizeof key) != 0) { s->levels_hint = 0; s->hint_params = key; }
        const uint32_t per_pass = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(spp, primary_budget() / fp.plane)));
        const uint32_t n_passes = (spp + per_pass - 1) / per_pass;
        reserve_flags(&q.h_flags, &q.h_flags_passes, n_passes);
        fp.n_samples = std::min(per_pass, spp);
        const uint64_t n0 = uint64_t(fp.plane) * fp.n_samples;
        reserve_pools(s, fp, n0, levels);
        uint32_t launches = 0;
        if (q.used) CK(cudaStreamWaitEvent(st, q.copied, 0));   // the device frame of this slot has been downloaded
        CK(cudaMemsetAsync(s->w->fc, 0, sizeof(FrameCounters), st));
        CK(cudaEventRecord(q.a, st));
        for (uint32_t k = 0; k < n_passes; ++k) {
            const uint32_t done = k * per_pass, ns = std::min(per_pass, spp - done);
            fp.n_samples = ns;
            fp.sample_first = p.sample_offset + done;
            PassLaunch P{fp, m, levels, s->levels_hint ? std::min(levels, s->levels_hint) : levels, has_gi, done == 0 ? 1 : 0,
                         (!(p.flags & RT_FLAG_RAW_SUM) && done + ns == spp) ? 1 : 0, k == 0 ? 1 : 0};
            enqueue_pass(s, P, target, q.h_flags + 4 * size_t(k), st, [&](int, auto&& launch) { launch(); CK(cudaGetLastError()); ++launches; });
        }
        CK(cudaEventRecord(q.b, st));
        CK(cudaMemcpyAsync(q.h_fc, s->w->fc, sizeof(FrameCounters), cudaMemcpyDeviceToHost, st));
        q.deferred = true; q.params = p; q.key = key; q.n0 = n0; q.launches = launches; q.n_passes = n_passes; q.per_pass = per_pass;
    }
    if (rgb8) quantise_rows(s, rect, target, q.fb8.p, st);     // 6.2 MB instead of 24.9 MB over PCIe for a 1080p frame
    if (rgb || rgb8) {
        CK(cudaEventRecord(q.rendered, st));
        CK(cudaStreamWaitEvent(s->copy_stream, q.rendered, 0));
        if (rgb) copy_rect_to_host(s, rect, target, rgb, s->copy_stream);
        if (rgb8) copy_rect8_to_host(s, rect, q.fb8.p, rgb8, s->copy_stream);
        CK(cudaEventRecord(q.copied, s->copy_stream));
    } else {
        CK(cudaEventRecord(q.copied, st));                     // "complete" = rendered
    }
    q.host_rgb = rgb; q.host_rgb8 = rgb8; q.target = target; q.stream = st; q.asked = asked; q.pool = overlap ? slot : 0; q.in_flight = true; q.used = true;
    *ticket = s->seq_issued++;
}

}  // namespace

// ======================================================================================================================
extern "C" {

int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

const char* rt_status_string(int status) {
    switch (status) {
        case RT_OK: return "ok";
        case RT_ERR_BAD_ARG: return "bad argument";
        case RT_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
        case RT_ERR_CUDA: return "CUDA runtime error";
        case RT_ERR_IO: return "i/o error";
        case RT_ERR_PARSE: return "parse error";
        case RT_ERR_OOM: return "out of memory";
        case RT_ERR_UNSUPPORTED: return "unsupported";
        case RT_ERR_TIMEOUT: return "a peer did not signal in time";
        case RT_FRAME_RERENDERED: return "frame was rendered again (its first, queued attempt outgrew the wavefront pools)";
        default: return "unknown status";
    }
}

const char* rt_last_error(void) { return g_last_error.c_str(); }

void rt_default_build_opts(rt_build_opts* o) {
    if (!o) return;
    o->kd_max_depth = 8; o->kd_max_leaf_size = 64; o->device = 0; o->accel_width = 0; o->accel_build = RT_ACCEL_BUILD_HOST;
}

void rt_default_params(rt_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof *p);
    p->fov_degrees = 90.;
    p->epsilon = float(1e-6);
    p->shadow_bias = float(1e-4); p->reflection_bias = float(1e-4); p->refraction_bias = float(1e-4);
    p->samples_per_pixel = 1; p->max_ray_depth = 5; p->diffuse_reflection_ray_count = 0; p->seed = 42;
}

int rt_scene_create(const rt_scene_desc* desc, const rt_build_opts* opts, rt_scene** out) {
    if (!desc) return fail(RT_ERR_BAD_ARG, "null scene description");
    return create_with([&] { return scene_from_desc(*desc); }, opts, out);
}

int rt_scene_create_from_crtscene(const char* path, const char* asset_root, const rt_build_opts* opts, rt_scene** out) {
    if (!path) return fail(RT_ERR_BAD_ARG, "null path");
    return create_with([&] { return scene_from_crtscene(path, asset_root ? asset_root : ""); }, opts, out);
}

int rt_scene_create_from_rtsc(const void* bytes, uint64_t n_bytes, const rt_build_opts* opts, rt_scene** out) {
    if (!bytes) return fail(RT_ERR_BAD_ARG, "null buffer");
    return create_with([&] { return scene_from_rtsc(bytes, n_bytes); }, opts, out);
}

void rt_scene_destroy(rt_scene* s) { delete s; }

int rt_scene_get_info(const rt_scene* s, rt_scene_info* info) {
    if (!s || !info) return fail(RT_ERR_BAD_ARG, "null argument");
    *info = s->info;
    return RT_OK;
}

int rt_scene_export_rtsc(const rt_scene* s, void* buf, uint64_t cap, uint64_t* n_bytes) {
    return guarded([&] {
        if (!s || !n_bytes) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        const std::vector<uint8_t> out = scene_to_rtsc(s->host);
        *n_bytes = out.size();
        if (buf && cap >= out.size()) std::memcpy(buf, out.data(), out.size());
        return int(RT_OK);
    });
}

int rt_scene_get_tree(const rt_scene* s, uint64_t* node5, float* boxes, uint32_t* refs) {
    if (!s) return fail(RT_ERR_BAD_ARG, "null scene");
    const auto& nodes = s->tree.nodes;
    for (size_t i = 0; i < nodes.size(); ++i) {
        const KdNode& n = nodes[i];
        if (node5) { node5[5 * i] = n.parent; node5[5 * i + 1] = n.child0; node5[5 * i + 2] = n.child1; node5[5 * i + 3] = n.first_ref; node5[5 * i + 4] = n.first_ref == KD_NONE ? 0 : n.ref_count; }
        if (boxes) { std::memcpy(boxes + 6 * i, n.bmin, 12); std::memcpy(boxes + 6 * i + 3, n.bmax, 12); }
    }
    if (refs && !s->tree.refs.empty()) std::memcpy(refs, s->tree.refs.data(), s->tree.refs.size() * 4);
    return RT_OK;
}

int rt_scene_get_device_layout(const rt_scene* s, uint32_t* nodes8, uint32_t* packets) {
    if (!s) return fail(RT_ERR_BAD_ARG, "null scene");
    if (nodes8) std::memcpy(nodes8, s->layout.nodes8.data(), s->layout.nodes8.size() * 4);
    if (packets) std::memcpy(packets, s->layout.packets.data(), size_t(s->layout.n_packets) * PACKET_WORDS * 4);
    return RT_OK;
}

int rt_scene_get_bvh_layout(const rt_scene* s, uint32_t* nodes16, uint32_t* tris12, float* root6) {
    if (!s) return fail(RT_ERR_BAD_ARG, "null scene");
    if (s->dev_bvh.built) {                      // built on the device: the arrays live there
        return guarded([&] {
            CK(cudaSetDevice(s->device));
            if (nodes16) CK(cudaMemcpy(nodes16, s->dev_bvh.nodes16, size_t(s->dev_bvh.n_nodes2) * 64, cudaMemcpyDeviceToHost));
            if (tris12) CK(cudaMemcpy(tris12, s->dev_bvh.tris12, size_t(s->dev_bvh.n_tris) * 48, cudaMemcpyDeviceToHost));
            if (root6) { std::memcpy(root6, s->geom.root_min, 12); std::memcpy(root6 + 3, s->geom.root_max, 12); }
            return int(RT_OK);
        });
    }
    if (nodes16) std::memcpy(nodes16, s->bvh_layout.nodes.data(), size_t(s->bvh_layout.n_nodes) * 16 * 4);
    if (tris12) std::memcpy(tris12, s->bvh_layout.tris.data(), size_t(s->bvh_layout.n_refs) * 12 * 4);
    if (root6) { std::memcpy(root6, s->geom.root_min, 12); std::memcpy(root6 + 3, s->geom.root_max, 12); }   // the reference's root box (bvh_init)
    return RT_OK;
}

int rt_scene_get_geometry(const rt_scene* s, float* tri9, float* face_normals, float* vertex_normals) {
    if (!s) return fail(RT_ERR_BAD_ARG, "null scene");
    const auto& tris = s->geom.tris;
    for (size_t i = 0; i < tris.size(); ++i) {
        if (tri9) { std::memcpy(tri9 + 9 * i, tris[i].v0, 12); std::memcpy(tri9 + 9 * i + 3, tris[i].e1, 12); std::memcpy(tri9 + 9 * i + 6, tris[i].e2, 12); }
        if (face_normals) std::memcpy(face_normals + 3 * i, tris[i].normal, 12);
    }
    if (vertex_normals && !s->geom.vertex_normals.empty())
        std::memcpy(vertex_normals, s->geom.vertex_normals.data(), s->geom.vertex_normals.size() * 4);
    return RT_OK;
}

int rt_trace_closest_device(rt_scene* s, const float* d_rays, uint64_t n, int backface_culling, float epsilon, uint32_t flags,
                            rt_hit* d_hits, void* stream) {
    return guarded([&] {
        require_device(s);
        if (n && (!d_rays || !d_hits)) throw rt_error(RT_ERR_BAD_ARG, "null buffer");
        check_epsilon(epsilon);
        CK(cudaSetDevice(s->device));
        static_assert(sizeof(rt_hit) == sizeof(Hit), "rt_hit layout");
        launch_trace_batch(s, d_rays, n, backface_culling != 0, epsilon, mode_of(flags), reinterpret_cast<Hit*>(d_hits),
                           stream ? static_cast<cudaStream_t>(stream) : s->stream);
        return RT_OK;
    });
}

int rt_trace_occluded_device(rt_scene* s, const float* d_rays, const float* d_max_t, uint64_t n, float epsilon, float shadow_bias,
                             uint32_t flags, uint8_t* d_occluded, void* stream) {
    return guarded([&] {
        require_device(s);
        if (n && (!d_rays || !d_max_t || !d_occluded)) throw rt_error(RT_ERR_BAD_ARG, "null buffer");
        check_epsilon(epsilon);
        CK(cudaSetDevice(s->device));
        launch_occluded_batch(s, d_rays, d_max_t, n, epsilon, shadow_bias, mode_of(flags), d_occluded,
                              stream ? static_cast<cudaStream_t>(stream) : s->stream);
        return RT_OK;
    });
}

int rt_trace_closest(rt_scene* s, const float* rays, uint64_t n, int backface_culling, float epsilon, uint32_t flags, rt_hit* hits) {
    return guarded([&] {
        require_device(s);
        if (n && (!rays || !hits)) throw rt_error(RT_ERR_BAD_ARG, "null buffer");
        check_epsilon(epsilon);
        if (!n) return int(RT_OK);
        std::lock_guard<std::mutex> lock(s->mtx);    // called concurrently by the reference's tile workers (render.hpp:93-101)
        CK(cudaSetDevice(s->device));
        s->q_rays.reserve(6 * n); s->q_hits.reserve(n);
        CK(cudaMemcpyAsync(s->q_rays.p, rays, 24 * n, cudaMemcpyHostToDevice, s->stream));
        launch_trace_batch(s, s->q_rays.p, n, backface_culling != 0, epsilon, mode_of(flags), s->q_hits.p, s->stream);
        CK(cudaMemcpyAsync(hits, s->q_hits.p, sizeof(Hit) * n, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return int(RT_OK);
    });
}

int rt_trace_occluded(rt_scene* s, const float* rays, const float* max_t, uint64_t n, float epsilon, float shadow_bias, uint32_t flags,
                      uint8_t* occluded) {
    return guarded([&] {
        require_device(s);
        if (n && (!rays || !max_t || !occluded)) throw rt_error(RT_ERR_BAD_ARG, "null buffer");
        check_epsilon(epsilon);
        if (!n) return int(RT_OK);
        std::lock_guard<std::mutex> lock(s->mtx);
        CK(cudaSetDevice(s->device));
        s->q_rays.reserve(6 * n); s->q_maxt.reserve(n); s->q_occ.reserve(n);
        CK(cudaMemcpyAsync(s->q_rays.p, rays, 24 * n, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(s->q_maxt.p, max_t, 4 * n, cudaMemcpyHostToDevice, s->stream));
        launch_occluded_batch(s, s->q_rays.p, s->q_maxt.p, n, epsilon, shadow_bias, mode_of(flags), s->q_occ.p, s->stream);
        CK(cudaMemcpyAsync(occluded, s->q_occ.p, n, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return int(RT_OK);
    });
}

int rt_render_frame_device(rt_scene* s, const rt_params* p, float* d_rgb, void* stream) {
    return guarded([&] {
        require_device(s);
        if (!p || !d_rgb) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        std::lock_guard<std::mutex> lock(s->mtx);
        CK(cudaSetDevice(s->device));
        render_device(s, *p, d_rgb, stream ? static_cast<cudaStream_t>(stream) : s->stream);
        return int(RT_OK);
    });
}

int rt_render_frame(rt_scene* s, const rt_params* p, float* rgb) {
    return guarded([&] {
        require_device(s);
        if (!p || !rgb) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        std::lock_guard<std::mutex> lock(s->mtx);
        CK(cudaSetDevice(s->device));
        const size_t n = size_t(s->host.width) * s->host.height * 3;
        s->fb.reserve(n);
        forbid_bands(*p);
        const Rect r = rect_of(s, *p);
        render_device(s, *p, s->fb.p, s->stream);
        // only the tile rectangle is defined on the device and only it is written to the caller's image
        const size_t row_bytes = size_t(r.x1 - r.x0) * 12;
        CK(cudaMemcpy2DAsync(rgb + (size_t(r.y0) * s->host.width + r.x0) * 3, size_t(s->host.width) * 12,
                             s->fb.p + (size_t(r.y0) * s->host.width + r.x0) * 3, size_t(s->host.width) * 12, row_bytes, r.y1 - r.y0,
                             cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return int(RT_OK);
    });
}

int rt_render_frame_rgb8(rt_scene* s, const rt_params* p, uint8_t* rgb8) {
    return guarded([&] {
        require_device(s);
        if (!p || !rgb8) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        if (p->flags & RT_FLAG_RAW_SUM) throw rt_error(RT_ERR_BAD_ARG, "RT_FLAG_RAW_SUM cannot be quantised");
        std::lock_guard<std::mutex> lock(s->mtx);
        CK(cudaSetDevice(s->device));
        const size_t n = size_t(s->host.width) * s->host.height * 3;
        s->fb.reserve(n); s->fb8.reserve(n);
        forbid_bands(*p);
        const Rect r = rect_of(s, *p);
        render_device(s, *p, s->fb.p, s->stream);
        const size_t first = size_t(r.y0) * s->host.width * 3, count = size_t(r.y1 - r.y0) * s->host.width * 3;
        k_quantise<<<unsigned((count + 255) / 256), 256, 0, s->stream>>>(s->fb.p + first, s->fb8.p + first, count);
        CK(cudaGetLastError());
        CK(cudaMemcpy2DAsync(rgb8 + (size_t(r.y0) * s->host.width + r.x0) * 3, size_t(s->host.width) * 3,
                             s->fb8.p + (size_t(r.y0) * s->host.width + r.x0) * 3, size_t(s->host.width) * 3, size_t(r.x1 - r.x0) * 3,
                             r.y1 - r.y0, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return int(RT_OK);
    });
}

int rt_render_frame_begin(rt_scene* s, const rt_params* p, float* rgb, uint64_t* ticket) {
    return guarded([&] {
        require_device(s);
        if (!p || !rgb || !ticket) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        std::lock_guard<std::mutex> lock(s->mtx);
        CK(cudaSetDevice(s->device));
        begin_frame(s, *p, rgb, nullptr, s->stream, ticket);
        return int(RT_OK);
    });
}

int rt_render_frame_rgb8_begin(rt_scene* s, const rt_params* p, uint8_t* rgb8, uint64_t* ticket) {
    return guarded([&] {
        require_device(s);
        if (!p || !rgb8 || !ticket) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        if (p->flags & RT_FLAG_RAW_SUM) throw rt_error(RT_ERR_BAD_ARG, "RT_FLAG_RAW_SUM cannot be quantised");
        std::lock_guard<std::mutex> lock(s->mtx);
        CK(cudaSetDevice(s->device));
        begin_frame(s, *p, nullptr, nullptr, s->stream, ticket, rgb8);
        return int(RT_OK);
    });
}

#ifdef RT_STREAM_STATS
// developer builds only (not declared in include/rt_b200.h): lane-occupancy counters of the stream kernels since the last reset
extern "C" __attribute__((visibility("default"))) int rt_debug_stream_stats(unsigned long long* out16, int reset) {
    if (out16 && cudaMemcpyFromSymbol(out16, rtb::g_stream_stats, sizeof(unsigned long long) * 16) != cudaSuccess) return RT_ERR_CUDA;
    if (reset) { unsigned long long z[16] = {}; if (cudaMemcpyToSymbol(rtb::g_stream_stats, z, sizeof z) != cudaSuccess) return RT_ERR_CUDA; }
    return RT_OK;
}
// per-warp log of the last stream kernel that ran: n_warps x { start ns, end ns, queries taken, node-step slots }
extern "C" __attribute__((visibility("default"))) int rt_debug_stream_log(unsigned long long* out, int n_warps) {
    if (cudaDeviceSynchronize() != cudaSuccess) return RT_ERR_CUDA;
    const size_t n = size_t(std::min(n_warps, rtb::STREAM_LOG_WARPS)) * 4 * sizeof(unsigned long long);
    return cudaMemcpyFromSymbol(out, rtb::g_stream_log, n) == cudaSuccess ? RT_OK : RT_ERR_CUDA;
}
#endif

void* rt_alloc_pinned(uint64_t bytes) {
    void* p = nullptr;
    const cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); fail(RT_ERR_OOM, std::string("cudaMallocHost: ") + cudaGetErrorString(e)); return nullptr; }
    return p;
}
void rt_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int rt_render_frame_device_begin(rt_scene* s, const rt_params* p, float* d_rgb, void* stream, uint64_t* ticket) {
    return guarded([&] {
        require_device(s);
        if (!p || !d_rgb || !ticket) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        std::lock_guard<std::mutex> lock(s->mtx);
        CK(cudaSetDevice(s->device));
        begin_frame(s, *p, nullptr, d_rgb, stream ? static_cast<cudaStream_t>(stream) : s->stream, ticket);
        return int(RT_OK);
    });
}

int rt_frame_wait(rt_scene* s, uint64_t ticket) {
    return guarded([&] {
        require_device(s);
        std::lock_guard<std::mutex> lock(s->mtx);
        if (ticket >= s->seq_issued) throw rt_error(RT_ERR_BAD_ARG, "no such frame ticket");
        CK(cudaSetDevice(s->device));
        // frames complete in ticket order; a ticket older than the two newest was finished when its slot was re-used
        if (ticket + 2 == s->seq_issued) finalize_slot(s, int(ticket & 1));
        else if (ticket + 1 == s->seq_issued) { finalize_slot(s, int((ticket & 1) ^ 1)); finalize_slot(s, int(ticket & 1)); }
        else return int(RT_OK);
        const rt_scene::SeqSlot& q = s->seq[ticket & 1];
        // a caller-owned device frame may already have been consumed (e.g. combined with other ranks) before the re-render
        return (q.rerendered && !q.host_rgb && !q.host_rgb8) ? int(RT_FRAME_RERENDERED) : int(RT_OK);
    });
}

int rt_trace_primary(rt_scene* s, const rt_params* p, rt_hit* hits) {
    return guarded([&] {
        require_device(s);
        if (!p || !hits) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        check_params(*p);
        forbid_bands(*p);
        std::lock_guard<std::mutex> lock(s->mtx);
        CK(cudaSetDevice(s->device));
        const Rect r = rect_of(s, *p);
        FrameParams fp = frame_params(s, *p, r);
        fp.n_samples = 1; fp.sample_first = p->sample_offset;
        const size_t n = size_t(fp.tw) * fp.th;
        s->q_hits.reserve(n);
        const Mode m = mode_of(p->flags);
#define CALL(F, O) k_primary_hits<F, O><<<(fp.plane + 255) / 256, 256, 0, s->stream>>>(s->d, fp, s->q_hits.p)
        DISPATCH_MODE(m, CALL);
#undef CALL
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(hits, s->q_hits.p, n * sizeof(Hit), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return int(RT_OK);
    });
}

int rt_get_counters(rt_scene* s, rt_counters* c) {
    return guarded([&] {
        require_device(s);
        if (!c) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        {
            std::lock_guard<std::mutex> lock(s->mtx);
            CK(cudaSetDevice(s->device));
            drain_sequence(s);
        }
        fetch_counters(s);
        *c = s->counters;
        return int(RT_OK);
    });
}

int rt_resolve_sum_device(rt_scene* s, const float* d_sum, uint32_t spp_total, float* d_rgb_out, uint8_t* d_rgb8_out, void* stream) {
    return guarded([&] {
        require_device(s);
        if (!d_sum || spp_total == 0) throw rt_error(RT_ERR_BAD_ARG, "bad argument");
        CK(cudaSetDevice(s->device));
        const size_t n = size_t(s->host.width) * s->host.height * 3;
        k_resolve_sum<<<unsigned((n + 255) / 256), 256, 0, stream ? static_cast<cudaStream_t>(stream) : s->stream>>>(
            d_sum, float(spp_total), d_rgb_out, d_rgb8_out, n);
        CK(cudaGetLastError());
        return int(RT_OK);
    });
}

}  // extern "C"

// ---- multi-GPU combine over NVLink peer memory (rt_peer.cuh) ---------------------------------------------------------------
struct rt_peer_group {
    uint32_t world = 1, rank = 0;
    int device = 0;
    uint64_t n = 0;                         // floats per framebuffer
    uint8_t* block = nullptr;               // local allocation: 2 x { [raw sums][result rgb][result rgb8] } + [flags]
    size_t slot_bytes = 0, off_rgb = 0, off_rgb8 = 0, off_flags = 0, bytes = 0;   // off_rgb / off_rgb8 are relative to a slot
    uint8_t* peer[PEER_MAX] = {};           // every rank's block as mapped here (own entry = block)
    bool opened[PEER_MAX] = {};
    bool connected = false;
    uint32_t epoch = 0;                     // frames signalled so far; frame e lives in slot (e - 1) & 1
    PeerTable table{};
    // shared host result: a POSIX shared-memory object every rank's process maps and pins; two frame slots like the device's
    float* host_map = nullptr;              // this process's mapping (2 x n floats)
    float* host_dev = nullptr;              // the same memory as this device addresses it
    size_t host_bytes = 0;
    uint32_t host_epoch = 0;                // last frame that was sent to the shared host frame
    uint8_t* stage = nullptr;               // copy-engine gather: this rank's slice of every OTHER rank's raw sums ((world - 1) x stage_stride bytes)
    size_t stage_stride = 0;
    // bounded waits (rt_peer.cuh PeerWait): a word in mapped host memory the wait kernels write when a peer's flag did not come
    uint32_t* err_host = nullptr;           // host view; 0 = no wait has timed out
    PeerWait wait{nullptr, 0};
    // throws RT_ERR_TIMEOUT once a wait kernel has reported a peer that never signalled (checked by every rt_peer_* call)
    void check_waits() const {
        const uint32_t e = err_host ? *reinterpret_cast<volatile const uint32_t*>(err_host) : 0u;
        if (e) throw rt_error(RT_ERR_TIMEOUT, "a peer did not signal frame " + std::to_string(e) + " in time (RT_B200_PEER_TIMEOUT_MS); the combined frame is not valid");
    }
    size_t next_slot() const { return size_t(epoch & 1u) * slot_bytes; }           // where the NEXT frame is rendered
    size_t last_slot() const { return size_t((epoch - 1u) & 1u) * slot_bytes; }    // the frame signalled last
};

namespace {
constexpr size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
void peer_fill_table(rt_peer_group* g) {
    for (uint32_t r = 0; r < g->world; ++r) {
        g->table.fb[r] = reinterpret_cast<const float*>(g->peer[r]);
        g->table.flags[r] = reinterpret_cast<uint32_t*>(g->peer[r] + g->off_flags);
    }
    g->connected = true;
}
}  // namespace

extern "C" {

int rt_peer_group_create(uint32_t world, uint32_t rank, int device, uint32_t width, uint32_t height, rt_peer_group** out, uint8_t* handle) {
    if (!out) return fail(RT_ERR_BAD_ARG, "null output handle");
    *out = nullptr;
    rt_peer_group* g = nullptr;
    const int st = guarded([&] {
        if (world == 0 || world > uint32_t(PEER_MAX) || rank >= world || !width || !height) throw rt_error(RT_ERR_BAD_ARG, "bad peer group shape");
        static_assert(sizeof(cudaIpcMemHandle_t) == RT_PEER_HANDLE_BYTES, "handle size");
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
            cudaGetLastError();
            throw rt_error(RT_ERR_NO_DEVICE, "no such CUDA device");
        }
        CK(cudaSetDevice(device));
        g = new rt_peer_group();
        g->world = world; g->rank = rank; g->device = device;
        g->n = uint64_t(width) * height * 3;
        g->off_rgb = align256(g->n * 4);
        g->off_rgb8 = g->off_rgb + align256(g->n * 4);
        g->slot_bytes = g->off_rgb8 + align256(g->n);
        g->off_flags = 2 * g->slot_bytes;
        g->bytes = g->off_flags + align256(PEER_FLAG_COUNT * 4);
        CK(cudaMalloc(&g->block, g->bytes));
        CK(cudaMemset(g->block, 0, g->bytes));
        CK(cudaHostAlloc(reinterpret_cast<void**>(&g->err_host), 64, cudaHostAllocMapped));
        *g->err_host = 0;
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g->wait.error), g->err_host, 0));
        unsigned timeout_ms = 30000;
        if (const char* e = std::getenv("RT_B200_PEER_TIMEOUT_MS")) std::sscanf(e, "%u", &timeout_ms);
        g->wait.timeout_ns = 1000000ull * timeout_ms;
        CK(cudaDeviceSynchronize());
        g->peer[rank] = g->block;
        if (handle) {
            cudaIpcMemHandle_t h;
            CK(cudaIpcGetMemHandle(&h, g->block));
            std::memcpy(handle, &h, sizeof h);
        }
        if (world == 1) peer_fill_table(g);
        *out = g;
        return int(RT_OK);
    });
    if (st != RT_OK) { if (g) { if (g->block) cudaFree(g->block); if (g->err_host) cudaFreeHost(g->err_host); delete g; } *out = nullptr; }
    return st;
}

int rt_peer_group_connect(rt_peer_group* g, const uint8_t* handles) {
    return guarded([&] {
        if (!g || !handles) throw rt_error(RT_ERR_BAD_ARG, "null argument");
        CK(cudaSetDevice(g->device));
        for (uint32_t r = 0; r < g->world; ++r) {
            if (r == g->rank || g->peer[r]) continue;
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles + size_t(r) * RT_PEER_HANDLE_BYTES, sizeof h);
            void* p = nullptr;
            CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            g->peer[r] = static_cast<uint8_t*>(p);
            g->opened[r] = true;
        }
        peer_fill_table(g);
        return int(RT_OK);
    });
}

int rt_peer_group_connect_local(rt_peer_group* const* groups, uint32_t world) {
    return guarded([&] {
        if (!groups || world == 0 || world > uint32_t(PEER_MAX)) throw rt_error(RT_ERR_BAD_ARG, "bad argument");
        for (uint32_t r = 0; r < world; ++r)
            if (!groups[r] || groups[r]->world != world || groups[r]->rank != r || groups[r]->n != groups[0]->n)
                throw rt_error(RT_ERR_BAD_ARG, "groups must be the ranks 0..world-1 of one shape");
        for (uint32_t a = 0; a < world; ++a) {
            rt_peer_group* g = groups[a];
            CK(cudaSetDevice(g->device));
            for (uint32_t r = 0; r < world; ++r) {
                if (r == a) continue;
                if (groups[r]->device != g->device) {
                    int can = 0;
                    CK(cudaDeviceCanAccessPeer(&can, g->device, groups[r]->device));
                    if (!can) throw rt_error(RT_ERR_UNSUPPORTED, "devices are not peers");
                    const cudaError_t e = cudaDeviceEnablePeerAccess(groups[r]->device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) throw cuda_error{e, "cudaDeviceEnablePeerAccess"};
                    cudaGetLastError();
                }
                g->peer[r] = groups[r]->block;
            }
            peer_fill_table(g);
        }
        return int(RT_OK);
    });
}

float* rt_peer_framebuffer(rt_peer_group* g) { return g ? reinterpret_cast<float*>(g->block + g->next_slot()) : nullptr; }
float* rt_peer_result_rgb(rt_peer_group* g) { return g && g->epoch ? reinterpret_cast<float*>(g->block + g->last_slot() + g->off_rgb) : nullptr; }
uint8_t* rt_peer_result_rgb8(rt_peer_group* g) { return g && g->epoch ? g->block + g->last_slot() + g->off_rgb8 : nullptr; }

int rt_peer_signal_ready(rt_peer_group* g, void* stream) {
    return guarded([&] {
        if (!g || !g->connected) throw rt_error(RT_ERR_BAD_ARG, "peer group is not connected");
        CK(cudaSetDevice(g->device));
        ++g->epoch;
        launch_ks(k_peer_signal_ready, 1, 32, 0, static_cast<cudaStream_t>(stream), g->table, int(g->world), int(g->rank), g->epoch);
        CK(cudaGetLastError());
        return int(RT_OK);
    });
}

int rt_peer_reduce_resolve(rt_peer_group* g, uint32_t spp_total, uint32_t outputs, void* stream) {
    return guarded([&] {
        if (!g || !g->connected || spp_total == 0) throw rt_error(RT_ERR_BAD_ARG, "bad argument");
        if (g->epoch == 0) throw rt_error(RT_ERR_BAD_ARG, "rt_peer_signal_ready has not been called for this frame");
        g->check_waits();
        if ((outputs & RT_PEER_OUT_HOST_RGB) && !g->host_dev) throw rt_error(RT_ERR_BAD_ARG, "RT_PEER_OUT_HOST_RGB without rt_peer_host_result_attach");
        CK(cudaSetDevice(g->device));
        const uint64_t n4 = g->n / 4;
        const uint64_t g0 = n4 * g->rank / g->world, g1 = n4 * (g->rank + 1) / g->world;
        int n_sm = 0;
        CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, g->device));
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        // How the peers' samples reach this rank.  "ce" (default): the rank's COPY ENGINES fetch its slice of every other rank's
        // framebuffer into a local staging buffer (one D2D copy per peer over NVLink, behind a one-block kernel that waits for
        // the READY flags), and the reduce kernel then reads local memory only - no SM waits on an NVLink round trip, which is
        // what slowed the render of the next frame the combine runs beside (N = 2, config 2: 422 -> 40x us per render with the
        // P2P-load kernel beside it).  "kernel": the kernel loads from the peers' memory itself (16 loads in flight per thread).
        // The results go to rank 0 by P2P stores either way.  RT_B200_PEER_GATHER / RT_B200_PEER_BLOCKS override (A/B runs).
        static const bool use_ce = [] { const char* e = std::getenv("RT_B200_PEER_GATHER"); return !(e && std::strcmp(e, "kernel") == 0); }();
        static const uint64_t max_blocks = [] { unsigned v = 0; if (const char* e = std::getenv("RT_B200_PEER_BLOCKS")) std::sscanf(e, "%u", &v); return uint64_t(v ? v : 128u); }();
        (void)n_sm;
        const bool last_rank = g->rank + 1 == g->world;
        const uint64_t f0 = g0 * 4, f1 = last_rank ? g->n : g1 * 4;               // this rank's float range, the frame's tail included
        PeerSources src{};
        const bool gather = use_ce && g->world > 1;
        if (gather) {
            if (!g->stage) {
                g->stage_stride = align256((g->n / g->world + 8) * sizeof(float));
                CK(cudaMalloc(&g->stage, g->stage_stride * (g->world - 1)));
            }
            k_peer_wait_ready<<<1, 32, 0, st>>>(g->table.flags[g->rank], int(g->world), g->epoch, g->wait);
            CK(cudaGetLastError());
            uint32_t k = 0;
            for (uint32_t r = 0; r < g->world; ++r) {
                const float* fb_r = reinterpret_cast<const float*>(g->peer[r] + g->last_slot());
                if (r == g->rank) { src.p[r] = fb_r; continue; }
                float* dst = reinterpret_cast<float*>(g->stage + size_t(k++) * g->stage_stride);
                CK(cudaMemcpyAsync(dst, fb_r + f0, (f1 - f0) * sizeof(float), cudaMemcpyDeviceToDevice, st));
                src.p[r] = dst - f0;                                             // indexed with the frame's element index
            }
        } else {
            for (uint32_t r = 0; r < g->world; ++r) src.p[r] = reinterpret_cast<const float*>(g->peer[r] + g->last_slot());
        }
        const int unroll = gather ? 8 : (g->world <= 2 ? 8 : g->world <= 4 ? 4 : g->world <= 8 ? 2 : 1);
        const uint64_t want = (g1 - g0 + uint64_t(256) * unroll - 1) / (uint64_t(256) * unroll);
        const int blocks = int(std::max<uint64_t>(1, std::min<uint64_t>(want, max_blocks)));
        uint8_t* root = g->peer[0] + g->last_slot();
#define PEER_REDUCE(U)                                                                                                                        \
        k_peer_reduce_resolve<U><<<blocks, 256, 0, st>>>(                                                                                     \
            g->table, src, gather ? 0 : 1, g->wait, int(g->world), int(g->rank), g0, g1, float(spp_total),                                            \
            (outputs & 1u) ? reinterpret_cast<float*>(root + g->off_rgb) : nullptr, (outputs & 2u) ? root + g->off_rgb8 : nullptr,            \
            g->epoch, n4 * 4, g->n, (outputs & RT_PEER_OUT_HOST_RGB) ? reinterpret_cast<float*>(g->block + g->last_slot() + g->off_rgb) : nullptr)
        if (unroll == 8) PEER_REDUCE(8); else if (unroll == 4) PEER_REDUCE(4); else if (unroll == 2) PEER_REDUCE(2); else PEER_REDUCE(1);
#undef PEER_REDUCE
        CK(cudaGetLastError());
        if (outputs & RT_PEER_OUT_HOST_RGB) {
            // this rank's slice: its own copy engine, its own PCIe link; then tell every rank (the consumer waits in rt_peer_wait_done)
            const uint64_t f0 = g0 * 4, f1 = (g->rank + 1 == g->world) ? g->n : g1 * 4;
            CK(cudaMemcpyAsync(g->host_map + size_t((g->epoch - 1u) & 1u) * g->n + f0,
                               reinterpret_cast<const float*>(g->block + g->last_slot() + g->off_rgb) + f0, (f1 - f0) * sizeof(float),
                               cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
            k_peer_signal_host<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(g->table, int(g->world), int(g->rank), g->epoch);
            g->host_epoch = g->epoch;
        }

        CK(cudaGetLastError());
        return int(RT_OK);
    });
}

int rt_peer_wait_done(rt_peer_group* g, void* stream) {
    return guarded([&] {
        if (!g || !g->connected) throw rt_error(RT_ERR_BAD_ARG, "peer group is not connected");
        g->check_waits();
        CK(cudaSetDevice(g->device));
        k_peer_wait_done<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(g->table.flags[g->rank], int(g->world), g->epoch, g->wait);
        CK(cudaGetLastError());
        if (g->host_epoch == g->epoch) {          // this frame also goes to the shared host frame: every rank's slice has landed
            k_peer_wait_host<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(g->table.flags[g->rank], int(g->world), g->epoch, g->wait);
            CK(cudaGetLastError());
        }
        return int(RT_OK);
    });
}

int rt_peer_combine(rt_peer_group* g, uint32_t spp_total, uint32_t outputs, void* stream) {
    int st = rt_peer_signal_ready(g, stream);
    if (st == RT_OK) st = rt_peer_reduce_resolve(g, spp_total, outputs, stream);
    if (st == RT_OK) st = rt_peer_wait_done(g, stream);
    return st;
}

int rt_peer_download_result(rt_peer_group* g, float* rgb, uint8_t* rgb8, void* stream) {
    return guarded([&] {
        if (!g || !g->epoch) throw rt_error(RT_ERR_BAD_ARG, "no combined frame yet");
        g->check_waits();
        CK(cudaSetDevice(g->device));
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (rgb) CK(cudaMemcpyAsync(rgb, g->block + g->last_slot() + g->off_rgb, g->n * 4, cudaMemcpyDeviceToHost, st));
        if (rgb8) CK(cudaMemcpyAsync(rgb8, g->block + g->last_slot() + g->off_rgb8, g->n, cudaMemcpyDeviceToHost, st));
        return int(RT_OK);
    });
}

int rt_peer_read_result(rt_peer_group* g, float* rgb, uint8_t* rgb8, void* stream) {
    const int st = rt_peer_download_result(g, rgb, rgb8, stream);
    if (st != RT_OK) return st;
    return guarded([&] {
        CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
        g->check_waits();                                   // a wait of THIS frame that ran out is reported here, not a frame later
        return int(RT_OK);
    });
}

int rt_peer_host_result_attach(rt_peer_group* g, const char* shm_name, int create, float** host_rgb) {
    return guarded([&] {
        if (!g || !shm_name || shm_name[0] != '/' || !host_rgb) throw rt_error(RT_ERR_BAD_ARG, "bad argument (the name must start with '/')");
        if (g->host_map) throw rt_error(RT_ERR_BAD_ARG, "a host result is already attached");
        CK(cudaSetDevice(g->device));
        const size_t bytes = (2 * g->n * sizeof(float) + 4095) & ~size_t(4095);
        const int fd = shm_open(shm_name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
        if (fd < 0) throw rt_error(RT_ERR_IO, std::string("shm_open ") + shm_name + ": " + std::strerror(errno));
        // posix_fallocate, not ftruncate: a /dev/shm that is too small must fail here (RT_ERR_IO), not with SIGBUS at the first store
        if (create) {
            const int fe = posix_fallocate(fd, 0, off_t(bytes));
            if (fe != 0) { close(fd); shm_unlink(shm_name); throw rt_error(RT_ERR_IO, std::string("posix_fallocate (is /dev/shm large enough?): ") + std::strerror(fe)); }
        }
        void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        const int e = errno;
        close(fd);
        if (p == MAP_FAILED) throw rt_error(RT_ERR_IO, std::string("mmap: ") + std::strerror(e));
        const cudaError_t ce = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
        if (ce != cudaSuccess) { munmap(p, bytes); throw cuda_error{ce, "cudaHostRegister"}; }
        void* d = nullptr;
        CK(cudaHostGetDevicePointer(&d, p, 0));
        g->host_map = static_cast<float*>(p); g->host_dev = static_cast<float*>(d); g->host_bytes = bytes;
        *host_rgb = g->host_map;
        return int(RT_OK);
    });
}

void rt_peer_group_destroy(rt_peer_group* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    if (g->host_map) { cudaHostUnregister(g->host_map); munmap(g->host_map, g->host_bytes); }
    for (uint32_t r = 0; r < g->world; ++r)
        if (g->opened[r] && g->peer[r]) cudaIpcCloseMemHandle(g->peer[r]);
    if (g->block) cudaFree(g->block);
    if (g->stage) cudaFree(g->stage);
    if (g->err_host) cudaFreeHost(g->err_host);
    delete g;
}

}  // extern "C"
