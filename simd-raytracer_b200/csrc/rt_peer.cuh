// csrc/rt_peer.cuh - the multi-GPU combine of an spp-sliced frame as ONE kernel over NVLink peer memory.
//
// the scene is replicated, every GPU renders its own samples of every pixel into its own
// framebuffer (RT_FLAG_RAW_SUM), and the frame is sum_over_ranks / spp followed by the PPM quantisation
// (render/render.hpp:66-74, io/image/ppm.hpp:17-19).  Instead of ncclReduce + a resolve kernel on the root, every rank runs
// k_peer_reduce_resolve: it waits (device side) until all peers have published "frame e rendered", reads ITS 1/world slice
// of every peer's framebuffer straight through NVLink (P2P loads on cudaIpc-mapped pointers), adds them in rank order - rank
// r holds sample slice r, so this is the reference's sample order and the result is bit-identical to the single-GPU frame
// when every rank renders one sample - divides, quantises, and stores the float and 8-bit slices into the root's result
// buffers (P2P stores).  With a shared host result attached (rt_peer_host_result_attach) the slice is also kept in the rank's
// OWN result buffer, from where the rank's copy engine moves it into the frame in HOST memory that every rank's process maps:
// N PCIe links carry the frame instead of rank 0's one (k_peer_signal_host / k_peer_wait_host order the consumer behind the
// copies; storing into host memory from the kernel itself was measured at a third of the copy engine's rate).  The last block publishes "rank r done with frame e" to every peer; k_peer_wait_done then orders the
// next frame behind everybody's reads.  No host round trip, no staging copy, one NVLink crossing per byte.
//
// Flags live in each rank's own block and are written remotely by peers with system-scope release stores.  Epochs only
// grow, so no flag is ever reset.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace rtb {

constexpr int PEER_MAX = 16;
constexpr int PEER_FLAG_READY = 0;            // flags[PEER_FLAG_READY + r] = last frame rank r has rendered
constexpr int PEER_FLAG_DONE = PEER_MAX;      // flags[PEER_FLAG_DONE + r]  = last frame rank r has finished reducing
constexpr int PEER_FLAG_HOST = 2 * PEER_MAX + 2;    // flags[PEER_FLAG_HOST + r]  = last frame whose slice rank r has copied into the shared host frame
constexpr int PEER_FLAG_COUNT = 3 * PEER_MAX + 2;   // [2*PEER_MAX] = block counter of the local reduce kernel

struct PeerTable {
    const float* fb[PEER_MAX];                // every rank's raw sample sums (height*width*3 floats), frame slot 0
    uint32_t* flags[PEER_MAX];                // every rank's flag block
};
// Two frame slots per rank (frame e lives in slot (e - 1) & 1): frame e+1 can be rendered while frame e is being reduced;
// a slot is rendered into again only after every rank has published DONE for the frame that used it (k_peer_wait_done).

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Every wait on a peer's flag is BOUNDED: a peer that died, or never signals, must not leave this rank's GPU spinning for ever.
// PeerWait::timeout_ns (rt_peer_group: RT_B200_PEER_TIMEOUT_MS, default 30 s; 0 = unbounded) is measured on %globaltimer; a wait
// that runs out writes the frame number into PeerWait::error - a word in mapped host memory the host side reads at its next
// rt_peer_* call (RT_ERR_TIMEOUT) - and returns, so the stream drains (with a frame that is not to be used).
struct PeerWait { uint32_t* error; unsigned long long timeout_ns; };
__device__ __forceinline__ unsigned long long peer_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void peer_wait_flag(const uint32_t* p, uint32_t epoch, PeerWait w, unsigned sleep_ns) {
    if (ld_acquire_sys(p) >= epoch) return;
    const unsigned long long t0 = peer_now_ns();
    while (ld_acquire_sys(p) < epoch) {
        __nanosleep(sleep_ns);
        if (w.timeout_ns && peer_now_ns() - t0 > w.timeout_ns) {
            if (w.error) { *reinterpret_cast<volatile uint32_t*>(w.error) = epoch; __threadfence_system(); }
            return;
        }
    }
}

// "my framebuffer holds frame `epoch`": one remote store per peer.  Launched with programmatic stream serialisation like the
// kernels of a pass (rt_api.cu launch_ks): it is set up while the frame's last kernel drains and waits here for its completion.
__global__ void k_peer_signal_ready(PeerTable t, int world, int rank, uint32_t epoch) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int i = threadIdx.x;
    if (i < world) st_release_sys(t.flags[i] + PEER_FLAG_READY + rank, epoch);
}

// "my slice of frame `epoch` is in the shared host frame": one remote store per peer, behind the copy in stream order
__global__ void k_peer_signal_host(PeerTable t, int world, int rank, uint32_t epoch) {
    const int i = threadIdx.x;
    if (i < world) { __threadfence_system(); st_release_sys(t.flags[i] + PEER_FLAG_HOST + rank, epoch); }
}
__global__ void k_peer_wait_host(const uint32_t* my_flags, int world, uint32_t epoch, PeerWait w) {
    const int i = threadIdx.x;
    if (i < world) peer_wait_flag(my_flags + PEER_FLAG_HOST + i, epoch, w, 64);
}

__global__ void k_peer_wait_done(const uint32_t* my_flags, int world, uint32_t epoch, PeerWait w) {
    const int i = threadIdx.x;
    if (i < world) peer_wait_flag(my_flags + PEER_FLAG_DONE + i, epoch, w, 64);
}

__device__ __forceinline__ uint8_t peer_quantise(float c) {          // io/image/ppm.hpp:17-19, product in double
    const float cl = fminf(fmaxf(c, 0.0f), 1.0f);
    return uint8_t(__dmul_rn(255.999, double(cl)));
}

// n4 = number of float4 groups of the whole frame; this rank owns groups [g0, g1).
// U groups per thread and round, U x RB = 16 peer loads of 16 bytes in flight per thread (RB = ranks per batch): the kernel is
// bound by the NVLink round trip, not by bandwidth, so what it needs is loads in flight - with them a SMALL grid moves the frame
// in a few tens of microseconds and leaves the SM slots to the render of the next frame it runs beside.
// Where rank r's samples of this frame are read from, as a pointer that is indexed with the FRAME's element index: the peer's
// framebuffer slot itself (P2P loads through NVLink), or this rank's staging copy of its slice of it - which the copy engines
// fetched, so that no SM waits on an NVLink round trip (rt_peer_reduce_resolve, RT_B200_PEER_GATHER) - shifted back by the
// slice's first element.
struct PeerSources { const float* p[PEER_MAX]; };

// every peer has published "frame `epoch` rendered" (in front of the copy-engine gather, which cannot wait on a flag itself)
__global__ void k_peer_wait_ready(const uint32_t* my_flags, int world, uint32_t epoch, PeerWait w) {
    const int i = threadIdx.x;
    if (i < world) peer_wait_flag(my_flags + PEER_FLAG_READY + i, epoch, w, 32);
}

template <int U>
__global__ void __launch_bounds__(256) k_peer_reduce_resolve(PeerTable t, PeerSources src, int wait_ready, PeerWait wait, int world, int rank, uint64_t g0, uint64_t g1, float div,
                                                             float* __restrict__ root_rgb, uint8_t* __restrict__ root_rgb8,
                                                             uint32_t epoch, uint64_t n_tail_begin, uint64_t n_total, float* __restrict__ own_rgb) {
    constexpr int RB = 16 / U;
    // ---- wait until every peer has rendered frame `epoch` (flags are in MY memory, peers store into them) ----
    if (wait_ready) {
        if (threadIdx.x < world) peer_wait_flag(t.flags[rank] + PEER_FLAG_READY + threadIdx.x, epoch, wait, 32);
        __syncthreads();
    }

    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t base = g0 + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; base < g1; base += stride * U) {
        float4 s[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + u * stride < g1) s[u] = __ldcg(reinterpret_cast<const float4*>(src.p[0]) + base + u * stride);
        // a batch of peers' loads is issued before the first add (one NVLink round trip per batch, not one per rank);
        // the adds stay in rank order = sample order (render.hpp:66-72)
        for (int r0 = 1; r0 < world; r0 += RB) {
            float4 v[U][RB];
#pragma unroll
            for (int k = 0; k < RB; ++k)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (r0 + k < world && base + u * stride < g1)
                        v[u][k] = __ldcg(reinterpret_cast<const float4*>(src.p[r0 + k]) + base + u * stride);
#pragma unroll
            for (int k = 0; k < RB; ++k)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (r0 + k < world && base + u * stride < g1) { s[u].x = s[u].x + v[u][k].x; s[u].y = s[u].y + v[u][k].y; s[u].z = s[u].z + v[u][k].z; s[u].w = s[u].w + v[u][k].w; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t g = base + u * stride;
            if (g >= g1) continue;
            float4 q = s[u];
            q.x = __fdiv_rn(q.x, div); q.y = __fdiv_rn(q.y, div); q.z = __fdiv_rn(q.z, div); q.w = __fdiv_rn(q.w, div);   // :74
            if (root_rgb) reinterpret_cast<float4*>(root_rgb)[g] = q;
            if (own_rgb) reinterpret_cast<float4*>(own_rgb)[g] = q;        // this rank's slice in its OWN memory: copied to the shared host frame next
            if (root_rgb8)
                reinterpret_cast<uchar4*>(root_rgb8)[g] = make_uchar4(peer_quantise(q.x), peer_quantise(q.y), peer_quantise(q.z), peer_quantise(q.w));
        }
    }
    // the last rank also owns the (< 4 element) tail of a frame whose size is not a multiple of four floats
    if (rank == world - 1 && blockIdx.x == 0) {
        for (uint64_t i = n_tail_begin + threadIdx.x; i < n_total; i += blockDim.x) {
            float s = __ldcg(src.p[0] + i);
            for (int r = 1; r < world; ++r) s = s + __ldcg(src.p[r] + i);
            s = __fdiv_rn(s, div);
            if (root_rgb) root_rgb[i] = s;
            if (own_rgb) own_rgb[i] = s;
            if (root_rgb8) root_rgb8[i] = peer_quantise(s);
        }
    }
    // ---- publish "rank done": last block to finish, after its stores are visible system-wide ----
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(t.flags[rank] + 2 * PEER_MAX, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) t.flags[rank][2 * PEER_MAX] = 0u;
        __threadfence_system();
        if (threadIdx.x < world) st_release_sys(t.flags[threadIdx.x] + PEER_FLAG_DONE + rank, epoch);
    }
}

}  // namespace rtb
