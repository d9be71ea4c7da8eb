// Gradient all-reduce of the data-parallel training step as ONE kernel over NVSwitch peer memory ("one step
// = local fwd+bwd then all-reduce of the hash-table and MLP gradients", core/train/trainers/occnerf/trainer.py:246-248 under
// nn.DataParallel upstream).
//
// Every rank holds the same flat fp32 buffer in SYMMETRIC memory (allocated and exchanged by torch's symmetric-memory rendezvous;
// this file only sees raw pointers): the hash-table gradient is scattered into it by occnerf_hashgrid_backward directly, the small
// gradients are copied in.  Two-shot all-reduce: rank r owns slice r of the buffer;
//   * with a multicast (NVLS) mapping: multimem.ld_reduce.add.v4.f32 on the multicast address -- the SWITCH adds the 16 bytes of
//     all ranks and returns the sum -- followed by multimem.st of the result -- the switch writes it into every rank's copy;
//   * without one: the slice is summed from the peers' buffers by plain 16-byte loads over NVLink and stored into every peer.
// Ranks rendezvous INSIDE the kernel through a signal pad in symmetric memory (one slot per (block, peer), monotonically increasing
// epochs kept in device memory, so the kernel is replayable from a CUDA graph with frozen arguments): block b of every rank waits for
// block b of every peer before it reads (the peers' backward kernels are then complete: they precede the all-reduce in stream order)
// and after it has written (so that the kernels that follow see every peer's slice).  Every spin has a limit that traps instead of
// hanging the GPU.
#include "common.cuh"

namespace {

constexpr int kMaxWorld = 8;
constexpr int kThreads = 512;
constexpr uint32_t kSpinLimit = 1u << 25;        // tens of seconds: covers start-up skew between ranks; a rank that never arrives traps

struct Peers {
    float *buf[kMaxWorld];         // every rank's copy of the flat buffer (peer-mapped device pointers; buf[rank] = local)
    uint32_t *pad[kMaxWorld];      // every rank's signal pad: [blocks][world] u32
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// block `blockIdx.x` of this rank <-> block `blockIdx.x` of every peer
__device__ __forceinline__ void peer_barrier(const Peers &P, int rank, int world, uint32_t epoch) {
    __syncthreads();
    if (threadIdx.x < world) {
        const int p = threadIdx.x;
        __threadfence_system();
        st_release_sys(P.pad[p] + blockIdx.x * world + rank, epoch);
        const uint32_t *mine = P.pad[rank] + blockIdx.x * world + p;
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0)
            if (++spins > kSpinLimit) __trap();
    }
    __syncthreads();
}

__device__ __forceinline__ float4 mc_ld_reduce(const float *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float *mc, const float4 &v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Where a launch spends its time (block 0, thread 0; nanoseconds of %globaltimer summed over launches):
// [0] waiting for the peers to arrive  [1] the data phase  [2] waiting for the peers to finish  [3] launches
__device__ unsigned long long g_ar_dbg[4];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// n4: number of float4 of the flat buffer.  epochs: [gridDim.x] u32 in LOCAL device memory, the barrier count of each block so far.
template <bool MULTICAST>
__global__ void __launch_bounds__(kThreads) allreduce_sum_kernel(Peers P, float *mc, long n4, int rank, int world, uint32_t *epochs) {
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) s_epoch = epochs[blockIdx.x];
    __syncthreads();
    const uint32_t e0 = s_epoch;
    const bool dbg = blockIdx.x == 0 && threadIdx.x == 0;
    const unsigned long long t0 = dbg ? gtime() : 0;
    peer_barrier(P, rank, world, e0 + 1);                            // every peer's gradients are complete
    const unsigned long long t1 = dbg ? gtime() : 0;
    const long per = (n4 + world - 1) / world;
    const long begin = min((long)rank * per, n4), end = min(begin + per, n4);
    const long stride = (long)gridDim.x * kThreads;
    if (MULTICAST) {
        float4 *m4 = reinterpret_cast<float4 *>(mc);
        long i = begin + (long)blockIdx.x * kThreads + threadIdx.x;
        constexpr int U = 8;                                         // independent 16-byte round trips in flight per thread
        for (; i + (U - 1) * stride < end; i += U * stride) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = mc_ld_reduce(reinterpret_cast<const float *>(m4 + i + u * stride));
#pragma unroll
            for (int u = 0; u < U; ++u) mc_st(reinterpret_cast<float *>(m4 + i + u * stride), v[u]);
        }
        for (; i < end; i += stride) mc_st(reinterpret_cast<float *>(m4 + i), mc_ld_reduce(reinterpret_cast<const float *>(m4 + i)));
    } else {
        for (long i = begin + (long)blockIdx.x * kThreads + threadIdx.x; i < end; i += stride) {
            float4 v[kMaxWorld];
#pragma unroll
            for (int p = 0; p < kMaxWorld; ++p)                        // all peers' pieces in flight together (system scope: never a
                if (p < world)                                         //  stale line of a peer's buffer)
                    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[p].x), "=f"(v[p].y), "=f"(v[p].z), "=f"(v[p].w) : "l"(reinterpret_cast<const float4 *>(P.buf[p]) + i) : "memory");
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int p = 0; p < kMaxWorld; ++p)                        // same order on every rank: the owner of a slice sums it once
                if (p < world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
#pragma unroll
            for (int p = 0; p < kMaxWorld; ++p)
                if (p < world) reinterpret_cast<float4 *>(P.buf[p])[i] = s;
        }
    }
    __threadfence_system();
    const unsigned long long t2 = dbg ? gtime() : 0;
    peer_barrier(P, rank, world, e0 + 2);                            // every peer has written its slice into my copy
    if (threadIdx.x == 0) epochs[blockIdx.x] = e0 + 2;
    if (dbg) {
        const unsigned long long t3 = gtime();
        g_ar_dbg[0] += t1 - t0; g_ar_dbg[1] += t2 - t1; g_ar_dbg[2] += t3 - t2; g_ar_dbg[3] += 1;
    }
}

}  // namespace

// debug only: the phase times described at g_ar_dbg (4 x u64 to host memory), optionally cleared
extern "C" int occnerf_allreduce_debug(unsigned long long *host4, int reset) {
    OCC_CUDA(cudaDeviceSynchronize());
    if (host4) OCC_CUDA(cudaMemcpyFromSymbol(host4, g_ar_dbg, sizeof(unsigned long long) * 4));
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        OCC_CUDA(cudaMemcpyToSymbol(g_ar_dbg, z, sizeof(z)));
    }
    return OCCNERF_OK;
}

// In-place sum over `world` ranks of the flat fp32 buffer each rank holds in symmetric memory.
// peer_bufs_host / peer_pads_host: host arrays of `world` device pointers (entry `rank` = the local copy); multicast: the multicast
// mapping of the buffer or NULL (peer loads/stores are used instead); n_floats: a multiple of 4; pad: >= blocks * world u32 per rank,
// zeroed once before the first call; epochs: [blocks] u32 local device memory, zeroed once; blocks: the same on every rank.
extern "C" int occnerf_allreduce_sum_f32(const void *const *peer_bufs_host, const void *const *peer_pads_host, void *multicast, long n_floats,
                                         int rank, int world, int blocks, void *epochs, occnerf_stream_t stream) {
    OCC_CHECK_ARG(peer_bufs_host && peer_pads_host && epochs, "allreduce_sum_f32: null pointer");
    OCC_CHECK_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "allreduce_sum_f32: rank=%d world=%d (at most %d ranks)", rank, world, kMaxWorld);
    OCC_CHECK_ARG(n_floats >= 0 && n_floats % 4 == 0, "allreduce_sum_f32: n_floats=%ld must be a multiple of 4", n_floats);
    OCC_CHECK_ARG(blocks >= 1 && blocks <= 148, "allreduce_sum_f32: blocks=%d (1..148: all blocks of all ranks must be resident at once)", blocks);
    if (world == 1 || n_floats == 0) return OCCNERF_OK;
    Peers P = {};
    for (int p = 0; p < world; ++p) {
        OCC_CHECK_ARG(peer_bufs_host[p] && peer_pads_host[p], "allreduce_sum_f32: peer %d has a null pointer", p);
        OCC_CHECK_ARG(((uintptr_t)peer_bufs_host[p] & 15) == 0, "allreduce_sum_f32: buffers must be 16-byte aligned");
        P.buf[p] = (float *)peer_bufs_host[p];
        P.pad[p] = (uint32_t *)peer_pads_host[p];
    }
    if (multicast) allreduce_sum_kernel<true><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(P, (float *)multicast, n_floats / 4, rank, world, (uint32_t *)epochs);
    else allreduce_sum_kernel<false><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(P, nullptr, n_floats / 4, rank, world, (uint32_t *)epochs);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
