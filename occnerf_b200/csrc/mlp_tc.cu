// K3: the canonical density/colour MLP as fused tcgen05/TMEM kernels: forward chain and data-gradient chain.
//
// Replaces the ten nn.Linear (+ReLU) launches of CanonicalMLP.forward and their autograd backward
// (core/nets/occnerf/canonical_mlps/occnerf_mlp.py:183-199):
//     geo  : [agg35,var1,h32] (68) -> 256 -> 256 -> 256 -> 256 -> 65 (sigma + 64 features)
//     rgb  : [geo64,agg35,h32] (131) -> 256 -> 256 -> 256 -> 256 -> 3
// A CTA owns a tile of 128 samples and walks a CHAIN of 10 GEMMs; the activation tile never leaves the SM:
//  * A operand: bf16 in shared memory, K-major, SWIZZLE_NONE, core-matrix-major (each epilogue thread writes one
//    16-byte chunk per 8 columns, conflict free), rewritten in place by the epilogue of the previous GEMM;
//  * accumulators: two 256-column fp32 buffers in TMEM (ping-pong by chain position);
//  * weights: pre-packed once per step (occnerf_mlp_pack_weights) into exactly the shared-memory image the tensor core
//    wants and streamed from L2 through a 3 x 32 KB ring with cp.async.bulk + mbarrier complete_tx (UBLKCP);
//  * fine-grained hand-off: the epilogue publishes the next A operand in 32-column groups (a_ready[g] mbarriers),
//    so the MMAs of GEMM l+1 run while the epilogue of GEMM l is still draining TMEM -- tensor pipe and epilogue
//    overlap inside one tile without a second activation buffer.
//
// Precision: n_pass = 1 is bf16 x bf16 -> fp32.  n_pass = 3 is split-bf16, x = hi + lo (both bf16),
//   x.w ~= hi.whi + hi.wlo + lo.whi (three MMAs per K-step on one fp32 accumulator, ~2^-16 relative error per product).
//   n_pass = 2 is kind::tf32: fp32 operands in shared memory, rounded to nearest-tf32 (10-bit mantissa, fp32 exponent
//   range) by the epilogue / the weight packer; one tf32 MMA covers K = 8 at the cycle cost of a K = 16 bf16 MMA, i.e. two
//   bf16-units per product instead of the three of the split (the A operand is 128 KB either way).
//
// Warp roles (576 threads): warps 0-15 epilogue -- warp w owns TMEM lane quarter w & 3 (32 samples) and, as member of
// epilogue set w >> 2, the 8-column chunks k8 = set (mod 4) of every layer (one chunk of each 32-column group), so each
// SM sub-partition holds FOUR epilogue warps that hide each other's TMEM-load / conversion latencies and every A group
// is completed by all four sets together, in consumption order; warp 16 lane 0 weight producer, warp 17 lane 0 MMA issuer.
// Every mbarrier wait has a spin limit that traps instead of hanging the GPU.
#include <cuda_bf16.h>
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 576;
constexpr int kEpiSets = 4;                 // epilogue warps per TMEM lane quarter
constexpr int kEpiThreads = kEpiSets * 128;
constexpr int kStages = 3;                  // (6 x 16 KB with half-size chunks was measured 15 % slower: more, smaller bulk copies)
constexpr int kStageBytes = 32768;
constexpr int kLayers = 10;
constexpr int kAPartBytes = 65536;          // 128 rows x 256 K x bf16
constexpr int kGroupCols = 64;              // hand-off granularity of the A operand: columns per a_ready group.  32 = one publish
                                            // (2 fences + arrive, ~250 cycles of latency per warp) per 8-column chunk of a thread;
                                            // 64 = one per two chunks: the epilogue then outruns the MMAs (trace: tools/mlp_trace.py)
constexpr int kGroups = 256 / kGroupCols;
constexpr int kGroupK8 = kGroupCols / 8;    // 8-column chunks per group
constexpr uint32_t kSpinLimit = 1u << 27;

// (Tried and reverted: evaluating the 256 -> 3 output layer -- 16 K-steps of at least 67 cycles each for 3 useful columns --
//  with fp32 FMAs inside the rgb3 epilogue, and its transpose at the head of the backward chain.  Inference forward -2 %,
//  but forward-with-saving +5 % and dgrad +7 %: the extra W_out loads and registers sit on the epilogue's critical path.)
// padded GEMM shapes (K = contraction, N = output width; both multiples of 16)
// forward chain : pts0, pts1-3, geo, rgb0, rgb1-3, out
// backward chain: out^T, rgb3..1^T, rgb0^T, geo^T, pts3..1^T, pts0^T   (position d uses forward layer 9-d)
__host__ __device__ inline int fwd_K(int l) { return l == 0 ? 80 : (l == 5 ? 144 : 256); }
__host__ __device__ inline int fwd_N(int l) { return l == 4 ? 80 : (l == 9 ? 16 : 256); }
// chain 2 = the non-rigid motion MLP (non_rigid_motion_mlps/mlp_offset.py:7-62), forward only, 7 GEMMs:
//   pe36 (+12 pad) -> 128 -> 128 -> 128 -> 128 -> [h128, pe36] (176) -> 128 -> 128 -> 3 (16)
// (the 69-value pose condition is the same for every sample of a frame and is folded into the bias of layer 0)
__host__ __device__ inline int nr_K(int l) { return l == 0 ? 48 : (l == 4 ? 176 : 128); }
__host__ __device__ inline int nr_N(int l) { return l == 6 ? 16 : 128; }
__host__ __device__ inline int n_layers(int chain) { return chain == 2 ? 7 : 10; }
__host__ __device__ inline int chain_K(int chain, int l) { return chain == 0 ? fwd_K(l) : (chain == 1 ? fwd_N(9 - l) : nr_K(l)); }
__host__ __device__ inline int chain_N(int chain, int l) { return chain == 0 ? fwd_N(l) : (chain == 1 ? fwd_K(9 - l) : nr_N(l)); }
// K columns per weight chunk (= one ring slot).  Pair mode streams half of the weight rows per CTA, so the same slot bytes hold twice
// the K: 64 columns for every engine (the MMA thread pays its per-chunk cost -- barrier polls, loop, commit -- half as often)
__host__ __device__ inline int chunk_K(int n_pass, int pair = 0) { return (n_pass == 1 || pair) ? 64 : 32; }
__host__ __device__ inline int parts(int n_pass) { return n_pass == 3 ? 2 : 1; }
__host__ __device__ inline int elem_bytes(int n_pass) { return n_pass == 2 ? 4 : 2; }
__host__ __device__ inline bool valid_pass(int n_pass) { return n_pass >= 1 && n_pass <= 3; }

struct PackedLayout {
    long w_off[kLayers];
    long bias_off;
    long total;
};

inline PackedLayout packed_layout(int n_pass, int chain, int pair) {
    PackedLayout p;
    long off = 0;
    const int KC = chunk_K(n_pass, pair);
    for (int l = 0; l < kLayers; ++l) p.w_off[l] = 0;
    for (int l = 0; l < n_layers(chain); ++l) {
        p.w_off[l] = off;
        const int K = chain_K(chain, l), N = chain_N(chain, l);
        off += (long)((K + KC - 1) / KC) * parts(n_pass) * N * KC * elem_bytes(n_pass);
    }
    p.bias_off = off;
    off += kLayers * 256 * 4;
    p.total = (off + 255) / 256 * 256;
    return p;
}

// ------------------------------------------------------------------------------------------ weight packing
// element [n][k] of the padded FORWARD GEMM of layer l, with the row/column re-ordering described in mlp.py
__device__ __forceinline__ float fwd_weight(const occnerf_mlp_params &P, int l, int n, int k) {
    if (l == 0) return (k < 68) ? P.w[0][n * 68 + k] : 0.f;                                   // pts0: [agg35,var,h32]
    if (l == 4) {                                                                              // geo: features first, sigma at 64
        if (n < 64) return P.w[4][(n + 1) * 256 + k];
        if (n == 64) return P.w[4][k];
        return 0.f;
    }
    if (l == 5) {                                                                              // rgb0: [geo64 | agg35, var(0), h32]
        if (k < 99) return P.w[5][n * 131 + k];
        if (k == 99 || k >= 132) return 0.f;
        return P.w[5][n * 131 + k - 1];
    }
    if (l == 9) return (n < 3) ? P.w[9][n * 256 + k] : 0.f;
    return P.w[l][n * 256 + k];
}
__device__ __forceinline__ float fwd_bias(const occnerf_mlp_params &P, int l, int n) {
    if (l == 4) return n < 64 ? P.b[4][n + 1] : (n == 64 ? P.b[4][0] : 0.f);
    if (l == 9) return n < 3 ? P.b[9][n] : 0.f;
    return P.b[l][n];
}

struct DevLayout { long w_off[kLayers]; long bias_off; };

__device__ __forceinline__ uint32_t to_tf32(float x) {       // round to nearest (ties away), low 13 mantissa bits cleared
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}

// element (n, k) of a padded N x K weight matrix -> its place in the packed operand image of layer `base`:
// per K-chunk of KC columns [part][k-slab][n/8][n%8][16 B] (K-major, no swizzle: 8 rows x 16 B core matrices, a k-slab =
// 8 bf16 / 4 tf32 columns of all N rows)
// With cta_pair the chunk image is split by weight-row HALVES, [half][part][k-slab][...]: CTA r of a pair streams half r.
__device__ __forceinline__ void pack_store(unsigned char *out, long base, int n_pass, int N, int n, int k, float w, int cta_pair = 0) {
    const int KC = chunk_K(n_pass, cta_pair), np = parts(n_pass);
    const int chunk = k / KC, kk = k % KC;
    long half_off = 0;
    if (cta_pair) {
        const int Nh = N / 2, h = n / Nh;
        half_off = (long)h * np * Nh * KC * elem_bytes(n_pass);
        n -= h * Nh;
        N = Nh;
    }
    if (n_pass == 2) {
        const long cb = base + (long)chunk * (cta_pair ? 2 : 1) * N * KC * 4 + half_off;
        const long inner = ((long)(kk / 4) * (N / 8) + n / 8) * 128 + (n % 8) * 16 + (kk % 4) * 4;
        *reinterpret_cast<uint32_t *>(out + cb + inner) = to_tf32(w);
        return;
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    const long pb = (long)N * KC * 2;
    const long cb = base + (long)chunk * (cta_pair ? 2 : 1) * np * pb + half_off;
    const long inner = ((long)(kk / 8) * (N / 8) + n / 8) * 128 + (n % 8) * 16 + (kk % 8) * 2;
    *reinterpret_cast<__nv_bfloat16 *>(out + cb + inner) = hi;
    if (np == 2) *reinterpret_cast<__nv_bfloat16 *>(out + cb + pb + inner) = lo;
}

__global__ void pack_weights_kernel(occnerf_mlp_params P, DevLayout L, int n_pass, int chain, int cta_pair, unsigned char *out) {
    const int l = blockIdx.y;
    const int K = chain_K(chain, l), N = chain_N(chain, l);
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (long)N * K) {
        const int n = (int)(idx / K), k = (int)(idx % K);
        const float w = chain == 0 ? fwd_weight(P, l, n, k) : fwd_weight(P, 9 - l, k, n);     // backward chain: W^T
        pack_store(out, L.w_off[l], n_pass, N, n, k, w, cta_pair);
    }
    if (chain == 0 && blockIdx.x == 0 && threadIdx.x < 256) {
        float *b = reinterpret_cast<float *>(out + L.bias_off) + l * 256;
        b[threadIdx.x] = threadIdx.x < N ? fwd_bias(P, l, threadIdx.x) : 0.f;
    }
}

struct NrParams { const float *w[7]; const float *b[7]; const float *cond; };   // cond: 69 floats on the device or NULL

__global__ void pack_nr_kernel(NrParams P, DevLayout L, int n_pass, int cta_pair, unsigned char *out) {
    const int l = blockIdx.y;
    const int K = nr_K(l), N = nr_N(l);
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (long)N * K) {
        const int n = (int)(idx / K), k = (int)(idx % K);
        float w = 0.f;
        if (l == 0) w = k < 36 ? P.w[0][n * 105 + 69 + k] : 0.f;                 // the PE columns of [cond69, pe36]
        else if (l == 4) w = k < 164 ? P.w[4][n * 164 + k] : 0.f;                // [h128, pe36]
        else if (l == 6) w = n < 3 ? P.w[6][n * 128 + k] : 0.f;
        else w = P.w[l][n * 128 + k];
        pack_store(out, L.w_off[l], n_pass, N, n, k, w, cta_pair);
    }
    if (blockIdx.x == 0 && threadIdx.x < 256) {
        const int n = threadIdx.x;
        float b = 0.f;
        if (n < N && (l != 6 || n < 3)) {
            b = P.b[l][n];
            if (l == 0 && P.cond)                                                 // fold the per-frame condition code
                for (int k = 0; k < 69; ++k) b = fmaf(P.w[0][n * 105 + k], P.cond[k], b);
        }
        reinterpret_cast<float *>(out + L.bias_off)[l * 256 + n] = b;
    }
}

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// the same copy, delivered to the same shared-memory offset (and mbarrier) of every CTA in `mask` of the cluster
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrives on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::f16 (bf16 operands, fp32 accumulate), issued by one thread
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with kind::tf32 (fp32 containers in shared memory, K = 8 per instruction)
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- CTA-pair mode (cta_group::2): one UMMA of M = 256 spans both CTAs of a cluster -- each CTA holds its own 128 rows of A and
// HALF of the B operand (N/2 weight rows), the leader CTA's elected thread issues for both, the accumulator rows land in each
// CTA's own TMEM.  Per CTA that halves the weight bytes streamed into shared memory and the B bytes the tensor core reads.
__device__ __forceinline__ void tc_mma_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all prior MMAs of this thread -> one arrival on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_cg2_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// (default .release.cta semantics as in CUTLASS' ClusterBarrier::arrive: the data the arrival publishes is consumed by the
//  async proxy -- fence.proxy.async orders it --, not by the waiting thread; an explicit .release.cluster was measured to
//  double the epilogue's time per publish)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a LOCAL mbarrier whose arrivals come from the peer CTA as well (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > kSpinLimit) __trap();
    }
}

// K-major, no-swizzle shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout, version 1 = Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128
__device__ __forceinline__ uint32_t instr_desc(int n, int m = kTileM) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::tf32 instruction descriptor: D=f32, A=B=tf32 (format 2), both K-major, M=128
__device__ __forceinline__ uint32_t instr_desc_tf32(int n, int m = kTileM) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// (a0, a1) += (b0, b1) as one packed fp32x2 add (FADD2 on sm_100): halves the bias adds of the epilogue
__device__ __forceinline__ void add2(float &a0, float &a1, float b0, float b1) {
    asm("{ .reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; add.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x; }"
        : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}

__device__ __forceinline__ uint4 pack_bf16x8(const float (&v)[8]) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        pk[i] = *reinterpret_cast<const uint32_t *>(&h);
    }
    return make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// bf16 split of 8 consecutive K values of one row -> one 16-byte store per part into the A operand image
// (returns the packed bf16 "hi" part: exactly what the backward pass wants saved)
template <int NPASS>
__device__ __forceinline__ uint4 store_a8(unsigned char *a_base, int row, int k8, const float (&v)[8]) {
    uint32_t hi[4], lo[4];
    if (NPASS == 2) {
        // tf32 engine: the operand is fp32-sized, 4 columns per 16-byte core-matrix row -> k-slabs 2*k8 and 2*k8+1
        const uint32_t off = (uint32_t)(k8 * 32 + (row >> 3)) * 128 + (row & 7) * 16;
        // round to nearest tf32, ties away (= cvt.rna.tf32.f32 for finite values, which ptxas expands to 5 instructions with
        // the inf/nan handling): add half an ulp of the 10-bit mantissa to the magnitude.  The 13 low mantissa bits stay as they
        // are: kind::tf32 ignores them (measured on B200: outputs bitwise identical with and without clearing them,
        // tools/mlp_exp.py `nomask`, gpurun_out/r2o_mlp_exp.json), which saves one LOP3 per value in the epilogue
        auto rna = [](float x) { return __float_as_uint(x) + 0x1000u; };
        *reinterpret_cast<uint4 *>(a_base + off) = make_uint4(rna(v[0]), rna(v[1]), rna(v[2]), rna(v[3]));
        *reinterpret_cast<uint4 *>(a_base + off + 2048) = make_uint4(rna(v[4]), rna(v[5]), rna(v[6]), rna(v[7]));
        return pack_bf16x8(v);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        hi[i] = *reinterpret_cast<const uint32_t *>(&h);
        if (NPASS == 3) {
            const float2 hf = __bfloat1622float2(h);
            const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
            lo[i] = *reinterpret_cast<const uint32_t *>(&l);
        }
    }
    const uint32_t off = (uint32_t)(k8 * 16 + (row >> 3)) * 128 + (row & 7) * 16;
    *reinterpret_cast<uint4 *>(a_base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (NPASS == 3) *reinterpret_cast<uint4 *>(a_base + kAPartBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    return make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

struct ChainArgs {
    int m;
    long slot_stride;                // rows per slot of act_save / act / g_save (>= m; a multiple of 64 for the wgrad kernel)
    int chain;                       // 0 forward, 1 backward
    const unsigned char *packed;
    long w_off[kLayers];
    long bias_off;
    // forward
    float *XB;                       // [m,132]: cols 64..131 in; cols 0..63 out (geo features) when act_dtype != 0
    float *raw;                      // [m, ldr]: cols 0..2 rgb_pre, col 3 sigma_pre
    int ldr;
    void *act_save;                  // post-ReLU activations or NULL: fp32 [8][slot_stride][256], or bf16 chunk-major
                                     // [10][32][slot_stride][8] (see saved_off) with slot 8 = pts0 input (80 cols) and
                                     // slot 9 = rgb0 input (144 cols)
    int act_dtype;                   // 0 none, 1 fp32, 2 bf16
    uint8_t *relu_mask;              // [8][32][slot_stride] bytes: bit i of byte (slot, k8, row) = [unit 8*k8+i > 0]; written by
                                     // the forward chain in bf16 mode and read by the backward chain (32 B per row and layer)
    int debug;
    // non-rigid chain (chain 2)
    const float *nr_xyz;             // [m,3]
    float nr_window[6];              // Hann window weights of the 6 frequency bands (ops.hann_window)
    float *nr_out;                   // [m,3] = xyz + MLP(pe; cond)
    // backward
    const float *g_raw;              // [m,5]
    float *gXB;                      // [m,132]: cols 64..131 written (d agg, d var, d h; both trunks summed)
    __nv_bfloat16 *g_save;           // bf16 chunk-major [10][32][slot_stride][8]: gradients w.r.t. the pre-activations
                                     // (slot 9 = d raw[:, :3])
};

// Debug instrumentation (tools/mlp_stalls.py): with OCCNERF_MLP_DEBUG=1 in the environment the role threads accumulate the
// cycles they spend blocked on each mbarrier; summed over all CTAs into g_dbg.  [0] MMA waits for weights, [1] MMA waits for
// the A operand, [2] epilogue (row 0, set 0) waits for the accumulator, [3] producer waits for a free ring slot,
// [4] MMA thread total, [5] epilogue thread total, [6] CTAs.
__device__ unsigned long long g_dbg[16];
__device__ __forceinline__ long long clk() { return clock64(); }
// Per-layer time stamps (OCCNERF_MLP_DEBUG bit 4) of CTA 0's third tile, clock64 of one SM: [layer][event]
//  0 MMA thread enters the layer   1 first MMA issued   2 last MMA issued   3 epilogue warp 0: accumulator ready
//  4 warp 0: first group published   5 warp 0: last group published   6 warp 15: accumulator ready   7 warp 15: last group published
//  8 producer: last chunk of the layer issued   10 MMA thread: cycles waiting for A in the layer   11 ... for weights
__device__ unsigned long long g_trace[16][12];
// Weight-stream round trip (same tile, layer 2 and 3): [layer - 2][chunk][event]  0 producer: slot seen empty  1 producer: copy issued
//  2 MMA thread: starts waiting for the chunk  3 own half landed  4 peer's half reported  5 MMAs + commit issued
__device__ unsigned long long g_trace_w[2][8][6];
#define TRACEW(cond, l, c, e) do { if ((cond) && ((l) == 2 || (l) == 3)) g_trace_w[(l) - 2][c][e] = (unsigned long long)clock64(); } while (0)
#define TRACE(cond, l, e) do { if (cond) g_trace[l][e] = (unsigned long long)clock64(); } while (0)

// bf16 activations / gradients saved for the weight-gradient kernel use a CHUNK-MAJOR layout [slot][k8 = col/8][row][8]:
// the 32 lanes of an epilogue warp (32 consecutive rows, one 8-column chunk) then write 512 contiguous bytes instead of
// 32 separate 16-byte pieces 512 B apart (which cost one L1 tag cycle each: measured 8k cycles per layer and tile), and
// the weight-gradient kernel's TMA boxes land in shared memory directly in the no-swizzle MN-major UMMA layout.
__device__ __forceinline__ long saved_off(int slot, int k8, long stride, long row) { return (((long)slot * 32 + k8) * stride + row) * 8; }

struct Smem {
    unsigned char *A, *W;
    float *bias;                     // 2 x 256 floats: the bias of the layer in flight and of the next one (double buffer)
    uint32_t bar_w_full, bar_w_empty, bar_a_ready, bar_acc_full;
    uint32_t bar_w_peer;             // pair mode, leader: "the peer's half of the chunk has landed" (arrived by the peer's relay)
    uint32_t a_ready_arrive;         // where the epilogue warps arrive: the local a_ready, or the leader's (cluster address)
    bool pair;
};

// ---- weight producer: streams every K-chunk of every GEMM of every tile through the ring, running ahead freely
template <int NPASS>
__device__ __forceinline__ void producer_loop(const ChainArgs &args, const Smem &sm, int num_tiles) {
    constexpr int KC = (NPASS == 1) ? 64 : 32;
    constexpr int NP = (NPASS == 3) ? 2 : 1;
    constexpr int EB = (NPASS == 2) ? 4 : 2;
    uint32_t it = 0;
    long long dbg_wait = 0;
    const uint32_t rank = cluster_ctarank();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const bool tr = (args.debug & 16) && blockIdx.x == 0 && tile == 2 * (int)gridDim.x;
        for (int l = 0; l < n_layers(args.chain); ++l) {
            const int K = chain_K(args.chain, l), N = chain_N(args.chain, l);
            const int nch = (K + KC - 1) / KC;
            const uint32_t pb = (uint32_t)N * KC * EB;
            const unsigned char *src = args.packed + args.w_off[l];
            for (int c = 0; c < nch; ++c, ++it) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                const long long t0 = args.debug ? clk() : 0;
                mbar_wait(sm.bar_w_empty + 8 * s, ph ^ 1);          // BOTH CTAs of the pair have consumed the slot
                if (args.debug) dbg_wait += clk() - t0;
                const int kc = min(KC, K - c * KC);
                const uint32_t bytes = (uint32_t)N * kc * EB;      // per part; the k-slab-outer image is contiguous
                const uint32_t all = bytes * NP;                   // what lands in MY slot (from both CTAs of the pair)
                const uint32_t dst = smem_u32(sm.W + s * kStageBytes);
                const uint32_t bar = sm.bar_w_full + 8 * s;
                if (args.debug & 14) {
                    // TIMING EXPERIMENTS (OCCNERF_MLP_DEBUG bits 1..3; results are garbage except for bit 3): they separate the
                    // L2-output bound from the peer-delivery bound of the weight stream.
                    //   bit 1: both CTAs fetch a quarter and multicast it -> half of the bytes everywhere
                    //   bit 2: every CTA fetches its own half only, nothing is multicast (normal L2 output, no peer traffic)
                    //   bit 3: every CTA fetches the whole chunk itself, nothing is multicast (2x L2 output; numerically valid)
                    const unsigned char *chunk = src + (long)c * NP * pb;
                    if (args.debug & 8) {
                        mbar_arrive_expect_tx(bar, all);
                        bulk_g2s(dst, chunk, bytes, bar);
                        if (NP == 2) bulk_g2s(dst + kStageBytes / 2, chunk + pb, bytes, bar);
                    } else if (args.debug & 4) {
                        mbar_arrive_expect_tx(bar, all / 2);
                        bulk_g2s(dst + rank * (all / 2), chunk + rank * (all / 2), all / 2, bar);
                    } else {
                        mbar_arrive_expect_tx(bar, all / 2);
                        bulk_g2s_mc(dst + rank * (all / 4), chunk + rank * (all / 4), all / 4, bar, 3);
                    }
                    continue;
                }
                mbar_arrive_expect_tx(bar, all);
                // the two CTAs of a cluster walk the same weight stream: each fetches one half from L2 and multicasts
                // it into both shared memories, halving the L2 -> SM weight traffic
                if (NP == 2) {
                    bulk_g2s_mc(dst + rank * (kStageBytes / 2), src + ((long)c * 2 + rank) * pb, bytes, bar, 3);
                } else {
                    const uint32_t half = bytes / 2;
                    bulk_g2s_mc(dst + rank * half, src + (long)c * pb + rank * half, half, bar, 3);
                }
            }
            TRACE(tr, l, 8);
        }
    }
    if (args.debug) atomicAdd(&g_dbg[3], (unsigned long long)dbg_wait);
}

// ---- MMA issuer: GEMM l accumulates into TMEM buffer (l & 1); it consumes the A operand group by group as the
//      epilogue of GEMM l-1 publishes it
// one lane of a converged warp (the warp-uniform way to issue single-thread instructions: everything around it stays in the
// uniform datapath, no per-operand R2UR / ELECT loops as from a divergent `if (lane == 0)` region)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}

template <int NPASS>
__device__ __forceinline__ void mma_step(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    if (NPASS == 2) {
        tc_mma_tf32(d_tmem, da, db, idesc, accumulate);
    } else {
        tc_mma(d_tmem, da, db, idesc, accumulate);
        if (NPASS == 3) {
            tc_mma(d_tmem, da, db + (uint64_t)((kStageBytes / 2) >> 4), idesc, 1u);       // hi . Wlo
            tc_mma(d_tmem, da + (uint64_t)(kAPartBytes >> 4), db, idesc, 1u);             // lo . Whi
        }
    }
}

// ---- MMA issuer (ALL 32 lanes of the MMA warp run this loop, one elected lane issues): GEMM l accumulates into TMEM buffer
//      (l & 1); it consumes the A operand group by group as the epilogue of GEMM l-1 publishes it
template <int NPASS>
__device__ __forceinline__ void mma_loop(const ChainArgs &args, const Smem &sm, int num_tiles, uint32_t tmem_base) {
    constexpr int KC = (NPASS == 1) ? 64 : 32;
    constexpr int KS = (NPASS == 2) ? 8 : 16;          // K per MMA instruction (32 bytes of operand row)
    constexpr int CPG = kGroupCols / KC;               // weight chunks per A group (kGroupCols >= KC)
    static_assert(kGroupCols % KC == 0, "an A group is a whole number of weight chunks");
    uint32_t it = 0, a_phase = 0;      // a_phase: one parity bit per A group
    long long dbg_w = 0, dbg_a = 0;
    const long long dbg_t0 = args.debug ? clk() : 0;
    const uint32_t a_base = smem_u32(sm.A);
    const bool lane0 = (threadIdx.x & 31) == 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const bool tr = (args.debug & 16) && blockIdx.x == 0 && tile == 2 * (int)gridDim.x && lane0;
        for (int l = 0; l < n_layers(args.chain); ++l) {
            TRACE(tr, l, 0);
            const long long w0 = dbg_w, a0 = dbg_a;
            const int K = chain_K(args.chain, l), N = chain_N(args.chain, l);
            const int nch = (K + KC - 1) / KC;
            const uint32_t idesc = NPASS == 2 ? instr_desc_tf32(N) : instr_desc(N);
            const uint32_t b_lbo = (uint32_t)(N / 8) * 128;
            const uint32_t d_tmem = tmem_base + (uint32_t)(l & 1) * 256;
            // operand descriptors advance by a constant per K-step (the start-address field holds addr >> 4): A by two k-slabs
            // (4096 B), B by two k-slabs of the chunk image (2 * b_lbo)
            uint64_t da = smem_desc(a_base, 2048, 128);
            const uint64_t db_step = (uint64_t)((2 * b_lbo) >> 4);
            for (int c = 0; c < nch; ++c, ++it) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                {
                    const long long t0 = args.debug ? clk() : 0;
                    mbar_wait(sm.bar_w_full + 8 * s, ph);
                    if (args.debug) dbg_w += clk() - t0;
                }
                if (c % CPG == 0) {                                // first chunk of an A group: wait for the epilogue
                    const int g = c / CPG;
                    const long long t0 = args.debug ? clk() : 0;
                    mbar_wait(sm.bar_a_ready + 8 * g, (a_phase >> g) & 1);
                    if (args.debug) dbg_a += clk() - t0;
                    a_phase ^= 1u << g;
                }
                tc_fence_after();
                TRACE(tr && c == 0, l, 1);
                const int kc = min(KC, K - c * KC);
                uint64_t db = smem_desc(smem_u32(sm.W + s * kStageBytes), b_lbo, 128);
                if (elect_one()) {
                    if (kc == KC) {
#pragma unroll
                        for (int ks = 0; ks < KC / KS; ++ks)
                            mma_step<NPASS>(d_tmem, da + (uint64_t)(ks * (4096 >> 4)), db + ks * db_step, idesc, (c | ks) ? 1u : 0u);
                    } else {
                        for (int ks = 0; ks < kc / KS; ++ks)
                            mma_step<NPASS>(d_tmem, da + (uint64_t)(ks * (4096 >> 4)), db + ks * db_step, idesc, (c | ks) ? 1u : 0u);
                    }
                    tc_commit_mc(sm.bar_w_empty + 8 * s, 3);  // frees the ring slot in both CTAs of the pair once these MMAs have read it
                }
                __syncwarp();
                da += (uint64_t)((KC / KS) * (4096 >> 4));
            }
            TRACE(tr, l, 2);
            if (tr) { g_trace[l][10] = (unsigned long long)(dbg_a - a0); g_trace[l][11] = (unsigned long long)(dbg_w - w0); }
            if (elect_one()) tc_commit(sm.bar_acc_full);   // accumulator of GEMM l complete
            __syncwarp();
        }
    }
    if (args.debug && lane0) {
        atomicAdd(&g_dbg[0], (unsigned long long)dbg_w);
        atomicAdd(&g_dbg[1], (unsigned long long)dbg_a);
        atomicAdd(&g_dbg[4], (unsigned long long)(clk() - dbg_t0));
        atomicAdd(&g_dbg[6], 1ull);
    }
}

// ================================================================== CTA-pair mode (cta_group::2) role loops
// Packed image per chunk = [half][part][k-slab][n/8][8][16 B] (occnerf_mlp_pack_weights with cta_pair = 1), so a CTA's half of a chunk is
// one contiguous piece.
// Ring geometry per engine: bf16 6 x 16 KB, tf32 / split-bf16 3 x 32 KB (a K-chunk of 64 columns of HALF the weight rows, hi [+ lo]).
// Round-2 measurement behind the 64-column chunks (tools/mlp_trace_w.py, profiles/r02_mlp_trace_w.txt): with 32-column chunks the
// weights sat in the ring ~4 k cycles before they were used -- the "weight wait" of the MMA thread was not waiting at all but the
// latency of its own two barrier polls (~120 cycles each, even on a completed phase), and together with ~75 cycles of issue per UMMA
// one chunk of four cost the thread 620-700 cycles against 412 cycles of tensor time.  The issuing thread was the bottleneck of the
// MMA phase, neither L2 bandwidth (fetching half of the bytes changed nothing) nor a same-address hot spot (replicated weight images
// changed nothing).
constexpr int kPairKC = 64;
template <int NPASS> struct PairRing {
    static constexpr int kStagesN = NPASS == 1 ? 6 : 3;
    static constexpr int kBytes = NPASS == 1 ? 16384 : 32768;
};

// both CTAs: stream MY half of every chunk (plain bulk copies, no multicast)
template <int NPASS>
__device__ __forceinline__ void producer_loop_pair(const ChainArgs &args, const Smem &sm, int num_tiles) {
    constexpr int KC = kPairKC;
    constexpr int NP = (NPASS == 3) ? 2 : 1;
    constexpr int EB = (NPASS == 2) ? 4 : 2;
    constexpr int NS = PairRing<NPASS>::kStagesN, SB = PairRing<NPASS>::kBytes;
    uint32_t it = 0;
    const uint32_t rank = cluster_ctarank();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const bool tr = (args.debug & 16) && blockIdx.x == 0 && tile == 2 * (int)gridDim.x;
        for (int l = 0; l < n_layers(args.chain); ++l) {
            const int K = chain_K(args.chain, l), Nh = chain_N(args.chain, l) / 2;
            const int nch = (K + KC - 1) / KC;
            const uint32_t pbh = (uint32_t)Nh * KC * EB;                         // one part of one half of a full chunk
            const unsigned char *src = args.packed + args.w_off[l] + (long)rank * NP * pbh;
            for (int c = 0; c < nch; ++c, ++it) {
                const uint32_t s = it % NS, ph = (it / NS) & 1;
                mbar_wait(sm.bar_w_empty + 8 * s, ph ^ 1);
                TRACEW(tr, l, c, 0);
                const int kc = min(KC, K - c * KC);
                const uint32_t bytes = (uint32_t)Nh * kc * EB;
                const uint32_t dst = smem_u32(sm.W + s * SB);
                const uint32_t bar = sm.bar_w_full + 8 * s;
                mbar_arrive_expect_tx(bar, bytes * NP);
                const unsigned char *chunk = src + (long)c * 2 * NP * pbh;
                bulk_g2s(dst, chunk, bytes, bar);
                if (NP == 2) bulk_g2s(dst + SB / 2, chunk + pbh, bytes, bar);
                TRACEW(tr, l, c, 1);
            }
        }
    }
}

// peer CTA (rank 1): tell the leader's MMA thread that MY half of chunk `it` has landed in MY shared memory
template <int NPASS>
__device__ __forceinline__ void relay_loop_pair(const ChainArgs &args, const Smem &sm, int num_tiles) {
    constexpr int KC = kPairKC;
    constexpr int NS = PairRing<NPASS>::kStagesN;
    uint32_t it = 0;
    const uint32_t leader_w_peer = map_to_cta(sm.bar_w_peer, 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int l = 0; l < n_layers(args.chain); ++l) {
            const int nch = (chain_K(args.chain, l) + KC - 1) / KC;
            for (int c = 0; c < nch; ++c, ++it) {
                const uint32_t s = it % NS, ph = (it / NS) & 1;
                mbar_wait(sm.bar_w_full + 8 * s, ph);
                mbar_arrive_cluster(leader_w_peer + 8 * s);
            }
        }
    }
}

template <int NPASS>
__device__ __forceinline__ void mma_step_pair(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    if (NPASS == 2) {
        tc_mma_tf32_cg2(d_tmem, da, db, idesc, accumulate);
    } else {
        tc_mma_cg2(d_tmem, da, db, idesc, accumulate);
        if (NPASS == 3) {
            tc_mma_cg2(d_tmem, da, db + (uint64_t)((PairRing<3>::kBytes / 2) >> 4), idesc, 1u);   // hi . Wlo
            tc_mma_cg2(d_tmem, da + (uint64_t)(kAPartBytes >> 4), db, idesc, 1u);                 // lo . Whi
        }
    }
}

// non-blocking poll of one phase of a local mbarrier: the result arrives ~100 cycles later, so it is issued one chunk AHEAD of its use
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

// leader CTA (rank 0), whole MMA warp: M = 256 UMMAs over both CTAs' A tiles and B halves.
// Software-pipelined barrier polls: the three barriers chunk i+1 depends on (my half, the peer's half, its A group) are polled BEFORE
// the MMAs of chunk i are issued and the answers are looked at after them; only a negative answer falls back to the blocking wait.
template <int NPASS>
__device__ __forceinline__ void mma_loop_pair(const ChainArgs &args, const Smem &sm, int num_tiles, uint32_t tmem_base) {
    constexpr int KC = kPairKC;
    constexpr int KS = (NPASS == 2) ? 8 : 16;
    constexpr int NS = PairRing<NPASS>::kStagesN, SB = PairRing<NPASS>::kBytes;
    static_assert(kGroupCols == KC, "pair mode: one A group per weight chunk");
    uint32_t it = 0, a_phase = 0;
    long long dbg_w = 0, dbg_a = 0;
    const long long dbg_t0 = args.debug ? clk() : 0;
    const uint32_t a_base = smem_u32(sm.A);
    const bool lane0 = (threadIdx.x & 31) == 0;
    const int nl = n_layers(args.chain);
    uint32_t ok_w = 0, ok_a = 0;                       // early answers for the chunk about to be issued
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const bool tr = (args.debug & 16) && blockIdx.x == 0 && tile == 2 * (int)gridDim.x && lane0;
        for (int l = 0; l < nl; ++l) {
            TRACE(tr, l, 0);
            const long long w0 = dbg_w, a0 = dbg_a;
            const int K = chain_K(args.chain, l), N = chain_N(args.chain, l);
            const int nch = (K + KC - 1) / KC;
            const uint32_t idesc = NPASS == 2 ? instr_desc_tf32(N, 256) : instr_desc(N, 256);
            const uint32_t b_lbo = (uint32_t)(N / 16) * 128;                   // k-slab stride of MY half (N/2 rows)
            const uint32_t d_tmem = tmem_base + (uint32_t)(l & 1) * 256;
            uint64_t da = smem_desc(a_base, 2048, 128);
            const uint64_t db_step = (uint64_t)((2 * b_lbo) >> 4);
            for (int c = 0; c < nch; ++c, ++it) {
                const uint32_t s = it % NS, ph = (it / NS) & 1;
                TRACEW(tr, l, c, 2);
                if (!ok_w) {
                    const long long t0 = args.debug ? clk() : 0;
                    mbar_wait(sm.bar_w_full + 8 * s, ph);                       // my half
                    TRACEW(tr, l, c, 3);
                    mbar_wait_cluster(sm.bar_w_peer + 8 * s, ph);               // the peer's half
                    if (args.debug) dbg_w += clk() - t0;
                }
                TRACEW(tr, l, c, 4);
                if (!ok_a) {
                    const long long t0 = args.debug ? clk() : 0;
                    mbar_wait_cluster(sm.bar_a_ready + 8 * c, (a_phase >> c) & 1);   // A group c: 16 warps of each CTA
                    if (args.debug) dbg_a += clk() - t0;
                }
                a_phase ^= 1u << c;
                tc_fence_after();
                TRACE(tr && c == 0, l, 1);
                // polls for the NEXT chunk (the next layer's first one after the last of this layer)
                {
                    const uint32_t it1 = it + 1, s1 = it1 % NS, ph1 = (it1 / NS) & 1;
                    const int c1 = c + 1 < nch ? c + 1 : 0;
                    ok_w = mbar_test(sm.bar_w_full + 8 * s1, ph1) & mbar_test(sm.bar_w_peer + 8 * s1, ph1);
                    ok_a = mbar_test(sm.bar_a_ready + 8 * c1, (a_phase >> c1) & 1);
                }
                const int nks = min(KC, K - c * KC) / KS;
                uint64_t db = smem_desc(smem_u32(sm.W + s * SB), b_lbo, 128);
                if (elect_one()) {
                    if (nks == KC / KS) {
#pragma unroll
                        for (int ks = 0; ks < KC / KS; ++ks)
                            mma_step_pair<NPASS>(d_tmem, da + (uint64_t)(ks * (4096 >> 4)), db + ks * db_step, idesc, (c | ks) ? 1u : 0u);
                    } else {
                        for (int ks = 0; ks < nks; ++ks)
                            mma_step_pair<NPASS>(d_tmem, da + (uint64_t)(ks * (4096 >> 4)), db + ks * db_step, idesc, (c | ks) ? 1u : 0u);
                    }
                    tc_commit_cg2_mc(sm.bar_w_empty + 8 * s, 3);                // frees the slot in BOTH CTAs
                    if (c + 1 == nch) tc_commit_cg2_mc(sm.bar_acc_full, 3);     // accumulator rows of BOTH CTAs complete
                }
                __syncwarp();
                TRACEW(tr, l, c, 5);
                da += (uint64_t)((KC / KS) * (4096 >> 4));
            }
            TRACE(tr, l, 2);
            if (tr) { g_trace[l][10] = (unsigned long long)(dbg_a - a0); g_trace[l][11] = (unsigned long long)(dbg_w - w0); }
        }
    }
    if (args.debug && lane0) {
        atomicAdd(&g_dbg[0], (unsigned long long)dbg_w);
        atomicAdd(&g_dbg[1], (unsigned long long)dbg_a);
        atomicAdd(&g_dbg[4], (unsigned long long)(clk() - dbg_t0));
        atomicAdd(&g_dbg[6], 1ull);
    }
}

// one arrival of this thread's WARP on A group g.  EVERY epilogue warp arrives exactly once per group and GEMM, after the chunk
// it owns in that group -- if any -- is in shared memory and after all of its TMEM reads of older accumulators: a group
// therefore completes only when all 512 threads are past the previous layer, which is what makes it safe for GEMM l+2 to
// overwrite the TMEM buffer of GEMM l.
// (with 64-column groups a thread owns two chunks per group; pub_after(cg) says whether chunk index cg = k8 / 4 of the
//  thread's sequence k8 = set, set + 4, ... closes a group, pub_group(cg) which one)
__device__ __forceinline__ constexpr bool pub_after(int cg) { return ((cg + 1) * 32) % kGroupCols == 0; }
__device__ __forceinline__ constexpr int pub_group(int cg) { return (cg * 32) / kGroupCols; }
// named barrier of the 512 epilogue threads (the two role warps never join it)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ void publish(const Smem &sm, int g) {
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();                                    // the warp's 32 fenced writes are ordered before lane 0's release-arrive:
    if ((threadIdx.x & 31) == 0) {                   // 16 arrivals per group (and CTA) instead of 512
        if (sm.pair) mbar_arrive_cluster(sm.a_ready_arrive + 8 * g);     // both CTAs report to the leader's barrier
        else mbar_arrive(sm.bar_a_ready + 8 * g);
    }
}

// (Tried and reverted in round 2: publishing the first 32 columns of the next operand on a barrier of their own, so that the first four
//  UMMAs of a layer start one epilogue chunk earlier.  The MMAs did start ~0.5 k cycles earlier, but the extra publish -- fences +
//  remote arrive on every warp's critical path -- slowed the epilogue, which is what the MMAs of the rest of the layer wait for:
//  tf32 forward 0.479 -> 0.531 ms per 262 144 samples, trace period 7.0 k -> 7.6 k cycles per layer.)
// ReLU-mask byte of an 8-column chunk from its packed bf16 activations (4 words of 2): bit i = [column 2i != 0], bit 4+i =
// [column 2i+1 != 0] for i < 4 -- 4 packed compares + 4 LOP3 + 2 instead of ~28 scalar instructions.  (v >= 0 here; bf16 keeps the
// fp32 exponent range, so bf16(v) != 0 <=> v > 0 down to 2^-134.)
__device__ __forceinline__ uint32_t relu_mask_byte(const uint4 &h) {
    const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
    auto nz = [&](uint32_t w) { return __hne2_mask(*reinterpret_cast<const __nv_bfloat162 *>(&w), z); };
    const uint32_t m = (nz(h.x) & 0x00100001u) | (nz(h.y) & 0x00200002u) | (nz(h.z) & 0x00400004u) | (nz(h.w) & 0x00800008u);
    return (m & 0xFu) | ((m >> 16) & 0xF0u);
}
// position of column i (0..7) of a chunk in that byte
__device__ __forceinline__ constexpr int relu_mask_bit(int i) { return (i >> 1) + 4 * (i & 1); }

// ---- forward epilogue.  Thread = (row, set): the 8-column chunks k8 = set, set+4, ... of every layer.
template <int NPASS, bool SAVE>
__device__ __forceinline__ void fwd_epilogue_loop(const ChainArgs &args, const Smem &sm, int num_tiles, uint32_t tmem_base, int warp) {
    const int quarter = warp & 3, set = warp >> 2;
    const int row = quarter * 32 + (threadIdx.x & 31);
    const int tid = threadIdx.x;                                   // 0 .. 511
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float *bias_all = reinterpret_cast<const float *>(args.packed + args.bias_off);
    uint32_t acc_cnt = 0, bl = 0;                                  // bl: running layer count = parity of the bias buffer
    const bool dbg_on = args.debug && threadIdx.x == 0;
    long long dbg_acc = 0;
    const long long dbg_t0 = dbg_on ? clk() : 0;
    // Biases live in shared memory, one layer ahead: thread t < 256 fetches element t of the NEXT layer's bias into a register
    // when it starts a layer and parks it in the other buffer when it is done; a named barrier of the epilogue threads (in
    // the window in which they would wait for the next accumulator anyway) publishes it.  The ncu source view had shown a
    // quarter of the epilogue's active time in the first FADD of every chunk, waiting for its two bias LDG.128.
    if (tid < 256) sm.bias[tid] = __ldg(bias_all + tid);
    epi_bar_sync();
    // x0: the (agg35, var, h32) input columns of the tile, zero padded to 80 -- the operand of GEMM 0 and the tail of the operand of
    // GEMM 5 --, fetched ONE TILE AHEAD (during layer 6 of the previous tile): the round-2 trace showed layers 0 and 5, the two with
    // the least arithmetic, taking 11 k cycles each against 6.8 k for a 256 x 256 layer, all of it the latency of these loads in front
    // of the first UMMA.  For this staging a thread owns chunks cs, cs + 4, cs + 8 (cs = lane & 3) of row 8 * warp + lane / 4 -- NOT
    // its TMEM row: a warp instruction then touches 8 rows x 64 B instead of 32 rows (528 B apart) x 16 B.  With lane = row the
    // 2.7 k uncoalesced line requests per tile slowed the UMMAs running beside them from ~100 to ~270 cycles each (trace: the MMA
    // phase of the layer during which the loads are in flight took 6.8-9 k cycles instead of 3 k; L1::no_allocate made it worse).
    const int row_s = warp * 8 + ((threadIdx.x & 31) >> 2), cs = threadIdx.x & 3;
    float x0[3][8];
    auto load_x0 = [&](int t) {
        const long g = (long)t * kTileM + row_s;
        const bool ok = t < num_tiles && g < args.m;
        const float *xr = args.XB + g * 132 + 64;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = (cs + 4 * j) * 8 + h * 4;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok && c < 68) x = __ldg(reinterpret_cast<const float4 *>(xr + c));
                x0[j][h * 4 + 0] = x.x; x0[j][h * 4 + 1] = x.y; x0[j][h * 4 + 2] = x.z; x0[j][h * 4 + 3] = x.w;
            }
    };
    load_x0(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long grow = (long)tile * kTileM + row;
        const bool valid = grow < args.m;
        const bool tr0 = (args.debug & 16) && blockIdx.x == 0 && tile == 2 * (int)gridDim.x && threadIdx.x == 0;
        const bool tr15 = (args.debug & 16) && blockIdx.x == 0 && tile == 2 * (int)gridDim.x && threadIdx.x == 15 * 32;
        __nv_bfloat16 *sv = (SAVE && valid) ? reinterpret_cast<__nv_bfloat16 *>(args.act_save) : nullptr;
        // chunk cs + 4 j of x0 -> A chunk k8_0 + cs + 4 j of row row_s (and its bf16 copy into slot `slot` of the saved activations)
        const long grow_s = (long)tile * kTileM + row_s;
        const bool save_s = SAVE && grow_s < args.m;
        // (the global stores of the bf16 copies come AFTER the publishes: fence.proxy.async is a MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in
        //  SASS and would wait for them)
        static_assert(kGroupCols == 64, "staging below publishes 64-column groups");
        auto stage_x0 = [&](int j, int k8_0) { return store_a8<NPASS>(sm.A, row_s, k8_0 + cs + 4 * j, x0[j]); };
        auto save_x0 = [&](int slot, int k8_0, const uint4 &h0, const uint4 &h1, const uint4 &h2) {
            if (!save_s) return;
            uint4 *dst = reinterpret_cast<uint4 *>(args.act_save) + ((long)(slot * 32 + k8_0 + cs) * args.slot_stride + grow_s);
            dst[0] = h0;
            dst[4 * args.slot_stride] = h1;
            if (cs < 2) dst[8 * args.slot_stride] = h2;
        };
        {   // GEMM 0 operand A[:, 0:80) = (agg35, var, h32, pad): chunks 0..9
            const uint4 h0 = stage_x0(0, 0), h1 = stage_x0(1, 0);
            publish(sm, 0);
            uint4 h2 = make_uint4(0u, 0u, 0u, 0u);
            if (cs < 2) h2 = stage_x0(2, 0);
            publish(sm, 1);
            save_x0(8, 0, h0, h1, h2);
        }
        for (int l = 0; l < kLayers; ++l, ++acc_cnt, ++bl) {
            const int l_next = l + 1 == kLayers ? 0 : l + 1;
            // The next tile's inputs (x0 is dead from layer 5 on).  While these loads are outstanding the MMA warp stalls for ~4-6 k
            // cycles wherever they are placed (measured at layers 5, 6 and 9: the issue span of that layer's UMMAs grows by that much,
            // independent of their number and shape -- the LSU path the MMA warp's barrier polls share with the epilogue warps' misses);
            // layer 6 was the cheapest of the three (tf32 forward with saves 0.544 / 0.559 / 0.553 ms per 262 144 samples).
            if (l == 6) load_x0(tile + (int)gridDim.x);
            float bias_next = 0.f;
            if (tid < 256) bias_next = __ldg(bias_all + l_next * 256 + tid);
            {
                const long long t0 = dbg_on ? clk() : 0;
                mbar_wait(sm.bar_acc_full, acc_cnt & 1);
                if (dbg_on) dbg_acc += clk() - t0;
            }
            TRACE(tr0, l, 3);
            TRACE(tr15, l, 6);
            tc_fence_after();
            const uint32_t t_acc = t_lane + (uint32_t)(l & 1) * 256;
            const float *bias = sm.bias + (bl & 1) * 256;
            if (l == 9) {
                if (set == 0) {
                    uint32_t r[8];
                    tmem_ld8_issue(t_acc, r);
                    tmem_ld_wait();
                    if (valid) {
                        float *o = args.raw + grow * args.ldr;
                        o[0] = __uint_as_float(r[0]) + bias[0];
                        o[1] = __uint_as_float(r[1]) + bias[1];
                        o[2] = __uint_as_float(r[2]) + bias[2];
                    }
                }
                tc_fence_before();
            } else if (l == 4) {
                // geometry head: columns 0..63 = features (-> A[:, 0:64) of the colour trunk), column 64 = sigma.
                // All TMEM reads of a thread come before its last publish (see the backward chain, d == 4); all of its global
                // stores come after it.
                uint32_t rs[8], ra[8], rb[8];
                if (set == 0) tmem_ld8_issue(t_acc + 64, rs);
                tmem_ld8_issue(t_acc + set * 8, ra);
                tmem_ld8_issue(t_acc + (4 + set) * 8, rb);
                float va[8], vb[8];
                {
                    const float4 a0 = *reinterpret_cast<const float4 *>(bias + set * 8), a1 = *reinterpret_cast<const float4 *>(bias + set * 8 + 4);
                    const float4 c0 = *reinterpret_cast<const float4 *>(bias + (4 + set) * 8), c1 = *reinterpret_cast<const float4 *>(bias + (4 + set) * 8 + 4);
                    const float ba[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) { va[i] = __uint_as_float(ra[i]) + ba[i]; vb[i] = __uint_as_float(rb[i]) + bb[i]; }
                }
                const uint4 fa = store_a8<NPASS>(sm.A, row, set, va), fb = store_a8<NPASS>(sm.A, row, 4 + set, vb);
                publish(sm, 0);
                // A[:, 64:144) = (agg35, var, h32, pad): chunks 8..17 -> A groups 1, 2
                const uint4 h0 = stage_x0(0, 8), h1 = stage_x0(1, 8);
                publish(sm, 1);
                uint4 h2 = make_uint4(0u, 0u, 0u, 0u);
                if (cs < 2) h2 = stage_x0(2, 8);
                publish(sm, 2);
                if (valid) {
                    if (set == 0) args.raw[grow * args.ldr + 3] = __uint_as_float(rs[0]) + bias[64];
                    if (SAVE) {
                        float4 *dst = reinterpret_cast<float4 *>(args.XB + grow * 132 + set * 8);
                        dst[0] = make_float4(va[0], va[1], va[2], va[3]);
                        dst[1] = make_float4(va[4], va[5], va[6], va[7]);
                        dst[8] = make_float4(vb[0], vb[1], vb[2], vb[3]);
                        dst[9] = make_float4(vb[4], vb[5], vb[6], vb[7]);
                        uint4 *sd = reinterpret_cast<uint4 *>(args.act_save) + ((long)(9 * 32 + set) * args.slot_stride + grow);
                        sd[0] = fa;
                        sd[4 * args.slot_stride] = fb;
                    }
                }
                save_x0(9, 8, h0, h1, h2);
            } else {
                // hidden layer: +bias, ReLU -> next A operand (and the saved activation / ReLU mask for the backward pass).
                // Fully unrolled over the thread's 8 chunks k8 = set, set + 4, ...: TMEM columns, bias and operand offsets become
                // immediates (round 2: the epilogue is what the MMAs of the next layer wait for -- ~120 instructions per chunk in
                // the rolled loop with its 64-bit index arithmetic, ~60 now).
                const int slot = l < 4 ? l : l - 1;                  // 0..3 = pts1..4, 4..7 = rgb1..4
                const long step4 = 4 * args.slot_stride;             // per chunk: 16-byte units of the bf16 save, bytes of the mask
                uint4 *sv_p = reinterpret_cast<uint4 *>(args.act_save) + ((long)(slot * 32 + set) * args.slot_stride + grow);
                uint8_t *mk_p = args.relu_mask + ((long)(slot * 32 + set) * args.slot_stride + grow);
                const float *bias_t = bias + set * 8;
                uint32_t ra[8], rb[8];
                auto process = [&](int cg, uint32_t (&r)[8]) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(bias_t + cg * 32);
                    const float4 b1 = *reinterpret_cast<const float4 *>(bias_t + cg * 32 + 4);
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
                    add2(v[0], v[1], b0.x, b0.y); add2(v[2], v[3], b0.z, b0.w);
                    add2(v[4], v[5], b1.x, b1.y); add2(v[6], v[7], b1.z, b1.w);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                    const uint4 hi = store_a8<NPASS>(sm.A, row, cg * 4 + set, v);
                    // publish FIRST: the arrive has release semantics and would otherwise wait for the global stores below
                    // (measured: 21-25 % of the epilogue's time with them in front of it)
                    if (pub_after(cg)) publish(sm, pub_group(cg));
                    TRACE(tr0 && cg == kGroupK8 / 4 - 1, l, 4);
                    TRACE(tr0 && cg == 7, l, 5);
                    TRACE(tr15 && cg == 7, l, 7);
                    if (SAVE && valid) {
                        sv_p[cg * step4] = hi;
                        mk_p[cg * step4] = (uint8_t)relu_mask_byte(hi);      // ReLU'(x) = [x > 0], from the packed bf16 words
                    }
                };
                tmem_ld8_issue(t_acc + set * 8, ra);
#pragma unroll
                for (int cg = 0; cg < 8; cg += 2) {
                    tmem_ld_wait();
                    tmem_ld8_issue(t_acc + ((cg + 1) * 4 + set) * 8, rb);
                    process(cg, ra);
                    tmem_ld_wait();
                    if (cg + 2 < 8) tmem_ld8_issue(t_acc + ((cg + 2) * 4 + set) * 8, ra);
                    process(cg + 1, rb);
                }
            }
            // park the next layer's bias (fetched at the top of this layer) in the other buffer; the barrier sits where the
            // threads would otherwise wait for the next accumulator
            if (tid < 256) sm.bias[((bl + 1) & 1) * 256 + tid] = bias_next;
            epi_bar_sync();
        }
    }
    if (dbg_on) { atomicAdd(&g_dbg[2], (unsigned long long)dbg_acc); atomicAdd(&g_dbg[5], (unsigned long long)(clk() - dbg_t0)); }
}

// ---- backward (data-gradient) epilogue.  Chain position d: 0 out^T, 1..3 rgb3..1^T, 4 rgb0^T, 5 geo^T, 6..8 pts3..1^T, 9 pts0^T
template <int NPASS>
__device__ __forceinline__ void bwd_epilogue_loop(const ChainArgs &args, const Smem &sm, int num_tiles, uint32_t tmem_base, int warp) {
    const int quarter = warp & 3, set = warp >> 2;
    const int row = quarter * 32 + (threadIdx.x & 31);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t acc_cnt = 0;
    const bool dbg_on = args.debug && threadIdx.x == 0;
    long long dbg_acc = 0;
    const long long dbg_t0 = dbg_on ? clk() : 0;
    // d raw of this thread's row (set 0 only), fetched one tile ahead: it is the operand of GEMM 0, in front of the first UMMA of a tile
    float gr[4] = {0.f, 0.f, 0.f, 0.f};
    auto load_gr = [&](int t) {
        const long g = (long)t * kTileM + row;
        const bool ok = set == 0 && t < num_tiles && g < args.m;
#pragma unroll
        for (int i = 0; i < 4; ++i) gr[i] = ok ? __ldg(args.g_raw + g * 5 + i) : 0.f;
    };
    load_gr(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long grow = (long)tile * kTileM + row;
        const bool valid = grow < args.m;
        const bool tr0 = (args.debug & 16) && blockIdx.x == 0 && tile == 2 * (int)gridDim.x && threadIdx.x == 0;
        const bool tr15 = (args.debug & 16) && blockIdx.x == 0 && tile == 2 * (int)gridDim.x && threadIdx.x == 15 * 32;
        const float g_sigma = gr[3];
        // the colour trunk's share of d(agg, var, h) (position 4) waits in registers for the geometry trunk's (position 9): chunks
        // set and set + 4 of the 68 columns, and columns 64..67 in set 0 -- one gXB store instead of store + load + store
        float gx[3][8];
        {   // GEMM 0 operand: d raw[:, 0:3] in A[:, 0:16): chunk 0 by set 0, chunk 1 (zeros) by set 1.  (The global store of the bf16
            // copy comes after the publish: fence.proxy.async would wait for it.)
            float v[8] = {gr[0], gr[1], gr[2], 0.f, 0.f, 0.f, 0.f, 0.f};
            if (set < 2) store_a8<NPASS>(sm.A, row, set, v);
            publish(sm, 0);
            if (valid && set < 2) *reinterpret_cast<uint4 *>(args.g_save + saved_off(9, set, args.slot_stride, grow)) = pack_bf16x8(v);
        }
        for (int d = 0; d < kLayers; ++d, ++acc_cnt) {
            uint32_t mw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};        // this thread's 8 ReLU-mask bytes of the layer, fetched before the wait
            if (valid && d != 9 && d != 4) {
                const int slot = d < 4 ? 7 - d : 8 - d;              // d0..3 -> R4..R1 (slots 7..4); d5..8 -> H4..H1 (slots 3..0)
                const uint8_t *mp = args.relu_mask + ((long)slot * 32 + set) * args.slot_stride + grow;
#pragma unroll
                for (int cg = 0; cg < 8; ++cg) mw[cg] = __ldg(mp + (long)cg * 4 * args.slot_stride);
            }
            {
                const long long t0 = dbg_on ? clk() : 0;
                mbar_wait(sm.bar_acc_full, acc_cnt & 1);
                if (dbg_on) dbg_acc += clk() - t0;
            }
            TRACE(tr0, d, 3);
            TRACE(tr15, d, 6);
            tc_fence_after();
            const uint32_t t_acc = t_lane + (uint32_t)(d & 1) * 256;
            if (d == 8) load_gr(tile + (int)gridDim.x);
            if (d == 9) {
                // pts0^T: 68 (+12 pad) columns, added to the colour trunk's share of d(agg,var,h) kept in gx since position 4
                uint32_t r0[8], r1[8], r2[8];
                tmem_ld8_issue(t_acc + set * 8, r0);
                tmem_ld8_issue(t_acc + (set + 4) * 8, r1);
                if (set == 0) tmem_ld8_issue(t_acc + 64, r2);
                tmem_ld_wait();
                tc_fence_before();
                if (valid) {
                    float4 *dst = reinterpret_cast<float4 *>(args.gXB + grow * 132 + 64 + set * 8);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { gx[0][i] += __uint_as_float(r0[i]); gx[1][i] += __uint_as_float(r1[i]); }
                    dst[0] = make_float4(gx[0][0], gx[0][1], gx[0][2], gx[0][3]);
                    dst[1] = make_float4(gx[0][4], gx[0][5], gx[0][6], gx[0][7]);
                    dst[8] = make_float4(gx[1][0], gx[1][1], gx[1][2], gx[1][3]);
                    dst[9] = make_float4(gx[1][4], gx[1][5], gx[1][6], gx[1][7]);
                    if (set == 0)
                        dst[16] = make_float4(gx[2][0] + __uint_as_float(r2[0]), gx[2][1] + __uint_as_float(r2[1]),
                                              gx[2][2] + __uint_as_float(r2[2]), gx[2][3] + __uint_as_float(r2[3]));
                }
            } else if (d == 4) {
                // rgb0^T: columns 0..63 = d geo features (-> operand of geo^T together with d sigma), 64..131 = d(agg,var,h).
                // Every TMEM read of a thread comes BEFORE its last publish: GEMM d+2 overwrites this buffer as soon as some
                // threads have published the first group of layer d+1, which they can only do after GEMM d+1 has consumed
                // everything published here.  Every global store comes after it.
                uint32_t r0[8], r1[8], r2[8], fa[8], fb[8];
                tmem_ld8_issue(t_acc + set * 8, fa);
                tmem_ld8_issue(t_acc + (4 + set) * 8, fb);
                tmem_ld8_issue(t_acc + 64 + set * 8, r0);
                tmem_ld8_issue(t_acc + 64 + (set + 4) * 8, r1);
                if (set == 0) tmem_ld8_issue(t_acc + 128, r2);
                tmem_ld_wait();
                float va[8], vb[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    va[i] = __uint_as_float(fa[i]); vb[i] = __uint_as_float(fb[i]);
                    gx[0][i] = __uint_as_float(r0[i]); gx[1][i] = __uint_as_float(r1[i]);
                    gx[2][i] = set == 0 ? __uint_as_float(r2[i]) : 0.f;
                }
                const uint4 ha = store_a8<NPASS>(sm.A, row, set, va), hb = store_a8<NPASS>(sm.A, row, 4 + set, vb);
                publish(sm, 0);
                // A[:, 64:80) = (d sigma, 0...): chunk 8 by set 0, chunk 9 (zeros) by set 1
                float vs[8] = {set == 0 ? g_sigma : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (set < 2) store_a8<NPASS>(sm.A, row, 8 + set, vs);
                publish(sm, 1);
                if (valid) {
                    uint4 *gs = reinterpret_cast<uint4 *>(args.g_save) + ((long)(4 * 32 + set) * args.slot_stride + grow);
                    gs[0] = ha;
                    gs[4 * args.slot_stride] = hb;
                    if (set < 2) gs[8 * args.slot_stride] = pack_bf16x8(vs);
                }
            } else {
                // through a ReLU: G = acc * (saved activation > 0) -> next A operand, and saved for the weight gradient
                const int gslot = d;                                 // g_save: 0..3 rgb3..rgb0, 4 geo, 5..8 pts3..pts0
                uint4 *gs_p = reinterpret_cast<uint4 *>(args.g_save) + ((long)(gslot * 32 + set) * args.slot_stride + grow);
                const long step4 = 4 * args.slot_stride;             // 16-byte chunks per step of this thread's k8 = set, set + 4, ...
                uint32_t ra[8], rb[8];
                auto process = [&](int cg, const uint32_t (&r)[8], uint32_t word) {
                    const uint32_t bits = word;
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = (bits & (1u << relu_mask_bit(i))) ? __uint_as_float(r[i]) : 0.f;
                    const uint4 hi = store_a8<NPASS>(sm.A, row, cg * 4 + set, v);
                    if (pub_after(cg)) publish(sm, pub_group(cg));   // before the global store (see the forward epilogue)
                    TRACE(tr0 && cg == kGroupK8 / 4 - 1, d, 4);
                    TRACE(tr0 && cg == 7, d, 5);
                    TRACE(tr15 && cg == 7, d, 7);
                    if (valid) gs_p[cg * step4] = hi;
                };
                tmem_ld8_issue(t_acc + set * 8, ra);
#pragma unroll
                for (int cg = 0; cg < 8; cg += 2) {
                    tmem_ld_wait();
                    tmem_ld8_issue(t_acc + ((cg + 1) * 4 + set) * 8, rb);
                    process(cg, ra, mw[cg]);
                    tmem_ld_wait();
                    if (cg + 2 < 8) tmem_ld8_issue(t_acc + ((cg + 2) * 4 + set) * 8, ra);
                    process(cg + 1, rb, mw[cg + 1]);
                }
            }
        }
    }
    if (dbg_on) { atomicAdd(&g_dbg[2], (unsigned long long)dbg_acc); atomicAdd(&g_dbg[5], (unsigned long long)(clk() - dbg_t0)); }
}

// ---- non-rigid chain epilogue (forward only, nothing saved)
// chunk g (8 values) of the Hann-windowed positional encoding of one point (hannw_fourier.py:27-45; the arithmetic of
// hann_pe_kernel in mlp_simt.cu: w_j * sinf / cosf(x_c * 2^j), layout [j][sin xyz, cos xyz]); zero beyond index 36
__device__ __forceinline__ void pe_chunk(const float (&x)[3], const float (&win)[6], int g, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = g * 8 + i;
        float val = 0.f;
        if (idx < 36) {
            const int j = idx / 6, r = idx - j * 6, c = r >= 3 ? r - 3 : r;
            const float a = x[c] * (float)(1 << j);
            val = win[j] * (r >= 3 ? cosf(a) : sinf(a));
        }
        v[i] = val;
    }
}

template <int NPASS>
__device__ __forceinline__ void nr_epilogue_loop(const ChainArgs &args, const Smem &sm, int num_tiles, uint32_t tmem_base, int warp) {
    const int quarter = warp & 3, set = warp >> 2;
    const int row = quarter * 32 + (threadIdx.x & 31);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float *bias_all = reinterpret_cast<const float *>(args.packed + args.bias_off);
    float win[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) win[j] = args.nr_window[j];
    uint32_t acc_cnt = 0, bl = 0;
    const int tid = threadIdx.x;
    // biases in shared memory one layer ahead, as in the forward chain (the per-chunk __ldg of the bias sat in front of every FADD)
    if (tid < 256) sm.bias[tid] = __ldg(bias_all + tid);
    epi_bar_sync();
    // The positional encoding of a tile (16 sinf / cosf per thread) is computed ONE TILE AHEAD, while the thread would otherwise wait
    // for accumulators: the point is loaded during layer 2, its encoding evaluated after the skip connection of layer 3 has consumed
    // the current one.  At the head of a tile it used to sit in front of the first UMMA (~4 k issue cycles per tile).
    float xn[3], pe0[8], pe1[8];
    auto load_point = [&](int t) {
        const long g = (long)t * kTileM + row;
        const bool ok = t < num_tiles && g < args.m;
#pragma unroll
        for (int c = 0; c < 3; ++c) xn[c] = ok ? __ldg(args.nr_xyz + g * 3 + c) : 0.f;
    };
    auto encode = [&]() {
        pe_chunk(xn, win, set, pe0);
        pe_chunk(xn, win, set + 4, pe1);                 // (all zeros for sets 2, 3: indices >= 48)
    };
    load_point(blockIdx.x);
    encode();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long grow = (long)tile * kTileM + row;
        const bool valid = grow < args.m;
        const float x[3] = {xn[0], xn[1], xn[2]};
        // GEMM 0 operand A[:, 0:48) = (pe36, pad): chunks 0..5 -> A group 0
        static_assert(kGroupCols == 64, "the non-rigid epilogue publishes 64-column groups");
        store_a8<NPASS>(sm.A, row, set, pe0);
        if (set < 2) store_a8<NPASS>(sm.A, row, set + 4, pe1);
        publish(sm, 0);
        for (int l = 0; l < 7; ++l, ++acc_cnt, ++bl) {
            const int l_next = l + 1 == 7 ? 0 : l + 1;
            float bias_next = 0.f;
            if (tid < 256) bias_next = __ldg(bias_all + l_next * 256 + tid);
            if (l == 2) load_point(tile + (int)gridDim.x);
            mbar_wait(sm.bar_acc_full, acc_cnt & 1);
            tc_fence_after();
            const uint32_t t_acc = t_lane + (uint32_t)(l & 1) * 256;
            const float *bias = sm.bias + (bl & 1) * 256;
            if (l == 6) {
                if (set == 0) {
                    uint32_t r[8];
                    tmem_ld8_issue(t_acc, r);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) args.nr_out[grow * 3 + c] = x[c] + (__uint_as_float(r[c]) + bias[c]);
                    }
                }
                tc_fence_before();
            } else {
                const float *bias_t = bias + set * 8;
                uint32_t ra[8], rb[8];
                auto process = [&](int cg, const uint32_t (&r)[8]) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(bias_t + cg * 32);
                    const float4 b1 = *reinterpret_cast<const float4 *>(bias_t + cg * 32 + 4);
                    const float bj[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = fmaxf(__uint_as_float(r[i]) + bj[i], 0.f);
                    store_a8<NPASS>(sm.A, row, cg * 4 + set, v);
                    if (pub_after(cg)) publish(sm, pub_group(cg));
                };
                tmem_ld8_issue(t_acc + set * 8, ra);
#pragma unroll
                for (int cg = 0; cg < 4; cg += 2) {
                    tmem_ld_wait();
                    tmem_ld8_issue(t_acc + ((cg + 1) * 4 + set) * 8, rb);
                    process(cg, ra);
                    tmem_ld_wait();
                    if (cg + 2 < 4) tmem_ld8_issue(t_acc + ((cg + 2) * 4 + set) * 8, ra);
                    process(cg + 1, rb);
                }
                if (l == 3) {   // skip connection: A[:, 128:176) = (pe36, pad): chunks 16..21 -> A group 2
                    store_a8<NPASS>(sm.A, row, 16 + set, pe0);
                    if (set < 2) store_a8<NPASS>(sm.A, row, 20 + set, pe1);
                    publish(sm, 2);
                    encode();                            // the NEXT tile's encoding (xn was loaded during layer 2)
                }
            }
            if (tid < 256) sm.bias[((bl + 1) & 1) * 256 + tid] = bias_next;
            epi_bar_sync();
        }
    }
}

template <int NPASS, int CHAIN, int CG, bool SAVE = false>
__global__ void __launch_bounds__(kThreads, 1) mlp_chain_tc_kernel(const __grid_constant__ ChainArgs args) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int kABytes = kAPartBytes * (NPASS == 1 ? 1 : 2);     // bf16: 64 KB; hi+lo or tf32: 128 KB
    constexpr int kRingBytes = kStages * kStageBytes;
    static_assert(kStages * kStageBytes == PairRing<NPASS>::kStagesN * PairRing<NPASS>::kBytes, "both ring geometries use the same bytes");
    constexpr int kNStages = CG == 2 ? PairRing<NPASS>::kStagesN : kStages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kABytes + kRingBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 30);
    Smem sm;
    sm.A = smem;
    sm.W = smem + kABytes;
    sm.bias = reinterpret_cast<float *>(smem + kABytes + kRingBytes + 256);
    sm.bar_w_full = smem_u32(bars);
    sm.bar_w_empty = smem_u32(bars + kNStages);
    sm.bar_w_peer = smem_u32(bars + 2 * kNStages);                  // (pair mode only)
    sm.bar_a_ready = smem_u32(bars + 3 * kNStages);
    sm.bar_acc_full = smem_u32(bars + 3 * kNStages + kGroups);
    static_assert(3 * 6 + kGroups + 1 <= 30, "barrier block");
    sm.pair = CG == 2;
    sm.a_ready_arrive = CG == 2 ? map_to_cta(sm.bar_a_ready, 0) : sm.bar_a_ready;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // both CTAs of a cluster must walk weight streams of equal length: the tile count is padded to an even number and
    // a padding tile simply has no valid rows
    const int num_tiles = ((args.m + kTileM - 1) / kTileM + 1) & ~1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kNStages; ++s) {
            mbar_init(sm.bar_w_full + 8 * s, 1);
            mbar_init(sm.bar_w_empty + 8 * s, CG == 2 ? 1 : 2);     // pair mode: one multicast commit of the leader covers both smems
            if (CG == 2) mbar_init(sm.bar_w_peer + 8 * s, 1);
        }
        for (int g = 0; g < kGroups; ++g) mbar_init(sm.bar_a_ready + 8 * g, (CG == 2 ? 2 : 1) * (kEpiThreads / 32));
        mbar_init(sm.bar_acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // the peer's barriers are initialised before anything is signalled to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4 * kEpiSets) {
        if (lane == 0) {
            if (CG == 2) producer_loop_pair<NPASS>(args, sm, num_tiles);
            else producer_loop<NPASS>(args, sm, num_tiles);
        }
    } else if (warp == 4 * kEpiSets + 1) {
        if (CG == 2) {
            if (cluster_ctarank() == 0) mma_loop_pair<NPASS>(args, sm, num_tiles, tmem_base);
            else if (lane == 0) relay_loop_pair<NPASS>(args, sm, num_tiles);
        } else {
            mma_loop<NPASS>(args, sm, num_tiles, tmem_base);      // whole warp, one elected lane issues
        }
    } else {
        if (CHAIN == 0) fwd_epilogue_loop<NPASS, SAVE>(args, sm, num_tiles, tmem_base, warp);
        else if (CHAIN == 1) bwd_epilogue_loop<NPASS>(args, sm, num_tiles, tmem_base, warp);
        else nr_epilogue_loop<NPASS>(args, sm, num_tiles, tmem_base, warp);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // nobody leaves while the peer may still signal into this CTA
    if (warp == 0) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

template <int NPASS, int CHAIN, int CG, bool SAVE = false>
int launch_chain_cg(const ChainArgs &a, cudaStream_t st) {
    constexpr int kABytes = kAPartBytes * (NPASS == 1 ? 1 : 2);
    const int smem_bytes = kABytes + kStages * kStageBytes + 256 + 2048;     // + barriers + the bias double buffer
    static bool configured = false;
    if (!configured) {
        OCC_CUDA(cudaFuncSetAttribute(mlp_chain_tc_kernel<NPASS, CHAIN, CG, SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = true;
    }
    int dev = 0, sms = 148;
    OCC_CUDA(cudaGetDevice(&dev));
    OCC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles = ((a.m + kTileM - 1) / kTileM + 1) & ~1;
    const int grid = tiles < (sms & ~1) ? tiles : (sms & ~1);          // CTA pairs
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OCC_CUDA(cudaLaunchKernelEx(&cfg, mlp_chain_tc_kernel<NPASS, CHAIN, CG, SAVE>, a));
    return OCCNERF_OK;
}

// cta_pair: 0 = cta_group::1 CTAs sharing the weight stream by multicast, 1 = cta_group::2 (weights packed with cta_pair = 1)
template <int NPASS, int CHAIN>
int launch_chain(const ChainArgs &a, cudaStream_t st, int cta_pair = 0) {
    if (CHAIN == 0) {      // the forward chain comes with and without the activation / ReLU-mask saves compiled in
        if (a.act_dtype) return cta_pair ? launch_chain_cg<NPASS, 0, 2, true>(a, st) : launch_chain_cg<NPASS, 0, 1, true>(a, st);
        return cta_pair ? launch_chain_cg<NPASS, 0, 2, false>(a, st) : launch_chain_cg<NPASS, 0, 1, false>(a, st);
    }
    if (cta_pair) return launch_chain_cg<NPASS, CHAIN, 2>(a, st);
    return launch_chain_cg<NPASS, CHAIN, 1>(a, st);
}

// Debug micro-benchmark: `iters` back-to-back tcgen05.mma (M=128, N=n, K=16, bf16, SS operands at fixed shared-memory
// addresses, one accumulator) per CTA, timed with clock64 between the first issue and the arrival of the final commit.
// Gives the issue-to-completion rate of the tensor pipe for exactly the instruction shape the chain kernels use.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int n, int tf32, unsigned long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 65536 + 32768);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 2);
    for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 65536);
        const uint32_t idesc = tf32 ? instr_desc_tf32(n) : instr_desc(n), b_lbo = (uint32_t)(n / 8) * 128;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint64_t da = smem_desc(a_base + (uint32_t)(i & 15) * 4096, 2048, 128);
            const uint64_t db = smem_desc(b_base + (uint32_t)(i & 1) * 2 * b_lbo, b_lbo, 128);
            if (tf32) tc_mma_tf32(tmem_base + (uint32_t)(i & 1) * 256, da, db, idesc, i > 1 ? 1u : 0u);
            else tc_mma(tmem_base + (uint32_t)(i & 1) * 256, da, db, idesc, i > 1 ? 1u : 0u);
        }
        tc_commit(smem_u32(bar));
        mbar_wait(smem_u32(bar), 0);
        out[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

int g_host_debug = -1;

void fill_layout(ChainArgs &a, int n_pass, int chain, const void *packed, int pair = 0) {
    static const int debug = getenv("OCCNERF_MLP_DEBUG") ? atoi(getenv("OCCNERF_MLP_DEBUG")) : 0;
    a.debug = g_host_debug >= 0 ? g_host_debug : debug;
    const PackedLayout pl = packed_layout(n_pass, chain, pair);
    for (int l = 0; l < kLayers; ++l) a.w_off[l] = pl.w_off[l];
    a.bias_off = pl.bias_off;
    a.chain = chain;
    a.packed = (const unsigned char *)packed;
}

}  // namespace

// debug only: how many clusters of `cluster_size` CTAs of the tc3 forward chain kernel the device can hold at once
extern "C" int occnerf_mlp_debug_max_clusters(int cluster_size) {
    constexpr int smem_bytes = 2 * kAPartBytes + kStages * kStageBytes + 256 + 2048;
    if (cudaFuncSetAttribute(mlp_chain_tc_kernel<3, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return -1;
    if (cluster_size > 8 && cudaFuncSetAttribute(mlp_chain_tc_kernel<3, 0, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cluster_size * cluster_size);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, mlp_chain_tc_kernel<3, 0, 1>, &cfg) != cudaSuccess) { cudaGetLastError(); return -3; }
    return n;
}

// debug only: cycles per tcgen05.mma (M=128, N=n, K=16) measured on every SM at once; out_dev [148+] device u64
extern "C" int occnerf_mlp_debug_mma_rate(int iters, int n, int tf32, unsigned long long *out_dev, int ctas, occnerf_stream_t stream) {
    OCC_CHECK_ARG(out_dev && iters >= 2 && n >= 16 && n <= 256 && n % 16 == 0 && ctas >= 1, "mlp_debug_mma_rate: bad arguments");
    const int smem_bytes = 65536 + 32768 + 64;
    OCC_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    mma_rate_kernel<<<ctas, 128, smem_bytes, (cudaStream_t)stream>>>(iters, n, tf32, out_dev);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

// debug only: overrides OCCNERF_MLP_DEBUG for the following launches (debug < 0: back to the environment's value)
extern "C" int occnerf_mlp_debug_set(int debug) {
    g_host_debug = debug;
    return OCCNERF_OK;
}

// debug only: reads (and optionally clears) the stall counters described at g_dbg
// debug only: the per-layer time stamps of g_trace (16 x 12 u64)
extern "C" int occnerf_mlp_debug_trace(unsigned long long *host192) {
    OCC_CUDA(cudaDeviceSynchronize());
    OCC_CHECK_ARG(host192, "mlp_debug_trace: null pointer");
    OCC_CUDA(cudaMemcpyFromSymbol(host192, g_trace, sizeof(unsigned long long) * 16 * 12));
    return OCCNERF_OK;
}

// debug only: the weight-stream round-trip stamps of g_trace_w (2 x 8 x 6 u64)
extern "C" int occnerf_mlp_debug_trace_w(unsigned long long *host96) {
    OCC_CUDA(cudaDeviceSynchronize());
    OCC_CHECK_ARG(host96, "mlp_debug_trace_w: null pointer");
    OCC_CUDA(cudaMemcpyFromSymbol(host96, g_trace_w, sizeof(unsigned long long) * 96));
    return OCCNERF_OK;
}

extern "C" int occnerf_mlp_debug_counters(unsigned long long *host8, int reset) {
    OCC_CUDA(cudaDeviceSynchronize());
    if (host8) OCC_CUDA(cudaMemcpyFromSymbol(host8, g_dbg, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        OCC_CUDA(cudaMemcpyToSymbol(g_dbg, z, sizeof(z)));
    }
    return OCCNERF_OK;
}

extern "C" long occnerf_mlp_packed_bytes(int n_pass, int chain) {
    if (!valid_pass(n_pass) || chain < 0 || chain > 2) return -1;
    const long a = packed_layout(n_pass, chain, 0).total, b = packed_layout(n_pass, chain, 1).total;
    return a > b ? a : b;                       // (the pair-mode image pads K to its larger chunks)
}

extern "C" int occnerf_mlp_pack_weights(const occnerf_mlp_params *p_host, int n_pass, int chain, int cta_pair, void *packed,
                                        occnerf_stream_t stream) {
    OCC_CHECK_ARG(p_host && packed, "mlp_pack_weights: null pointer");
    OCC_CHECK_ARG(valid_pass(n_pass), "mlp_pack_weights: n_pass=%d (supported: 1 bf16, 2 tf32, 3 split-bf16)", n_pass);
    OCC_CHECK_ARG(chain == 0 || chain == 1, "mlp_pack_weights: chain=%d (0 forward, 1 backward)", chain);
    for (int l = 0; l < kLayers; ++l) OCC_CHECK_ARG(p_host->w[l] && p_host->b[l], "mlp_pack_weights: layer %d has a null pointer", l);
    const PackedLayout pl = packed_layout(n_pass, chain, cta_pair ? 1 : 0);
    DevLayout L;
    for (int l = 0; l < kLayers; ++l) L.w_off[l] = pl.w_off[l];
    L.bias_off = pl.bias_off;
    dim3 grid(occ_div_up(256 * 256, 256), kLayers);
    pack_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p_host, L, n_pass, chain, cta_pair ? 1 : 0, (unsigned char *)packed);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_mlp_forward_tc(float *XB, int m, const void *packed, int n_pass, int cta_pair, float *raw, int ldr, void *act_save,
                                      int act_dtype, long slot_stride, void *relu_mask, occnerf_stream_t stream) {
    if (m == 0) return OCCNERF_OK;
    OCC_CHECK_ARG(XB && packed && raw, "mlp_forward_tc: null pointer");
    OCC_CHECK_ARG(valid_pass(n_pass), "mlp_forward_tc: n_pass=%d (supported: 1 bf16, 2 tf32, 3 split-bf16)", n_pass);
    OCC_CHECK_ARG(m > 0 && ldr >= 4, "mlp_forward_tc: m=%d ldr=%d", m, ldr);
    OCC_CHECK_ARG((act_dtype == 0 || act_dtype == 2) && (act_dtype == 0 || (act_save && relu_mask)),
                  "mlp_forward_tc: act_dtype=%d (0 none, 2 bf16 chunk-major + ReLU mask; both buffers required)", act_dtype);
    OCC_CHECK_ARG(((uintptr_t)XB & 15) == 0 && ((uintptr_t)packed & 15) == 0 && ((uintptr_t)act_save & 15) == 0,
                  "mlp_forward_tc: XB/packed/act_save must be 16-byte aligned");
    ChainArgs a = {};
    a.m = m;
    fill_layout(a, n_pass, 0, packed, cta_pair ? 1 : 0);
    OCC_CHECK_ARG(act_dtype == 0 || slot_stride >= m, "mlp_forward_tc: slot_stride=%ld < m=%d", slot_stride, m);
    a.XB = XB; a.raw = raw; a.ldr = ldr; a.act_save = act_save; a.act_dtype = act_dtype; a.slot_stride = slot_stride;
    a.relu_mask = (uint8_t *)relu_mask;
    OCC_CHECK_ARG(((uintptr_t)relu_mask & 15) == 0, "mlp_forward_tc: relu_mask must be 16-byte aligned");
    return n_pass == 1 ? launch_chain<1, 0>(a, (cudaStream_t)stream, cta_pair) : n_pass == 2 ? launch_chain<2, 0>(a, (cudaStream_t)stream, cta_pair)
                                                                                : launch_chain<3, 0>(a, (cudaStream_t)stream, cta_pair);
}

extern "C" int occnerf_mlp_backward_tc(const float *g_raw, int m, const void *packed_bwd, int n_pass, int cta_pair, const void *relu_mask,
                                       float *gXB, void *g_save, long slot_stride, occnerf_stream_t stream) {
    if (m == 0) return OCCNERF_OK;
    OCC_CHECK_ARG(g_raw && packed_bwd && relu_mask && gXB && g_save, "mlp_backward_tc: null pointer");
    OCC_CHECK_ARG(valid_pass(n_pass), "mlp_backward_tc: n_pass=%d (supported: 1 bf16, 2 tf32, 3 split-bf16)", n_pass);
    OCC_CHECK_ARG(((uintptr_t)gXB & 15) == 0 && ((uintptr_t)packed_bwd & 15) == 0 && ((uintptr_t)relu_mask & 15) == 0 &&
                  ((uintptr_t)g_save & 15) == 0, "mlp_backward_tc: buffers must be 16-byte aligned");
    ChainArgs a = {};
    a.m = m;
    fill_layout(a, n_pass, 1, packed_bwd, cta_pair ? 1 : 0);
    OCC_CHECK_ARG(slot_stride >= m, "mlp_backward_tc: slot_stride=%ld < m=%d", slot_stride, m);
    a.g_raw = g_raw; a.relu_mask = (uint8_t *)relu_mask; a.gXB = gXB; a.g_save = (__nv_bfloat16 *)g_save; a.slot_stride = slot_stride;
    return n_pass == 1 ? launch_chain<1, 1>(a, (cudaStream_t)stream, cta_pair) : n_pass == 2 ? launch_chain<2, 1>(a, (cudaStream_t)stream, cta_pair)
                                                                                : launch_chain<3, 1>(a, (cudaStream_t)stream, cta_pair);
}

// ---- non-rigid motion MLP on the same chain machinery (chain 2)
extern "C" int occnerf_nonrigid_pack_weights(const void *const *w7_host, const void *const *b7_host, const float *cond_dev, int n_pass,
                                             int cta_pair, void *packed, occnerf_stream_t stream) {
    OCC_CHECK_ARG(w7_host && b7_host && packed, "nonrigid_pack_weights: null pointer");
    OCC_CHECK_ARG(valid_pass(n_pass), "nonrigid_pack_weights: n_pass=%d (supported: 1 bf16, 2 tf32, 3 split-bf16)", n_pass);
    NrParams P;
    for (int l = 0; l < 7; ++l) {
        OCC_CHECK_ARG(w7_host[l] && b7_host[l], "nonrigid_pack_weights: layer %d has a null pointer", l);
        P.w[l] = (const float *)w7_host[l];
        P.b[l] = (const float *)b7_host[l];
    }
    P.cond = cond_dev;
    const PackedLayout pl = packed_layout(n_pass, 2, cta_pair ? 1 : 0);
    DevLayout L;
    for (int l = 0; l < kLayers; ++l) L.w_off[l] = pl.w_off[l];
    L.bias_off = pl.bias_off;
    dim3 grid(occ_div_up(176 * 128, 256), 7);
    pack_nr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P, L, n_pass, cta_pair ? 1 : 0, (unsigned char *)packed);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_nonrigid_forward_tc(const float *xyz, const float *window6_host, int m, const void *packed, int n_pass, int cta_pair,
                                           float *out, occnerf_stream_t stream) {
    if (m == 0) return OCCNERF_OK;
    OCC_CHECK_ARG(xyz && window6_host && packed && out && m > 0, "nonrigid_forward_tc: null pointer / m=%d", m);
    OCC_CHECK_ARG(valid_pass(n_pass), "nonrigid_forward_tc: n_pass=%d (supported: 1 bf16, 2 tf32, 3 split-bf16)", n_pass);
    OCC_CHECK_ARG(((uintptr_t)packed & 15) == 0, "nonrigid_forward_tc: packed must be 16-byte aligned");
    ChainArgs a = {};
    a.m = m;
    fill_layout(a, n_pass, 2, packed, cta_pair ? 1 : 0);
    a.nr_xyz = xyz; a.nr_out = out;
    for (int j = 0; j < 6; ++j) a.nr_window[j] = window6_host[j];
    cudaStream_t st = (cudaStream_t)stream;
    return n_pass == 1 ? launch_chain<1, 2>(a, st, cta_pair) : n_pass == 2 ? launch_chain<2, 2>(a, st, cta_pair) : launch_chain<3, 2>(a, st, cta_pair);
}
