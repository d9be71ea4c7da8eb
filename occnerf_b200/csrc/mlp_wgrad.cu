// K3c: weight (and bias) gradients of the canonical MLP on tcgen05/TMEM:  dW_l[n,k] = sum_m G_l[m,n] * X_l[m,k].
//
// The contraction runs over the SAMPLE axis, so both operands are "MN-major" for the tensor core.  The fused forward (X_l)
// and data-gradient (G_l) kernels leave their bf16 tensors in HBM chunk-major, [slot][feature/8][sample][8 features]
// (the layout in which their epilogue stores coalesce); a TMA box of 8 features x 64 samples x 8 chunks
// (cp.async.bulk.tensor.3d, no swizzle) lands in shared memory as [chunk][sample][16 B], which IS the canonical
// no-swizzle MN-major UMMA layout: 8 consecutive samples x 16 B = one 128-byte core matrix, core matrices 128 B apart
// along K (samples), 1024 B apart along MN (feature chunks).  No transposes anywhere.
//
// One persistent CTA per SM; for every layer a CTA owns a contiguous range of 64-sample tiles and accumulates its
// partial dW (256 x 256 fp32 = 2 x 128 TMEM lanes x 256 columns = all 512 TMEM columns) over that range, then adds it
// into the global fp32 result with red.global.add.v4.f32.  While the tensor core works, the four otherwise idle
// epilogue warps read the G tiles from shared memory and build the bias gradient (column sums) for free.
//
// Roofline: per layer 2 x M x 256 x 2 B are read once (HBM-bound: the MMAs of a 64-sample stage take ~1k cycles but
// the stage is 64 KB), i.e. ~8 GB per 786k-sample step.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kStages = 3;
constexpr int kTileS = 64;                          // samples per stage
constexpr int kBlockBytes = kTileS * 128;           // one 64-feature x 64-sample block = 8 chunks x 64 samples x 16 B = 8 KB
constexpr int kOperandBytes = 4 * kBlockBytes;      // up to 256 features
constexpr int kStageBytes = 2 * kOperandBytes;      // G tile + X tile = 64 KB
constexpr int kLayers = 10;
constexpr uint32_t kSpinLimit = 1u << 27;

struct LayerDesc { int g_slot, n_pad, x_slot, k_pad; };
// forward layer l: (G slot in g_save, padded output width, X slot in act, padded input width)
__constant__ LayerDesc c_layers[kLayers] = {
    {8, 256, 8, 80},    // pts0 : X = (agg35,var,h32,pad)
    {7, 256, 0, 256},   // pts1
    {6, 256, 1, 256},   // pts2
    {5, 256, 2, 256},   // pts3
    {4, 80, 3, 256},    // geo  : G columns 0..63 features, 64 sigma
    {3, 256, 9, 144},   // rgb0 : X = (geo64, agg35, var, h32, pad)
    {2, 256, 4, 256},   // rgb1
    {1, 256, 5, 256},   // rgb2
    {0, 256, 6, 256},   // rgb3
    {9, 16, 7, 256},    // out  : G = d raw[:, :3]
};

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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// MN-major, no swizzle (INTERLEAVE): core matrix = 8 samples (K) x 8 features (16 B, MN-contiguous) = 128 B;
// LBO = stride between core matrices along K (next 8 samples) = 128 B, SBO = stride along MN (next 8-feature chunk) = 1024 B
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((128 >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((kTileS * 16 >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16: D=f32, A=B=bf16, both MN-major, M=128
__device__ __forceinline__ uint32_t instr_desc_mn(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct WgradArgs {
    int m;                  // samples
    long slot_stride;       // rows per slot (multiple of 64, rows >= m zero-filled)
    float *dW;              // [10][256][256] fp32, accumulated (caller zeroes)
    float *dB;              // [10][256] fp32, accumulated (caller zeroes)
};

__global__ void __launch_bounds__(kThreads, 1)
mlp_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x, const WgradArgs args) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
    const uint32_t bar_acc_full = smem_u32(bars + 2 * kStages), bar_acc_free = smem_u32(bars + 2 * kStages + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_total = (int)((args.m + kTileS - 1) / kTileS);
    // contiguous tile range of this CTA (the same for every layer)
    const int t0 = (int)((long)tiles_total * blockIdx.x / gridDim.x), t1 = (int)((long)tiles_total * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1 + 128); }
        mbar_init(bar_acc_full, 1);
        mbar_init(bar_acc_free, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (t1 > t0) {
        if (warp == 4) {
            // ================= TMA producer =================
            if (lane == 0) {
                uint32_t it = 0;
                for (int l = 0; l < kLayers; ++l) {
                    const LayerDesc L = c_layers[l];
                    const int gb = (L.n_pad <= 128) ? 2 : 4;              // G feature blocks (UMMA M = 128 needs two)
                    const int xb = (L.k_pad + 63) / 64;                   // X feature blocks
                    for (int t = t0; t < t1; ++t, ++it) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                        mbar_wait(bar_empty + 8 * s, ph ^ 1);
                        mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(gb + xb) * kBlockBytes);
                        const uint32_t dst = smem_u32(smem + s * kStageBytes);
                        const int row = t * (kTileS / 32);              // in units of 32 rows (see make_map)
                        for (int b = 0; b < gb; ++b) tma_load_3d(dst + b * kBlockBytes, &map_g, 0, row, L.g_slot * 32 + b * 8, bar_full + 8 * s);
                        for (int b = 0; b < xb; ++b) tma_load_3d(dst + kOperandBytes + b * kBlockBytes, &map_x, 0, row, L.x_slot * 32 + b * 8, bar_full + 8 * s);
                    }
                }
            }
        } else if (warp == 5) {
            // ================= MMA issuer =================
            if (lane == 0) {
                uint32_t it = 0;
                for (int l = 0; l < kLayers; ++l) {
                    const LayerDesc L = c_layers[l];
                    const int halves = (L.n_pad <= 128) ? 1 : 2;
                    const uint32_t idesc = instr_desc_mn(L.k_pad);
                    if (l > 0) { mbar_wait(bar_acc_free, (l - 1) & 1); tc_fence_after(); }   // previous layer flushed out of TMEM
                    uint32_t first = 1;
                    for (int t = t0; t < t1; ++t, ++it) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                        mbar_wait(bar_full + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t g_base = smem_u32(smem + s * kStageBytes), x_base = g_base + kOperandBytes;
#pragma unroll
                        for (int k16 = 0; k16 < kTileS / 16; ++k16) {
                            const uint64_t db = smem_desc_mn(x_base + k16 * 256);          // 16 samples = 2 core matrices along K
                            for (int h = 0; h < halves; ++h) {
                                const uint64_t da = smem_desc_mn(g_base + h * 2 * kBlockBytes + k16 * 256);
                                tc_mma(tmem_base + h * 256, da, db, idesc, first ? 0u : 1u);
                            }
                            first = 0;
                        }
                        tc_commit(bar_empty + 8 * s);
                    }
                    tc_commit(bar_acc_full);
                }
            }
        } else {
            // ================= epilogue warps: bias sums while the MMAs run, then flush dW =================
            const int tid = threadIdx.x;                       // 0..127; owns features 2*tid, 2*tid+1 for the bias sums
            const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
            uint32_t it = 0;
            for (int l = 0; l < kLayers; ++l) {
                const LayerDesc L = c_layers[l];
                const int halves = (L.n_pad <= 128) ? 1 : 2;
                const bool bias_cols = 2 * tid < L.n_pad;
                float s0 = 0.f, s1 = 0.f;
                const int chunk = (2 * tid) >> 3, within = ((2 * tid) & 7) * 2;   // feature chunk (1 KB slab) and byte offset in its 16 B
                for (int t = t0; t < t1; ++t, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(bar_full + 8 * s, ph);
                    if (bias_cols) {
                        const unsigned char *g = smem + s * kStageBytes + chunk * (kTileS * 16) + within;
#pragma unroll 8
                        for (int r = 0; r < kTileS; ++r) {
                            const int rr = (r + chunk) & (kTileS - 1);       // rotate per chunk: the 8 chunks of a warp hit different banks
                            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162 *>(g + rr * 16);
                            const float2 f = __bfloat1622float2(v);
                            s0 += f.x; s1 += f.y;
                        }
                    }
                    mbar_arrive(bar_empty + 8 * s);
                }
                if (bias_cols) {
                    atomicAdd(args.dB + l * 256 + 2 * tid, s0);
                    atomicAdd(args.dB + l * 256 + 2 * tid + 1, s1);
                }
                // flush the accumulator: TMEM lane = output row n (per half), column = k
                mbar_wait(bar_acc_full, l & 1);
                tc_fence_after();
                for (int h = 0; h < halves; ++h) {
                    const int n = h * 128 + tid;
                    float *drow = args.dW + ((long)l * 256 + n) * 256;
                    for (int cg = 0; cg < (L.k_pad + 31) / 32; ++cg) {
                        uint32_t r[32];
                        tmem_ld32(t_lane + h * 256 + cg * 32, r);
                        if (n < L.n_pad) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (cg * 32 + i * 4 < L.k_pad)
                                    red_add_v4(drow + cg * 32 + i * 4, __uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                               __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_acc_free);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap *map, const void *base, long rows, int slots) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        OCC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        OCC_CHECK_ARG(fn && q == cudaDriverEntryPointSuccess, "mlp_wgrad: cuTensorMapEncodeTiled is not available from this driver");
        encode = (EncodeTiledFn)fn;
    }
    // chunk-major [slots*32][rows][8].  Rows of one chunk are contiguous (16 B each), so the map folds 32 of them into the
    // innermost dimension (512-byte bursts instead of 16-byte ones): dim0 = 32 rows x 8 features, dim1 = row / 32,
    // dim2 = slot*32 + chunk.  The box (256, 2, 8) lands in shared memory as [chunk][64 rows][16 B] all the same.
    const cuuint64_t dims[3] = {256, (cuuint64_t)rows / 32, (cuuint64_t)slots * 32};
    const cuuint64_t strides[2] = {256 * sizeof(__nv_bfloat16), (cuuint64_t)rows * 8 * sizeof(__nv_bfloat16)};
    const cuuint32_t box[3] = {256, (cuuint32_t)kTileS / 32, 8};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    OCC_CHECK_ARG(r == CUDA_SUCCESS, "mlp_wgrad: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return OCCNERF_OK;
}

}  // namespace

extern "C" int occnerf_mlp_wgrad_tc(const void *g_save, const void *act_bf16, int m, long slot_stride, float *dW, float *dB,
                                    occnerf_stream_t stream) {
    if (m == 0) return OCCNERF_OK;
    OCC_CHECK_ARG(g_save && act_bf16 && dW && dB, "mlp_wgrad_tc: null pointer");
    OCC_CHECK_ARG(slot_stride >= m && slot_stride % kTileS == 0, "mlp_wgrad_tc: slot_stride=%ld must be a multiple of %d and >= m=%d",
                  slot_stride, kTileS, m);
    OCC_CHECK_ARG(slot_stride < (1l << 31), "mlp_wgrad_tc: too many rows for 32-bit TMA coordinates");
    OCC_CHECK_ARG((((uintptr_t)g_save | (uintptr_t)act_bf16 | (uintptr_t)dW) & 15) == 0, "mlp_wgrad_tc: buffers must be 16-byte aligned");
    CUtensorMap map_g, map_x;
    if (int e = make_map(&map_g, g_save, slot_stride, 10)) return e;
    if (int e = make_map(&map_x, act_bf16, slot_stride, 10)) return e;
    const int smem_bytes = kStages * kStageBytes + 256;
    static bool configured = false;
    if (!configured) {
        OCC_CUDA(cudaFuncSetAttribute(mlp_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = true;
    }
    int dev = 0, sms = 148;
    OCC_CUDA(cudaGetDevice(&dev));
    OCC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles = (m + kTileS - 1) / kTileS;
    WgradArgs a;
    a.m = m; a.slot_stride = slot_stride; a.dW = dW; a.dB = dB;
    mlp_wgrad_tc_kernel<<<tiles < sms ? tiles : sms, kThreads, smem_bytes, (cudaStream_t)stream>>>(map_g, map_x, a);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
