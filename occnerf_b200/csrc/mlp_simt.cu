// Exact-fp32 MLP path: a strided SIMT SGEMM with fused bias / ReLU / ReLU-mask / accumulate epilogues,
// a masked column sum (bias gradients) and the Hann-windowed positional encoding.
//
// This is the "precise" mode of the canonical MLP and of the non-rigid offset MLP
// (core/nets/occnerf/canonical_mlps/occnerf_mlp.py:183-199, non_rigid_motion_mlps/mlp_offset.py:45-62,
// embedders/hannw_fourier.py:27-45): the same fp32 multiply-adds the reference's nn.Linear layers do through
// cuBLAS sgemm, in one generic kernel so that forward (X.W^T), data gradient (dY.W) and weight gradient
// (dY^T.X, split over the sample axis) all share it.  The tcgen05 kernel in mlp_tc.cu is the fast path.
//
// C[i,j] = epi( sum_r A(i,r) * B(r,j) ),  A(i,r) = A[i*sAi + r*sAr],  B(r,j) = B[r*sBr + j*sBj].
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile split 4+4 in both directions so that the LDS.128
// operand reads are bank-conflict free.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, kThreads = 256, kPad = 4;

// tile loader: T = "r is the contiguous axis" (needs a transpose into the [BK][BM] shared layout)
template <bool T, bool VEC>
__device__ __forceinline__ void load_tile(float (*sm)[BM + kPad], const float *__restrict__ P, long s_outer, long s_r,
                                          int o0, int r0, int n_outer, int r_end) {
    // logical element (o, r): P[o*s_outer + r*s_r]; o in [o0, o0+BM), r in [r0, r0+BK)
    if constexpr (VEC) {
        if constexpr (T) {   // s_r == 1: float4 along r, scatter into 4 smem rows
            for (int v = threadIdx.x; v < BM * BK / 4; v += kThreads) {
                const int o = v / (BK / 4), r4 = (v % (BK / 4)) * 4;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (o0 + o < n_outer && r0 + r4 < r_end)   // r_end % 4 == 0 in VEC mode
                    x = __ldg(reinterpret_cast<const float4 *>(P + (long)(o0 + o) * s_outer + r0 + r4));
                sm[r4 + 0][o] = x.x; sm[r4 + 1][o] = x.y; sm[r4 + 2][o] = x.z; sm[r4 + 3][o] = x.w;
            }
        } else {             // s_outer == 1: float4 along o, direct
            for (int v = threadIdx.x; v < BM * BK / 4; v += kThreads) {
                const int r = v / (BM / 4), o4 = (v % (BM / 4)) * 4;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + r < r_end && o0 + o4 < n_outer)   // n_outer % 4 == 0 in VEC mode
                    x = __ldg(reinterpret_cast<const float4 *>(P + (long)(r0 + r) * s_r + o0 + o4));
                *reinterpret_cast<float4 *>(&sm[r][o4]) = x;
            }
        }
    } else {
        for (int v = threadIdx.x; v < BM * BK; v += kThreads) {
            int o, r;
            if constexpr (T) { o = v / BK; r = v % BK; } else { r = v / BM; o = v % BM; }
            float x = 0.f;
            if (o0 + o < n_outer && r0 + r < r_end) x = __ldg(P + (long)(o0 + o) * s_outer + (long)(r0 + r) * s_r);
            sm[r][o] = x;
        }
    }
}

template <bool AT, bool BT, bool VEC>
__global__ void __launch_bounds__(kThreads)
sgemm_kernel(const float *__restrict__ A, long sAi, long sAr, const float *__restrict__ B, long sBr, long sBj,
             float *__restrict__ C, long ldc, const float *__restrict__ bias, const float *__restrict__ mask,
             long ldmask, int Mi, int Nj, int Kr, int flags, int k_per_split) {
    __shared__ __align__(16) float As[BK][BM + kPad];
    __shared__ __align__(16) float Bs[BK][BN + kPad];
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int r_begin = blockIdx.z * k_per_split, r_end = min(Kr, r_begin + k_per_split);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

    for (int r0 = r_begin; r0 < r_end; r0 += BK) {
        load_tile<AT, VEC>(As, A, sAi, sAr, i0, r0, Mi, r_end);
        load_tile<BT, VEC>(Bs, B, sBj, sBr, j0, r0, Nj, r_end);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int i = i0 + (a < 4 ? ty * 4 + a : 64 + ty * 4 + (a - 4));
        if (i >= Mi) continue;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int j = j0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + (b - 4));
            if (j >= Nj) continue;
            float v = acc[a][b];
            float *c = C + (long)i * ldc + j;
            if (split) {   // partial sums: epilogue terms are applied by split 0 only; ReLU is not allowed here
                if (blockIdx.z == 0 && (flags & OCCNERF_GEMM_BIAS)) v += __ldg(bias + j);
                atomicAdd(c, v);
            } else {
                if (flags & OCCNERF_GEMM_BIAS) v += __ldg(bias + j);
                if (flags & OCCNERF_GEMM_ACCUM) v += *c;
                if (flags & OCCNERF_GEMM_RELU) v = fmaxf(v, 0.f);
                if (flags & OCCNERF_GEMM_RELUMASK) v = __ldg(mask + (long)i * ldmask + j) > 0.f ? v : 0.f;
                *c = v;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
colsum_kernel(const float *__restrict__ A, long lda, const float *__restrict__ mask, long ldmask, int Mi, int Nj,
              int rows_per_block, float *__restrict__ out) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= Nj) return;
    const int i0 = blockIdx.y * rows_per_block, i1 = min(Mi, i0 + rows_per_block);
    float s = 0.f;
    for (int i = i0; i < i1; ++i) {
        float v = __ldg(A + (long)i * lda + j);
        if (mask && !(__ldg(mask + (long)i * ldmask + j) > 0.f)) v = 0.f;
        s += v;
    }
    atomicAdd(out + j, s);
}

struct HannW { float w[16]; };

__global__ void __launch_bounds__(256)
hann_pe_kernel(const float *__restrict__ xyz, int m, HannW win, int multires, float *__restrict__ out, long ldo) {
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= m) return;
    const float x[3] = {__ldg(xyz + (long)q * 3), __ldg(xyz + (long)q * 3 + 1), __ldg(xyz + (long)q * 3 + 2)};
    float *o = out + (long)q * ldo;
    float freq = 1.0f;
    for (int j = 0; j < multires; ++j) {
        const float w = win.w[j];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = x[c] * freq;
            o[j * 6 + c] = w * sinf(a);
            o[j * 6 + 3 + c] = w * cosf(a);
        }
        freq *= 2.0f;
    }
}

template <bool AT, bool BT>
void launch_sgemm(bool vec, dim3 grid, cudaStream_t st, const float *A, long sAi, long sAr, const float *B, long sBr,
                  long sBj, float *C, long ldc, const float *bias, const float *mask, long ldmask, int Mi, int Nj,
                  int Kr, int flags, int kps) {
    if (vec)
        sgemm_kernel<AT, BT, true><<<grid, kThreads, 0, st>>>(A, sAi, sAr, B, sBr, sBj, C, ldc, bias, mask, ldmask, Mi, Nj, Kr, flags, kps);
    else
        sgemm_kernel<AT, BT, false><<<grid, kThreads, 0, st>>>(A, sAi, sAr, B, sBr, sBj, C, ldc, bias, mask, ldmask, Mi, Nj, Kr, flags, kps);
}

}  // namespace

extern "C" int occnerf_sgemm(const float *A, long sAi, long sAr, const float *B, long sBr, long sBj, float *C, long ldc,
                             const float *bias, const float *mask, long ldmask, int Mi, int Nj, int Kr, int flags,
                             int split_k, occnerf_stream_t stream) {
    OCC_CHECK_ARG(A && B && C, "sgemm: null pointer");
    OCC_CHECK_ARG(Mi >= 0 && Nj >= 0 && Kr >= 0 && split_k >= 1, "sgemm: bad sizes %d %d %d split %d", Mi, Nj, Kr, split_k);
    OCC_CHECK_ARG(!(flags & OCCNERF_GEMM_BIAS) || bias, "sgemm: BIAS flag without bias");
    OCC_CHECK_ARG(!(flags & OCCNERF_GEMM_RELUMASK) || mask, "sgemm: RELUMASK flag without mask");
    OCC_CHECK_ARG(sAi == 1 || sAr == 1, "sgemm: A must be contiguous along i or r");
    OCC_CHECK_ARG(sBj == 1 || sBr == 1, "sgemm: B must be contiguous along j or r");
    OCC_CHECK_ARG(split_k == 1 || !(flags & (OCCNERF_GEMM_RELU | OCCNERF_GEMM_RELUMASK)),
                  "sgemm: split_k > 1 cannot be combined with RELU / RELUMASK");
    if (Mi == 0 || Nj == 0) return OCCNERF_OK;
    const bool a_r_contig = sAr == 1;          // r contiguous -> transposing loader
    const bool b_r_contig = sBr == 1 && sBj != 1;
    // vector path: every float4 must be aligned and fully inside the logical extent
    auto ok16 = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
    int kps = (Kr + split_k - 1) / split_k;
    kps = ((kps + BK - 1) / BK) * BK;
    bool vec = ok16(A) && ok16(B) && kps % 4 == 0;
    if (a_r_contig) vec = vec && sAi % 4 == 0 && Kr % 4 == 0; else vec = vec && sAr % 4 == 0 && Mi % 4 == 0;
    if (b_r_contig) vec = vec && sBj % 4 == 0 && Kr % 4 == 0; else vec = vec && sBr % 4 == 0 && Nj % 4 == 0;
    const int splits = (Kr + kps - 1) / (kps > 0 ? kps : 1);
    dim3 grid(occ_div_up(Nj, BN), occ_div_up(Mi, BM), splits > 0 ? splits : 1);
    OCC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "sgemm: grid too large (%u,%u)", grid.y, grid.z);
    cudaStream_t st = (cudaStream_t)stream;
    if (a_r_contig && b_r_contig)
        launch_sgemm<true, true>(vec, grid, st, A, sAi, sAr, B, sBr, sBj, C, ldc, bias, mask, ldmask, Mi, Nj, Kr, flags, kps);
    else if (a_r_contig && !b_r_contig)
        launch_sgemm<true, false>(vec, grid, st, A, sAi, sAr, B, sBr, sBj, C, ldc, bias, mask, ldmask, Mi, Nj, Kr, flags, kps);
    else if (!a_r_contig && b_r_contig)
        launch_sgemm<false, true>(vec, grid, st, A, sAi, sAr, B, sBr, sBj, C, ldc, bias, mask, ldmask, Mi, Nj, Kr, flags, kps);
    else
        launch_sgemm<false, false>(vec, grid, st, A, sAi, sAr, B, sBr, sBj, C, ldc, bias, mask, ldmask, Mi, Nj, Kr, flags, kps);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_colsum(const float *A, long lda, const float *mask, long ldmask, int Mi, int Nj, float *out,
                              occnerf_stream_t stream) {
    OCC_CHECK_ARG(A && out && Mi >= 0 && Nj >= 0, "colsum: bad arguments");
    if (Mi == 0 || Nj == 0) return OCCNERF_OK;
    const int rows = 128;
    dim3 grid(occ_div_up(Nj, 256), occ_div_up(Mi, rows));
    colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, mask, ldmask, Mi, Nj, rows, out);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_hann_pe(const float *xyz, int m, const float *window_host, int multires, float *out, int ldo,
                               occnerf_stream_t stream) {
    OCC_CHECK_ARG(xyz && window_host && out, "hann_pe: null pointer");
    OCC_CHECK_ARG(multires >= 1 && multires <= 16 && ldo >= multires * 6, "hann_pe: multires=%d ldo=%d", multires, ldo);
    if (m <= 0) return OCCNERF_OK;
    HannW w;
    for (int j = 0; j < 16; ++j) w.w[j] = j < multires ? window_host[j] : 0.f;
    hann_pe_kernel<<<occ_div_up(m, 256), 256, 0, (cudaStream_t)stream>>>(xyz, m, w, multires, out, ldo);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
