// Per-frame prologue, the decoder of the motion-weight volume: five ConvTranspose3d(k = 4, s = 2, p = 1) with LeakyReLU(0.2) between
// them, batch 1, 1^3 -> 32^3 voxels, 63.3 M weights (core/nets/occnerf/mweight_vol_decoders/deconv_vol_decoder.py:25-33 +
// core/utils/network_util.py:12-50).  Forward, data gradient and weight gradient as tf32 tensor-core GEMMs that read the weights
// in the reference layout [Cin][Cout][4][4][4] -- no NCHW <-> NHWC passes (cuDNN spent 0.44 of its 0.93 ms per step on them) and
// no col2im / im2col buffers: the transposed convolution's index map lives in the operand loaders and in the epilogue.
//
// With q = (o, k) = o * 64 + tap (the contiguous axis of the weight tensor), V input voxels of edge D, output edge 2 D and
// out(v, k) = 2 v - 1 + k per axis (valid inside [0, 2 D)):
//   forward        C[q][v]  = sum_ci W[ci][q] * act(Yin[ci][v])            scatter-ADDED into Yout[o][out(v, k)]   (Yout starts as the bias)
//   weight gradient dW[ci][q] = sum_v  act(Yin[ci][v]) * dYout[o][out(v, k)]
//   data gradient   dYin[ci][v] = act'(Yin[ci][v]) * sum_q W[ci][q] * dYout[o][out(v, k)]
// act = LeakyReLU(0.2) of the previous layer's pre-activation (identity derivative / value handled by `slope`): activations are stored
// as pre-activations [C][D^3] (the reference's NCDHW with N = 1) and the non-linearity is applied while loading.
// One kernel template, mma.sync.m16n8k8 tf32 (fp32 operands rounded to nearest tf32 when staged, fp32 accumulation -- the arithmetic
// the library path used, torch.backends.cudnn.allow_tf32), 64 x 64 x 16 CTA tiles (128 x 8 where a layer has at most 8 voxels), operands
// staged through registers into shared memory with the next tile's loads in flight during the MMAs, optional split-K with fp32 reductions.
// The three small layers (1, 8, 64 input voxels) are weight-bandwidth-bound (254 MB of weights are streamed once per pass).
#include "common.cuh"

namespace {

constexpr int kTaps = 64;
constexpr int BK = 16;

enum Mode { FWD = 0, WGRAD = 1, DGRAD = 2 };

struct DeconvArgs {
    const float *W;        // [Cin][Cout][64]
    const float *Yin;      // [Cin][V]   pre-activations of the layer's input (after bias)
    const float *dYout;    // [Cout][8 V] gradient w.r.t. the layer's output pre-activations      (WGRAD, DGRAD)
    float *out;            // FWD: Yout [Cout][8 V] (+=)   WGRAD: dW [Cin][Cout][64]   DGRAD: dYin [Cin][V]
    int Cin, Cout, D;      // D = input edge (a power of two)
    int lg;                // log2(D)
    float slope;           // LeakyReLU slope applied to Yin (1 = the input is used as it is)
    int k_len;             // K range of one split (a multiple of BK)
    int accumulate;        // WGRAD / DGRAD: reduce into `out` (split-K, or the caller wants +=) instead of storing
    int exact;             // 0: one tf32 MMA per product;  1: 3 x tf32 (hi * hi + hi * lo + lo * hi of the operands' tf32 splits, fp32-grade)
};

__device__ __forceinline__ uint32_t tf32_rna(float x) { return __float_as_uint(x) + 0x1000u; }      // low mantissa bits are ignored by the MMA
__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

// dYout[o][out(v, k)] or 0 outside the output volume
__device__ __forceinline__ float gather_dy(const DeconvArgs &a, int q, int v) {
    const int o = q >> 6, k = q & 63;
    const int D = a.D, Do = 2 * D;
    const int x = v & (D - 1), y = (v >> a.lg) & (D - 1), z = v >> (2 * a.lg);
    const int xo = 2 * x - 1 + (k & 3), yo = 2 * y - 1 + ((k >> 2) & 3), zo = 2 * z - 1 + (k >> 4);
    if ((unsigned)xo >= (unsigned)Do || (unsigned)yo >= (unsigned)Do || (unsigned)zo >= (unsigned)Do) return 0.f;
    return __ldg(a.dYout + ((long)o * Do + zo) * Do * Do + yo * Do + xo);
}

// GEMM roles   FWD: M = q, N = v, K = ci     WGRAD: M = ci, N = q, K = v     DGRAD: M = ci, N = v, K = q
template <int MODE>
__device__ __forceinline__ void dims(const DeconvArgs &a, int &M, int &N, int &K) {
    const int V = a.D * a.D * a.D, Q = a.Cout * kTaps;
    if (MODE == FWD) { M = Q; N = V; K = a.Cin; }
    else if (MODE == WGRAD) { M = a.Cin; N = Q; K = V; }
    else { M = a.Cin; N = V; K = Q; }
}
template <int MODE>
__device__ __forceinline__ float load_a(const DeconvArgs &a, int m, int k, int M, int K) {
    if (m >= M || k >= K) return 0.f;
    const int V = a.D * a.D * a.D;
    if (MODE == FWD) return __ldg(a.W + (long)k * M + m);                                   // W[ci = k][q = m]
    if (MODE == WGRAD) return leaky(__ldg(a.Yin + (long)m * V + k), a.slope);               // act(Yin[ci = m][v = k])
    return __ldg(a.W + (long)m * K + k);                                                     // W[ci = m][q = k]
}
template <int MODE>
__device__ __forceinline__ float load_b(const DeconvArgs &a, int k, int n, int K, int N) {
    if (k >= K || n >= N) return 0.f;
    if (MODE == FWD) return leaky(__ldg(a.Yin + (long)k * N + n), a.slope);                 // act(Yin[ci = k][v = n])
    if (MODE == WGRAD) return gather_dy(a, n, k);                                            // dY[(o, k) = n][out(v = k)]
    return gather_dy(a, k, n);                                                               // dY[(o, k) = k][out(v = n)]
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>
__device__ __forceinline__ void epilogue_store(const DeconvArgs &a, int m, int n, int M, int N, float c) {
    if (m >= M || n >= N) return;
    if (MODE == FWD) {
        const int o = m >> 6, k = m & 63, D = a.D, Do = 2 * D;
        const int x = n & (D - 1), y = (n >> a.lg) & (D - 1), z = n >> (2 * a.lg);
        const int xo = 2 * x - 1 + (k & 3), yo = 2 * y - 1 + ((k >> 2) & 3), zo = 2 * z - 1 + (k >> 4);
        if ((unsigned)xo >= (unsigned)Do || (unsigned)yo >= (unsigned)Do || (unsigned)zo >= (unsigned)Do) return;
        atomicAdd(a.out + ((long)o * Do + zo) * Do * Do + yo * Do + xo, c);
    } else if (MODE == WGRAD) {
        float *p = a.out + (long)m * N + n;
        if (a.accumulate) atomicAdd(p, c); else *p = c;
    } else {
        const float y = __ldg(a.Yin + (long)m * N + n);
        const float g = c * (y > 0.f ? 1.f : a.slope);
        float *p = a.out + (long)m * N + n;
        if (a.accumulate) atomicAdd(p, g); else *p = g;
    }
}

// 4 warps.  BN = 64: 2 x 2 warps of 32 x 32.  BN = 8: 4 x 1 warps of 32 x 8 (BM = 128).
// SPLIT: operands are staged as tf32 (hi, lo) pairs and every product takes three MMAs (what torch.backends.cudnn.allow_tf32 = False
// asks of the library path: fp32-grade results).
template <int MODE, int BM, int BN, bool SPLIT>
__global__ void __launch_bounds__(128) deconv_gemm_kernel(const DeconvArgs a) {
    constexpr int LDA = BM + 8, LDB = BN + 8 - (BN == 8 ? 8 : 0);      // row strides = 8 (mod 32) words: conflict-free fragment loads
    static_assert(LDA % 32 == 8 && (LDB % 32 == 8), "fragment loads must be conflict free");
    constexpr int WN = BN == 8 ? 1 : 2, WM = 4 / WN;                   // warp grid
    constexpr int TM = BM / WM, TN = BN / WN;                          // warp tile
    constexpr int FM = TM / 16, FN = TN / 8;                           // m16n8 fragments per warp
    constexpr int NA = BM * BK / 128, NB = (BK * BN + 127) / 128;      // staged elements per thread
    constexpr int P = SPLIT ? 2 : 1;
    __shared__ uint32_t As[P][BK][LDA], Bs[P][BK][LDB];
    int M, N, K;
    dims<MODE>(a, M, N, K);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * a.k_len, k_end = min(K, k_begin + a.k_len);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = (warp / WN) * TM, wn = (warp % WN) * TN;
    // element (row, col) of the staged tile handled by this thread in round i: consecutive threads walk the operand's contiguous axis
    constexpr bool A_M_FAST = MODE == FWD;                             // A is contiguous along m (FWD) or along k
    float ra[NA], rb[NB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int e = tid + i * 128;
            const int mm = A_M_FAST ? e % BM : e / BK, kk = A_M_FAST ? e / BM : e % BK;
            ra[i] = (k0 + kk < k_end) ? load_a<MODE>(a, m0 + mm, k0 + kk, M, K) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int e = tid + i * 128;
            const int nn = e % BN, kk = e / BN;                        // B is walked along n in all three modes
            rb[i] = (e < BK * BN && k0 + kk < k_end) ? load_b<MODE>(a, k0 + kk, n0 + nn, K, N) : 0.f;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int e = tid + i * 128;
            const int mm = A_M_FAST ? e % BM : e / BK, kk = A_M_FAST ? e / BM : e % BK;
            const uint32_t hi = tf32_rna(ra[i]) & 0xffffe000u;
            As[0][kk][mm] = hi;
            if (SPLIT) As[1][kk][mm] = tf32_rna(ra[i] - __uint_as_float(hi));
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int e = tid + i * 128;
            if (e < BK * BN) {
                const uint32_t hi = tf32_rna(rb[i]) & 0xffffe000u;
                Bs[0][e / BN][e % BN] = hi;
                if (SPLIT) Bs[1][e / BN][e % BN] = tf32_rna(rb[i] - __uint_as_float(hi));
            }
        }
    };
    float acc[FM][FN][4];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;
    fetch(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        __syncthreads();                                               // the previous tile has been consumed
        stage();
        __syncthreads();
        if (k0 + BK < k_end) fetch(k0 + BK);                           // next tile's global loads fly during the MMAs
#pragma unroll
        for (int ks = 0; ks < BK; ks += 8) {
            uint32_t fa[P][FM][4], fb[P][FN][2];
#pragma unroll
            for (int p = 0; p < P; ++p) {
#pragma unroll
                for (int i = 0; i < FM; ++i) {
                    const int r = wm + i * 16 + g;
                    fa[p][i][0] = As[p][ks + t][r]; fa[p][i][1] = As[p][ks + t][r + 8];
                    fa[p][i][2] = As[p][ks + t + 4][r]; fa[p][i][3] = As[p][ks + t + 4][r + 8];
                }
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    const int c = wn + j * 8 + g;
                    fb[p][j][0] = Bs[p][ks + t][c]; fb[p][j][1] = Bs[p][ks + t + 4][c];
                }
            }
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    if (SPLIT) {                                       // small terms first
                        mma_tf32(acc[i][j], fa[1][i], fb[0][j]);
                        mma_tf32(acc[i][j], fa[0][i], fb[1][j]);
                    }
                    mma_tf32(acc[i][j], fa[0][i], fb[0][j]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int r = m0 + wm + i * 16 + g, c = n0 + wn + j * 8 + 2 * t;
            epilogue_store<MODE>(a, r, c, M, N, acc[i][j][0]);
            epilogue_store<MODE>(a, r, c + 1, M, N, acc[i][j][1]);
            epilogue_store<MODE>(a, r + 8, c, M, N, acc[i][j][2]);
            epilogue_store<MODE>(a, r + 8, c + 1, M, N, acc[i][j][3]);
        }
}

// out[c][v] = bias[c]  (the forward kernel scatter-adds onto it);  V voxels per channel
__global__ void __launch_bounds__(256) bias_fill_kernel(const float *__restrict__ bias, int C, long V, float *__restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long)C * V) out[i] = __ldg(bias + i / V);
}

// g_bias[c] = sum_v dY[c][v]: one warp per channel
__global__ void __launch_bounds__(256) bias_grad_kernel(const float *__restrict__ dY, int C, long V, float *__restrict__ g_bias) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    float s = 0.f;
    for (long v = lane; v < V; v += 32) s += __ldg(dY + (long)c * V + v);
    s = warp_sum(s);
    if (lane == 0) g_bias[c] = s;
}

// The 256 -> 1024 linear layer in front of the stack (network_util.py:25-28), batch 1.
//   forward : y[n] = b[n] + sum_k w[n][k] e[k]                      (pre-activation; the first deconvolution applies the LeakyReLU)
//   backward: given g[n] = dL/dy[n]:  dw[n][k] = g[n] e[k],  db[n] = g[n],  de[k] = sum_n w[n][k] g[n]
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float *__restrict__ w, const float *__restrict__ b, const float *__restrict__ e,
                                                         int n_out, int n_in, float *__restrict__ y) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= n_out) return;
    float s = 0.f;
    for (int k = lane; k < n_in; k += 32) s = fmaf(__ldg(w + (long)n * n_in + k), __ldg(e + k), s);
    s = warp_sum(s);
    if (lane == 0) y[n] = s + __ldg(b + n);
}
__global__ void __launch_bounds__(256) linear_bwd_kernel(const float *__restrict__ w, const float *__restrict__ e, const float *__restrict__ g,
                                                         int n_out, int n_in, float *__restrict__ dw, float *__restrict__ db, float *__restrict__ de) {
    // block b: rows [8 b, 8 b + 8) of dw and db (warp per row), and their share of de (reduced into the caller-zeroed de)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + warp;
    __shared__ float s_de[8][256];
    const float gn = n < n_out ? __ldg(g + n) : 0.f;
    for (int k0 = 0; k0 < n_in; k0 += 256) {
        for (int k = k0 + lane; k < min(k0 + 256, n_in); k += 32) {
            if (n < n_out) dw[(long)n * n_in + k] = gn * __ldg(e + k);
            s_de[warp][k - k0] = n < n_out ? __ldg(w + (long)n * n_in + k) * gn : 0.f;
        }
        __syncthreads();
        const int k = k0 + threadIdx.x;
        if (k < n_in) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) s += s_de[r][threadIdx.x];
            atomicAdd(de + k, s);
        }
        __syncthreads();
    }
    if (lane == 0 && n < n_out) db[n] = gn;
}

template <int MODE>
int launch_gemm(const DeconvArgs &a0, int splits, cudaStream_t st) {
    DeconvArgs a = a0;
    a.lg = 0;
    while ((1 << a.lg) < a.D) ++a.lg;
    const int V = a.D * a.D * a.D, Q = a.Cout * kTaps;
    const int M = MODE == FWD ? Q : a.Cin, N = MODE == WGRAD ? Q : V, K = MODE == FWD ? a.Cin : (MODE == WGRAD ? V : Q);
    splits = splits < 1 ? 1 : splits;
    int k_len = (K + splits - 1) / splits;
    k_len = (k_len + BK - 1) / BK * BK;
    splits = (K + k_len - 1) / k_len;
    a.k_len = k_len;
    if (MODE != FWD && splits > 1) a.accumulate = 1;
    if (N <= 8) {
        dim3 grid(1, occ_div_up(M, 128), splits);
        if (a.exact) deconv_gemm_kernel<MODE, 128, 8, true><<<grid, 128, 0, st>>>(a);
        else deconv_gemm_kernel<MODE, 128, 8, false><<<grid, 128, 0, st>>>(a);
    } else {
        dim3 grid(occ_div_up(N, 64), occ_div_up(M, 64), splits);
        if (a.exact) deconv_gemm_kernel<MODE, 64, 64, true><<<grid, 128, 0, st>>>(a);
        else deconv_gemm_kernel<MODE, 64, 64, false><<<grid, 128, 0, st>>>(a);
    }
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

bool bad_layer(int Cin, int Cout, int D) { return Cin < 1 || Cout < 1 || D < 1 || D > 64 || (D & (D - 1)) != 0 || (long)Cout * kTaps > (1 << 24); }

}  // namespace

// Yout [Cout][(2 D)^3] = bias + ConvTranspose3d(act(Yin)), act = LeakyReLU(slope) (slope = 1: none).  W [Cin][Cout][4][4][4] as the
// reference stores it, Yin [Cin][D^3].  `splits` > 1 splits the contraction over Cin across CTAs (the output is accumulated either way).
extern "C" int occnerf_deconv3d_forward(const float *W, const float *bias, const float *Yin, int Cin, int Cout, int D, float slope, int splits,
                                        int exact, float *Yout, occnerf_stream_t stream) {
    OCC_CHECK_ARG(W && bias && Yin && Yout, "deconv3d_forward: null pointer");
    OCC_CHECK_ARG(!bad_layer(Cin, Cout, D), "deconv3d_forward: Cin=%d Cout=%d D=%d", Cin, Cout, D);
    const long Vo = 8L * D * D * D;
    bias_fill_kernel<<<occ_div_up(Cout * Vo, 256), 256, 0, (cudaStream_t)stream>>>(bias, Cout, Vo, Yout);
    OCC_LAUNCH_CHECK();
    DeconvArgs a = {};
    a.W = W; a.Yin = Yin; a.out = Yout; a.Cin = Cin; a.Cout = Cout; a.D = D; a.slope = slope; a.exact = exact ? 1 : 0;
    return launch_gemm<FWD>(a, splits, (cudaStream_t)stream);
}

// Given dYout [Cout][(2 D)^3] (gradient w.r.t. the layer's output pre-activation):
//   dW [Cin][Cout][64] (stored; or reduced into when `accumulate` / splits > 1 -- then the caller zeroes it), dbias [Cout],
//   dYin [Cin][D^3] = act'(Yin) * (data gradient), or NULL to skip it (first layer of a stack whose input needs no gradient).
// With split-K the reduced outputs must be zeroed by the caller: dW if w_splits > 1, dYin if d_splits > 1.
extern "C" int occnerf_deconv3d_backward(const float *W, const float *Yin, const float *dYout, int Cin, int Cout, int D, float slope,
                                         int w_splits, int d_splits, int accumulate_dw, int exact, float *dW, float *dbias, float *dYin,
                                         occnerf_stream_t stream) {
    OCC_CHECK_ARG(W && Yin && dYout && dW && dbias, "deconv3d_backward: null pointer");
    OCC_CHECK_ARG(!bad_layer(Cin, Cout, D), "deconv3d_backward: Cin=%d Cout=%d D=%d", Cin, Cout, D);
    const long Vo = 8L * D * D * D;
    bias_grad_kernel<<<occ_div_up(Cout, 8), 256, 0, (cudaStream_t)stream>>>(dYout, Cout, Vo, dbias);
    OCC_LAUNCH_CHECK();
    DeconvArgs a = {};
    a.W = W; a.Yin = Yin; a.dYout = dYout; a.Cin = Cin; a.Cout = Cout; a.D = D; a.slope = slope; a.exact = exact ? 1 : 0;
    a.out = dW; a.accumulate = accumulate_dw ? 1 : 0;
    int rc = launch_gemm<WGRAD>(a, w_splits, (cudaStream_t)stream);
    if (rc != OCCNERF_OK || !dYin) return rc;
    a.out = dYin; a.accumulate = 0;
    return launch_gemm<DGRAD>(a, d_splits, (cudaStream_t)stream);
}

extern "C" int occnerf_decoder_linear_forward(const float *w, const float *b, const float *e, int n_out, int n_in, float *y, occnerf_stream_t stream) {
    OCC_CHECK_ARG(w && b && e && y && n_out > 0 && n_in > 0, "decoder_linear_forward: bad arguments");
    linear_fwd_kernel<<<occ_div_up(n_out, 8), 256, 0, (cudaStream_t)stream>>>(w, b, e, n_out, n_in, y);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_decoder_linear_backward(const float *w, const float *e, const float *g, int n_out, int n_in, float *dw, float *db, float *de,
                                               occnerf_stream_t stream) {
    OCC_CHECK_ARG(w && e && g && dw && db && de && n_out > 0 && n_in > 0, "decoder_linear_backward: bad arguments");
    OCC_CUDA(cudaMemsetAsync(de, 0, sizeof(float) * n_in, (cudaStream_t)stream));
    linear_bwd_kernel<<<occ_div_up(n_out, 8), 256, 0, (cudaStream_t)stream>>>(w, e, g, n_out, n_in, dw, db, de);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
