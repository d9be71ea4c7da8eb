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
    int T;                 // taps per output channel that can touch the output at all: 64, or 8 when D = 1 (a 1^3 input reaches only the
                           // taps {1, 2}^3: the other 56 of the first layer are never read, and their gradient is never written -- it is zero)
};

__device__ __forceinline__ uint32_t tf32_rna(float x) { return __float_as_uint(x) + 0x1000u; }      // low mantissa bits are ignored by the MMA
__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

// effective column index qe (o * T + j) -> (o, tap k):  T = 64: the identity;  T = 8: j enumerates the taps {1, 2}^3
__device__ __forceinline__ void split_q(const DeconvArgs &a, int qe, int &o, int &k) {
    if (a.T == 64) { o = qe >> 6; k = qe & 63; return; }
    o = qe >> 3;
    const int j = qe & 7;
    k = (1 + (j >> 2)) * 16 + (1 + ((j >> 1) & 1)) * 4 + 1 + (j & 1);
}
__device__ __forceinline__ long q_index(const DeconvArgs &a, int qe) {      // position of (o, k) along the [Cout][64] axis of W / dW
    int o, k;
    split_q(a, qe, o, k);
    return (long)o * kTaps + k;
}

// dYout[o][out(v, k)] or 0 outside the output volume
__device__ __forceinline__ float gather_dy(const DeconvArgs &a, int q, int v) {
    int o, k;
    split_q(a, q, o, k);
    const int D = a.D, Do = 2 * D;
    const int x = v & (D - 1), y = (v >> a.lg) & (D - 1), z = v >> (2 * a.lg);
    const int xo = 2 * x - 1 + (k & 3), yo = 2 * y - 1 + ((k >> 2) & 3), zo = 2 * z - 1 + (k >> 4);
    if ((unsigned)xo >= (unsigned)Do || (unsigned)yo >= (unsigned)Do || (unsigned)zo >= (unsigned)Do) return 0.f;
    return __ldg(a.dYout + ((long)o * Do + zo) * Do * Do + yo * Do + xo);
}

// GEMM roles   FWD: M = q, N = v, K = ci     WGRAD: M = ci, N = q, K = v     DGRAD: M = ci, N = v, K = q
template <int MODE>
__device__ __forceinline__ void dims(const DeconvArgs &a, int &M, int &N, int &K) {
    const int V = a.D * a.D * a.D, Q = a.Cout * a.T;
    if (MODE == FWD) { M = Q; N = V; K = a.Cin; }
    else if (MODE == WGRAD) { M = a.Cin; N = Q; K = V; }
    else { M = a.Cin; N = V; K = Q; }
}
template <int MODE>
__device__ __forceinline__ float load_a(const DeconvArgs &a, int m, int k, int M, int K) {
    if (m >= M || k >= K) return 0.f;
    const int V = a.D * a.D * a.D;
    const long Qs = (long)a.Cout * kTaps;                                                    // row stride of W
    if (MODE == FWD) return __ldg(a.W + (long)k * Qs + q_index(a, m));                       // W[ci = k][q = m]
    if (MODE == WGRAD) return leaky(__ldg(a.Yin + (long)m * V + k), a.slope);               // act(Yin[ci = m][v = k])
    return __ldg(a.W + (long)m * Qs + q_index(a, k));                                        // W[ci = m][q = k]
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
        int o, k;
        split_q(a, m, o, k);
        const int D = a.D, Do = 2 * D;
        const int x = n & (D - 1), y = (n >> a.lg) & (D - 1), z = n >> (2 * a.lg);
        const int xo = 2 * x - 1 + (k & 3), yo = 2 * y - 1 + ((k >> 2) & 3), zo = 2 * z - 1 + (k >> 4);
        if ((unsigned)xo >= (unsigned)Do || (unsigned)yo >= (unsigned)Do || (unsigned)zo >= (unsigned)Do) return;
        atomicAdd(a.out + ((long)o * Do + zo) * Do * Do + yo * Do + xo, c);
    } else if (MODE == WGRAD) {
        float *p = a.out + (long)m * a.Cout * kTaps + q_index(a, n);
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
// Staging: the dense operand A is fetched with 16-byte loads along its contiguous axis whenever the tile allows it -- along m in the
// forward pass (shared-memory image [k][m]), along k in the two gradient passes (image [m][k]; both images have conflict-free fragment
// loads) --; the gathered operand (dYout through the transposed convolution's index map) is fetched element by element with the
// per-thread part of the index arithmetic (the thread's column n is the same for every K step) done once.
template <int MODE, int BM, int BN, bool SPLIT, int NT>
__global__ void __launch_bounds__(NT) deconv_gemm_kernel(const DeconvArgs a) {
    constexpr bool A_T = MODE != FWD;                                   // A image [m][k] (k contiguous) instead of [k][m]
    constexpr int LDA = A_T ? BK + 4 : BM + 8, LDB = BN + 8 - (BN == 8 ? 8 : 0);
    static_assert(A_T ? (LDA % 8 == 4) : (LDA % 32 == 8), "A fragment loads must be conflict free");
    static_assert(LDB % 32 == 8, "B fragment loads must be conflict free");
    constexpr int WN = BN == 8 ? 1 : 2, WM = (NT / 32) / WN;           // warp grid
    constexpr int TM = BM / WM, TN = BN / WN;                          // warp tile
    constexpr int FM = TM / 16, FN = TN / 8;                           // m16n8 fragments per warp
    constexpr int NA = BM * BK / NT, NB = (BK * BN + NT - 1) / NT;     // staged elements per thread
    static_assert(NA % 4 == 0 && NT % BN == 0, "staging maps");
    constexpr int P = SPLIT ? 2 : 1;
    constexpr int A_ROWS = A_T ? BM : BK;
    __shared__ __align__(16) uint32_t As[P][A_ROWS][LDA];
    __shared__ uint32_t Bs[P][BK][LDB];
    int M, N, K;
    dims<MODE>(a, M, N, K);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * a.k_len, k_end = min(K, k_begin + a.k_len);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = (warp / WN) * TM, wn = (warp % WN) * TN;
    const int V = a.D * a.D * a.D;
    const long Qs = (long)a.Cout * kTaps;
    // ---- A: 16-byte path when the whole tile is inside the matrix, rows are 16-byte aligned and the tap axis is not remapped
    bool vec_a;
    if (MODE == FWD) vec_a = a.T == kTaps && m0 + BM <= M;
    else if (MODE == WGRAD) vec_a = (V % 4 == 0) && m0 + BM <= M && (a.k_len % 4 == 0);
    else vec_a = a.T == kTaps && m0 + BM <= M;
    float ra[NA], rb[NB];
    // ---- B: the thread's column nn is fixed (128 % BN == 0); its share of the gather arithmetic
    const int nn_b = tid % BN, kk_b0 = tid / BN;
    constexpr int KK_STEP = NT / BN;
    const int n_b = n0 + nn_b;
    int gb0 = 0, gb1 = 0, gb2 = 0, gb3 = 0;                            // WGRAD: (o * Do^3, kz-1, ky-1, kx-1);  DGRAD: (2z-1, 2y-1, 2x-1)
    const int D = a.D, Do = 2 * a.D;
    if (MODE == WGRAD && n_b < N) {
        int o, k;
        split_q(a, n_b, o, k);
        gb0 = o; gb1 = (k >> 4) - 1; gb2 = ((k >> 2) & 3) - 1; gb3 = (k & 3) - 1;
    } else if (MODE == DGRAD && n_b < N) {
        gb1 = 2 * (n_b >> (2 * a.lg)) - 1; gb2 = 2 * ((n_b >> a.lg) & (D - 1)) - 1; gb3 = 2 * (n_b & (D - 1)) - 1;
    }
    auto fetch = [&](int k0) {
        if (vec_a) {
#pragma unroll
            for (int i = 0; i < NA / 4; ++i) {
                const int e = tid + i * NT;
                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MODE == FWD) {                                     // 4 consecutive m of one k:  W[ci = k][q = m]
                    const int m4 = e % (BM / 4), kk = e / (BM / 4);
                    if (k0 + kk < k_end) v4 = __ldg(reinterpret_cast<const float4 *>(a.W + (long)(k0 + kk) * Qs + m0 + 4 * m4));
                } else {                                               // 4 consecutive k of one m
                    const int k4 = e % (BK / 4), mm = e / (BK / 4);
                    const int kq = k0 + 4 * k4;
                    if (kq < k_end) {                                  // (k_end and K are multiples of 4 on this path)
                        if (MODE == WGRAD) {
                            v4 = __ldg(reinterpret_cast<const float4 *>(a.Yin + (long)(m0 + mm) * V + kq));
                            v4.x = leaky(v4.x, a.slope); v4.y = leaky(v4.y, a.slope); v4.z = leaky(v4.z, a.slope); v4.w = leaky(v4.w, a.slope);
                        } else {
                            v4 = __ldg(reinterpret_cast<const float4 *>(a.W + (long)(m0 + mm) * Qs + kq));
                        }
                    }
                }
                ra[4 * i] = v4.x; ra[4 * i + 1] = v4.y; ra[4 * i + 2] = v4.z; ra[4 * i + 3] = v4.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                const int e = tid + i * NT;
                const int mm = A_T ? e / BK : e % BM, kk = A_T ? e % BK : e / BM;
                ra[i] = (k0 + kk < k_end) ? load_a<MODE>(a, m0 + mm, k0 + kk, M, K) : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int kk = kk_b0 + i * KK_STEP, kg = k0 + kk;
            float v = 0.f;
            if (kk < BK && kg < k_end && n_b < N) {
                if (MODE == FWD) {
                    v = leaky(__ldg(a.Yin + (long)kg * N + n_b), a.slope);
                } else if (MODE == WGRAD) {                            // kg = input voxel, the thread's column = (o, tap)
                    const int zo = 2 * (kg >> (2 * a.lg)) + gb1, yo = 2 * ((kg >> a.lg) & (D - 1)) + gb2, xo = 2 * (kg & (D - 1)) + gb3;
                    if ((unsigned)xo < (unsigned)Do && (unsigned)yo < (unsigned)Do && (unsigned)zo < (unsigned)Do)
                        v = __ldg(a.dYout + (((long)gb0 * Do + zo) * Do + yo) * Do + xo);
                } else {                                               // kg = (o, tap), the thread's column = input voxel
                    int o, k;
                    split_q(a, kg, o, k);
                    const int zo = gb1 + (k >> 4), yo = gb2 + ((k >> 2) & 3), xo = gb3 + (k & 3);
                    if ((unsigned)xo < (unsigned)Do && (unsigned)yo < (unsigned)Do && (unsigned)zo < (unsigned)Do)
                        v = __ldg(a.dYout + (((long)o * Do + zo) * Do + yo) * Do + xo);
                }
            }
            rb[i] = v;
        }
    };
    auto put_a = [&](int row, int col, float x) {
        const uint32_t hi = tf32_rna(x) & 0xffffe000u;
        As[0][row][col] = hi;
        if (SPLIT) As[1][row][col] = tf32_rna(x - __uint_as_float(hi));
    };
    auto stage = [&]() {
        if (vec_a) {
#pragma unroll
            for (int i = 0; i < NA / 4; ++i) {
                const int e = tid + i * NT;
                int row, col;
                if (MODE == FWD) { row = e / (BM / 4); col = 4 * (e % (BM / 4)); }          // [k][m .. m+3]
                else { row = e / (BK / 4); col = 4 * (e % (BK / 4)); }                      // [m][k .. k+3]
                uint32_t hi[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) hi[j] = tf32_rna(ra[4 * i + j]) & 0xffffe000u;
                *reinterpret_cast<uint4 *>(&As[0][row][col]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                if (SPLIT)
                    *reinterpret_cast<uint4 *>(&As[1][row][col]) =
                        make_uint4(tf32_rna(ra[4 * i] - __uint_as_float(hi[0])), tf32_rna(ra[4 * i + 1] - __uint_as_float(hi[1])),
                                   tf32_rna(ra[4 * i + 2] - __uint_as_float(hi[2])), tf32_rna(ra[4 * i + 3] - __uint_as_float(hi[3])));
            }
        } else {
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                const int e = tid + i * NT;
                const int mm = A_T ? e / BK : e % BM, kk = A_T ? e % BK : e / BM;
                if (A_T) put_a(mm, kk, ra[i]); else put_a(kk, mm, ra[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int kk = kk_b0 + i * KK_STEP;
            if (kk < BK) {
                const uint32_t hi = tf32_rna(rb[i]) & 0xffffe000u;
                Bs[0][kk][nn_b] = hi;
                if (SPLIT) Bs[1][kk][nn_b] = tf32_rna(rb[i] - __uint_as_float(hi));
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
                    if (A_T) {
                        fa[p][i][0] = As[p][r][ks + t]; fa[p][i][1] = As[p][r + 8][ks + t];
                        fa[p][i][2] = As[p][r][ks + t + 4]; fa[p][i][3] = As[p][r + 8][ks + t + 4];
                    } else {
                        fa[p][i][0] = As[p][ks + t][r]; fa[p][i][1] = As[p][ks + t][r + 8];
                        fa[p][i][2] = As[p][ks + t + 4][r]; fa[p][i][3] = As[p][ks + t + 4][r + 8];
                    }
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

// g_bias[c] += sum_v dY[c][v]: block (c, chunk of 2048 voxels), the caller zeroes g_bias (the one-warp-per-channel form took 60 us for the
// 25 x 32768 last layer: four blocks on the whole GPU)
__global__ void __launch_bounds__(256) bias_grad_kernel(const float *__restrict__ dY, int C, long V, float *__restrict__ g_bias) {
    const int c = blockIdx.y;
    const long v0 = (long)blockIdx.x * 2048;
    float s = 0.f;
    for (long v = v0 + threadIdx.x; v < min(V, v0 + 2048); v += 256) s += __ldg(dY + (long)c * V + v);
    s = warp_sum(s);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w];
        atomicAdd(g_bias + c, t);
    }
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
    a.T = a.D == 1 ? 8 : kTaps;
    const int V = a.D * a.D * a.D, Q = a.Cout * a.T;
    const int M = MODE == FWD ? Q : a.Cin, N = MODE == WGRAD ? Q : V, K = MODE == FWD ? a.Cin : (MODE == WGRAD ? V : Q);
    splits = splits < 1 ? 1 : splits;
    int k_len = (K + splits - 1) / splits;
    k_len = (k_len + BK - 1) / BK * BK;
    splits = (K + k_len - 1) / k_len;
    a.k_len = k_len;
    if (MODE != FWD && splits > 1) a.accumulate = 1;
    if (N <= 8) {
        dim3 grid(1, occ_div_up(M, 128), splits);
        if (a.exact) deconv_gemm_kernel<MODE, 128, 8, true, 128><<<grid, 128, 0, st>>>(a);
        else deconv_gemm_kernel<MODE, 128, 8, false, 128><<<grid, 128, 0, st>>>(a);
    } else if (MODE != FWD && M >= 256 && !a.exact) {
        // gradient passes of the wide layers: 256-row tiles (8 warps), so that a gathered dYout element feeds 256 rows instead of 64
        // (tf32 mode only: the (hi, lo) images of a 256-row tile would not fit the 48 KB of static shared memory)
        dim3 grid(occ_div_up(N, 64), occ_div_up(M, 256), splits);
        deconv_gemm_kernel<MODE, 256, 64, false, 256><<<grid, 256, 0, st>>>(a);
    } else {
        dim3 grid(occ_div_up(N, 64), occ_div_up(M, 64), splits);
        if (a.exact) deconv_gemm_kernel<MODE, 64, 64, true, 128><<<grid, 128, 0, st>>>(a);
        else deconv_gemm_kernel<MODE, 64, 64, false, 128><<<grid, 128, 0, st>>>(a);
    }
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

// side stream of the backward entry point (one per device, created on first use, never destroyed)
struct Side { cudaStream_t st; cudaEvent_t fork, join; };
int g_overlap = 1;
Side *side_stream() {
    static Side pool[64];
    static bool made[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { occnerf_set_error("deconv3d_backward: cudaGetDevice failed"); return nullptr; }
    if (!made[dev]) {
        Side s;
        if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
            occnerf_set_error("deconv3d_backward: side stream: %s", cudaGetErrorString(cudaGetLastError()));
            return nullptr;
        }
        pool[dev] = s;
        made[dev] = true;
    }
    return &pool[dev];
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
    cudaStream_t main_st = (cudaStream_t)stream, wst = main_st;
    // The weight-gradient and the data-gradient GEMM of a layer share their inputs and neither fills the GPU (~300 CTAs of 128 - 256
    // threads, latency-bound): with a data gradient to compute, bias + weight gradient go to a side stream forked from `stream` and
    // joined again before the call returns (event dependencies: capturable in a CUDA graph, where they become parallel branches).
    Side *side = nullptr;
    if (dYin && g_overlap) {
        side = side_stream();
        if (!side) return OCCNERF_ECUDA;
        OCC_CUDA(cudaEventRecord(side->fork, main_st));
        OCC_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
        wst = side->st;
    }
    OCC_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Cout, wst));
    bias_grad_kernel<<<dim3(occ_div_up(Vo, 2048), Cout), 256, 0, wst>>>(dYout, Cout, Vo, dbias);
    OCC_LAUNCH_CHECK();
    DeconvArgs a = {};
    a.W = W; a.Yin = Yin; a.dYout = dYout; a.Cin = Cin; a.Cout = Cout; a.D = D; a.slope = slope; a.exact = exact ? 1 : 0;
    a.out = dW; a.accumulate = accumulate_dw ? 1 : 0;
    int rc = launch_gemm<WGRAD>(a, w_splits, wst);
    if (rc == OCCNERF_OK && dYin) {
        a.out = dYin; a.accumulate = 0;
        rc = launch_gemm<DGRAD>(a, d_splits, main_st);
    }
    if (side) {                                                       // a capturing stream must not be left forked
        OCC_CUDA(cudaEventRecord(side->join, side->st));
        OCC_CUDA(cudaStreamWaitEvent(main_st, side->join, 0));
    }
    return rc;
}

// 0: bias / weight gradient and data gradient one after the other on the caller's stream (for A/B timing); default 1
extern "C" int occnerf_deconv_set_overlap(int on) {
    g_overlap = on ? 1 : 0;
    return OCCNERF_OK;
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
