// Visibility-weighted aggregation of per-vertex features over the 4x10 multi-scale neighbours.
//
// Replaces CanonicalMLP.simple_agg and the (m,40,35) gather feeding it
// (core/nets/occnerf/canonical_mlps/occnerf_mlp.py:110-126,175-178): the reference materialises
// feats[knn_idxs] = 5.6 KB per sample in HBM; here one warp owns one sample, the 40 attention weights
// live in registers/shared memory and each neighbour row (36 floats, 144 B, L2/L1 resident: the whole
// table is 6890 x 144 B = 0.97 MB) is read with 9 x LDG.128 by one of three 9-lane groups.
// Output goes straight into the MLP input row: X[0..34] = agg, X[35] = unbiased variance of the
// normalised attention.
//
// Backward: att is detached upstream (occnerf_mlp.py:123), so only feats receives a gradient:
// g_feats[idx_n] += att_n * gX[0..34] with red.global.add.v4.f32.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxNN = 64;
constexpr int kRowF4 = 9;   // 36 floats per feature row

// attention weights of one sample -> shared (idx, w); returns the variance (valid in all lanes)
__device__ __forceinline__ float attention(const int32_t *__restrict__ idx_row, const float *__restrict__ counter, int nn,
                                           int lane, int *s_idx, float *s_w) {
    const int n0 = lane, n1 = lane + 32;
    const bool v0 = n0 < nn, v1 = n1 < nn;
    const int i0 = v0 ? __ldg(idx_row + n0) : 0, i1 = v1 ? __ldg(idx_row + n1) : 0;
    float a0 = v0 ? __ldg(counter + i0) : 0.f, a1 = v1 ? __ldg(counter + i1) : 0.f;
    const float mn = warp_min(fminf(v0 ? a0 : INFINITY, v1 ? a1 : INFINITY));
    a0 = __fadd_rn(a0, __fsub_rn(1.0f, mn));
    a1 = __fadd_rn(a1, __fsub_rn(1.0f, mn));
    const float mx = warp_max(fmaxf(v0 ? a0 : -INFINITY, v1 ? a1 : -INFINITY));
    a0 = __fdiv_rn(a0, mx);
    a1 = __fdiv_rn(a1, mx);
    const float mean = warp_sum((v0 ? a0 : 0.f) + (v1 ? a1 : 0.f)) / (float)nn;
    const float d0 = v0 ? a0 - mean : 0.f, d1 = v1 ? a1 - mean : 0.f;
    const float var = warp_sum(d0 * d0 + d1 * d1) / (float)(nn - 1);
    const float amax = warp_max(fmaxf(v0 ? a0 : -INFINITY, v1 ? a1 : -INFINITY));
    const float e0 = v0 ? expf(a0 - amax) : 0.f, e1 = v1 ? expf(a1 - amax) : 0.f;
    const float den = warp_sum(e0 + e1);
    if (v0) { s_idx[n0] = i0; s_w[n0] = e0 / den; }
    if (v1) { s_idx[n1] = i1; s_w[n1] = e1 / den; }
    __syncwarp();
    return var;
}

__global__ void __launch_bounds__(kWarps * 32)
aggregate_fwd_kernel(const int32_t *__restrict__ knn_idx, const float *__restrict__ counter,
                     const float4 *__restrict__ feats, int m, int nn, float *__restrict__ X, long ldx) {
    __shared__ int s_idx[kWarps][kMaxNN];
    __shared__ float s_w[kWarps][kMaxNN];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long q = (long)blockIdx.x * kWarps + wib;
    if (q >= m) return;
    const float var = attention(knn_idx + q * nn, counter, nn, lane, s_idx[wib], s_w[wib]);
    const int grp = lane / kRowF4, col = lane - grp * kRowF4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grp < 3) {
        for (int n = grp; n < nn; n += 3) {
            const float w = s_w[wib][n];
            const float4 f = __ldg(feats + (size_t)s_idx[wib][n] * kRowF4 + col);
            acc.x = fmaf(w, f.x, acc.x); acc.y = fmaf(w, f.y, acc.y);
            acc.z = fmaf(w, f.z, acc.z); acc.w = fmaf(w, f.w, acc.w);
        }
    }
    // lanes 0..8 gather the partial sums of the other two 9-lane groups (both read before either is added)
    const float4 p1 = make_float4(__shfl_down_sync(OCC_FULL, acc.x, kRowF4), __shfl_down_sync(OCC_FULL, acc.y, kRowF4),
                                  __shfl_down_sync(OCC_FULL, acc.z, kRowF4), __shfl_down_sync(OCC_FULL, acc.w, kRowF4));
    const float4 p2 = make_float4(__shfl_down_sync(OCC_FULL, acc.x, 2 * kRowF4), __shfl_down_sync(OCC_FULL, acc.y, 2 * kRowF4),
                                  __shfl_down_sync(OCC_FULL, acc.z, 2 * kRowF4), __shfl_down_sync(OCC_FULL, acc.w, 2 * kRowF4));
    acc.x += p1.x + p2.x; acc.y += p1.y + p2.y; acc.z += p1.z + p2.z; acc.w += p1.w + p2.w;
    if (lane < kRowF4) {
        if (lane == kRowF4 - 1) acc.w = var;   // column 35 carries the variance
        *reinterpret_cast<float4 *>(X + q * ldx + lane * 4) = acc;
    }
}

__global__ void __launch_bounds__(kWarps * 32)
aggregate_bwd_kernel(const int32_t *__restrict__ knn_idx, const float *__restrict__ counter,
                     const float *__restrict__ gX, long ldg, int m, int nn, float *__restrict__ g_feats) {
    __shared__ int s_idx[kWarps][kMaxNN];
    __shared__ float s_w[kWarps][kMaxNN];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long q = (long)blockIdx.x * kWarps + wib;
    if (q >= m) return;
    attention(knn_idx + q * nn, counter, nn, lane, s_idx[wib], s_w[wib]);
    const int grp = lane / kRowF4, col = lane - grp * kRowF4;
    if (grp >= 3) return;
    float4 g = __ldg(reinterpret_cast<const float4 *>(gX + q * ldg) + col);
    if (col == kRowF4 - 1) g.w = 0.f;   // column 35 is the variance slot, not a feature
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) return;
    for (int n = grp; n < nn; n += 3) {
        const float w = s_w[wib][n];
        red_add_v4(g_feats + ((size_t)s_idx[wib][n] * kRowF4 + col) * 4, w * g.x, w * g.y, w * g.z, w * g.w);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward, v2: pre-reduction in shared memory.  31 M (sample, neighbour) contributions per step land on only 6890
// feature rows, so v1's one-RED-per-contribution is bound by the L2 atomic units.  Here a warp owns 32 consecutive RAYS
// at one sample index (a few cm apart: ~64 distinct vertices among their 1280 neighbour slots) and accumulates
// into a private 128-entry open-addressing table in shared memory, without atomics on the values: lanes are split into
// three 9-lane groups (one float4 column chunk per lane) that take three neighbour slots OF THE SAME LEVEL of one sample
// per iteration -- ids within a level are distinct, so the three groups never touch the same row.  Keys are claimed
// with a shared-memory CAS.  The table is flushed to global memory with red.v4 when it fills up and at the end:
// ~20-60x fewer L2 reductions.
constexpr int kCap = 128;
constexpr int kW2 = 6;                   // warps per CTA
constexpr int kIdxStride = 41;           // 40 + 1: conflict-free per-lane rows
constexpr int kEmpty = -1;

struct __align__(16) WarpScratch {
    float vals[kCap][36];
    float grow[32][36];
    int keys[kCap];
    int sidx[32][kIdxStride];
    float sw[32][kIdxStride];
};

__device__ __forceinline__ void table_flush(WarpScratch &ws, int lane, float *__restrict__ g_feats) {
    const int grp = lane / kRowF4, col = lane - grp * kRowF4;
    __syncwarp();
    for (int e0 = 0; e0 < kCap; e0 += 3) {
        const int e = e0 + grp;
        if (grp < 3 && e < kCap) {
            const int key = ws.keys[e];
            if (key != kEmpty) {
                float4 *vp = reinterpret_cast<float4 *>(&ws.vals[e][col * 4]);
                const float4 v = *vp;
                red_add_v4(g_feats + ((size_t)key * kRowF4 + col) * 4, v.x, v.y, v.z, v.w);
                *vp = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    __syncwarp();
    for (int e = lane; e < kCap; e += 32) ws.keys[e] = kEmpty;
    __syncwarp();
}

__global__ void __launch_bounds__(kW2 * 32)
aggregate_bwd2_kernel(const int32_t *__restrict__ knn_idx, const float *__restrict__ counter, const float *__restrict__ gX,
                      long ldg, int m, int group_stride, long n_tasks, float *__restrict__ g_feats) {
    extern __shared__ __align__(16) unsigned char agg_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpScratch &ws = reinterpret_cast<WarpScratch *>(agg_smem)[wib];
    const int grp = lane / kRowF4, col = lane - grp * kRowF4;
    for (int e = lane; e < kCap; e += 32) ws.keys[e] = kEmpty;
    for (int e = lane; e < kCap * 36; e += 32) (&ws.vals[0][0])[e] = 0.f;
    __syncwarp();
    int count = 0;                                                   // used table entries (warp-uniform)
    const long warp_global = (long)blockIdx.x * kW2 + wib, total_warps = (long)gridDim.x * kW2;
    // consecutive tasks of a warp are consecutive sample indices of the same ray block: maximal vertex sharing
    const long per_warp = (n_tasks + total_warps - 1) / total_warps;
    const long t_begin = warp_global * per_warp, t_end = min(n_tasks, t_begin + per_warp);
    for (long t = t_begin; t < t_end; ++t) {
        const long j = t % group_stride, r0 = (t / group_stride) * 32;
        const long q = (r0 + lane) * group_stride + j;
        const bool valid = q < m;
        // ---- this lane's 40 attention weights (same arithmetic as attention())
        {
            float a[40];
            int id[40];
            float mn = INFINITY;
#pragma unroll
            for (int n4 = 0; n4 < 10; ++n4) {
                const int4 v = valid ? __ldg(reinterpret_cast<const int4 *>(knn_idx + q * 40) + n4) : make_int4(0, 0, 0, 0);
                id[4 * n4] = v.x; id[4 * n4 + 1] = v.y; id[4 * n4 + 2] = v.z; id[4 * n4 + 3] = v.w;
            }
#pragma unroll
            for (int n = 0; n < 40; ++n) { a[n] = __ldg(counter + id[n]); mn = fminf(mn, a[n]); }
            float mx = -INFINITY;
#pragma unroll
            for (int n = 0; n < 40; ++n) { a[n] = __fadd_rn(a[n], __fsub_rn(1.0f, mn)); mx = fmaxf(mx, a[n]); }
            float amax = -INFINITY;
#pragma unroll
            for (int n = 0; n < 40; ++n) { a[n] = __fdiv_rn(a[n], mx); amax = fmaxf(amax, a[n]); }
            float den = 0.f;
#pragma unroll
            for (int n = 0; n < 40; ++n) { a[n] = expf(a[n] - amax); den += a[n]; }
            const float inv = valid ? 1.0f / den : 0.f;
#pragma unroll
            for (int n = 0; n < 40; ++n) { ws.sidx[lane][n] = id[n]; ws.sw[lane][n] = a[n] * inv; }
        }
        // ---- this lane's gradient row (column 35 is the variance slot, not a feature)
#pragma unroll
        for (int c = 0; c < kRowF4; ++c) {
            float4 g = valid ? __ldg(reinterpret_cast<const float4 *>(gX + q * ldg) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (c == kRowF4 - 1) g.w = 0.f;
            *reinterpret_cast<float4 *>(&ws.grow[lane][c * 4]) = g;
        }
        __syncwarp();
        for (int s = 0; s < 32; ++s) {
            if (count > kCap - 40) { table_flush(ws, lane, g_feats); count = 0; }
#pragma unroll 1
            for (int it = 0; it < 16; ++it) {                        // 4 levels x (3+3+3+1) slots
                const int level = it >> 2, sub = (it & 3) * 3 + grp;
                bool act = grp < 3 && sub < 10;
                int v = kEmpty;
                float w = 0.f;
                if (act) {
                    v = ws.sidx[s][level * 10 + sub];
                    w = ws.sw[s][level * 10 + sub];
                    act = w != 0.f;
                }
                // find or claim the table entry of v (one CAS per group, by its first lane)
                uint32_t h = ((uint32_t)v * 2654435761u) >> 25;     // 7 bits
                bool found = !act;
                int fresh = 0;
                while (__any_sync(OCC_FULL, !found)) {
                    int ok = 1;
                    if (!found && col == 0) {
                        const int old = atomicCAS(&ws.keys[h], kEmpty, v);
                        ok = (old == kEmpty || old == v);
                        fresh += (old == kEmpty);
                    }
                    ok = __shfl_sync(OCC_FULL, ok, grp < 3 ? grp * kRowF4 : 27);
                    if (!found) { if (ok) found = true; else h = (h + 1) & (kCap - 1); }
                }
                count += __popc(__ballot_sync(OCC_FULL, fresh != 0));
                if (act) {
                    const float4 g = *reinterpret_cast<const float4 *>(&ws.grow[s][col * 4]);
                    float4 *vp = reinterpret_cast<float4 *>(&ws.vals[h][col * 4]);
                    float4 a = *vp;
                    a.x = fmaf(w, g.x, a.x); a.y = fmaf(w, g.y, a.y); a.z = fmaf(w, g.z, a.z); a.w = fmaf(w, g.w, a.w);
                    *vp = a;
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
    table_flush(ws, lane, g_feats);
}

}  // namespace

extern "C" int occnerf_aggregate_backward2(const int32_t *knn_idx, const float *point_counter, const float *gX, int ldg, int m,
                                           int nn, int group_stride, float *g_feats, occnerf_stream_t stream) {
    if (m <= 0) return OCCNERF_OK;
    OCC_CHECK_ARG(knn_idx && point_counter && gX && g_feats, "aggregate_backward2: null pointer");
    OCC_CHECK_ARG(nn == 40, "aggregate_backward2: nn=%d (built for 4 levels x 10 neighbours)", nn);
    OCC_CHECK_ARG(group_stride >= 1, "aggregate_backward2: group_stride=%d", group_stride);
    OCC_CHECK_ARG(ldg >= 36 && ldg % 4 == 0 && ((uintptr_t)gX & 15) == 0 && ((uintptr_t)g_feats & 15) == 0 &&
                  ((uintptr_t)knn_idx & 15) == 0, "aggregate_backward2: buffers must be 16-byte aligned with ldg %% 4 == 0");
    const size_t smem = sizeof(WarpScratch) * kW2;
    static bool configured = false;
    if (!configured) {
        OCC_CUDA(cudaFuncSetAttribute(aggregate_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int dev = 0, sms = 148;
    OCC_CUDA(cudaGetDevice(&dev));
    OCC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long rays = ((long)m + group_stride - 1) / group_stride;
    const long n_tasks = ((rays + 31) / 32) * group_stride;
    long grid = (n_tasks + kW2 - 1) / kW2;
    if (grid > sms) grid = sms;
    aggregate_bwd2_kernel<<<(unsigned)grid, kW2 * 32, smem, (cudaStream_t)stream>>>(knn_idx, point_counter, gX, ldg, m, group_stride,
                                                                                n_tasks, g_feats);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_aggregate_forward(const int32_t *knn_idx, const float *point_counter, const float *feats, int m,
                                         int nn, float *X, int ldx, occnerf_stream_t stream) {
    OCC_CHECK_ARG(knn_idx && point_counter && feats && X, "aggregate_forward: null pointer");
    OCC_CHECK_ARG(nn >= 2 && nn <= kMaxNN, "aggregate_forward: nn=%d outside [2,%d]", nn, kMaxNN);
    OCC_CHECK_ARG(ldx >= 36 && ldx % 4 == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)feats & 15) == 0,
                  "aggregate_forward: X/feats must be 16-byte aligned with ldx %% 4 == 0 (ldx=%d)", ldx);
    if (m <= 0) return OCCNERF_OK;
    aggregate_fwd_kernel<<<occ_div_up(m, kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(
        knn_idx, point_counter, (const float4 *)feats, m, nn, X, ldx);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_aggregate_backward(const int32_t *knn_idx, const float *point_counter, const float *gX, int ldg,
                                          int m, int nn, float *g_feats, occnerf_stream_t stream) {
    OCC_CHECK_ARG(knn_idx && point_counter && gX && g_feats, "aggregate_backward: null pointer");
    OCC_CHECK_ARG(nn >= 2 && nn <= kMaxNN, "aggregate_backward: nn=%d outside [2,%d]", nn, kMaxNN);
    OCC_CHECK_ARG(ldg >= 36 && ldg % 4 == 0 && ((uintptr_t)gX & 15) == 0 && ((uintptr_t)g_feats & 15) == 0,
                  "aggregate_backward: gX/g_feats must be 16-byte aligned with ldg %% 4 == 0 (ldg=%d)", ldg);
    if (m <= 0) return OCCNERF_OK;
    aggregate_bwd_kernel<<<occ_div_up(m, kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(knn_idx, point_counter, gX,
                                                                                         ldg, m, nn, g_feats);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
