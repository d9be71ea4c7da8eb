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
// g_feats[idx_n] += att_n * gX[0..34] with red.global.add.v4.f32 into privatised replicas of the table (all 31 M
// contributions per step land on 6890 rows and the L2 atomic units serialise per address: 64 replicas cut the kernel
// 3.2x.  Two shared-memory pre-reductions were tried and were slower: hash probing per CTA (latency-bound at 6 warps/SM),
// and a per-CTA 431 x 36 table for the vertices of the two coarsest levels, which receive half of all contributions
// (2.5 ms vs 1.0 ms: fp32 atomicAdd on shared memory is a compare-and-swap loop, slower than L2's native RED.ADD.F32)).
// What did work is merging runs of equal vertices along a ray in registers: aggregate_bwd_slot_kernel below.
// (The same idea for the FORWARD gathers -- neighbour rows cached in registers per slot, in three shapes: a warp walking
// the run, the same with a separate attention pass, and one thread per (run, level, column chunk) with shared-memory
// partial sums -- was 1.7-2.5x slower than the per-sample forward kernel, which already runs at 90 % of the L1 rate.
// Round 2: a bf16 copy of the table (80-byte rows, five lanes per row, six rows per warp instruction: half the gathered bytes and L1
// wavefronts; results fp32-exact against the oracle on the rounded table) bought 4 % (0.517 -> 0.496 ms per step): the forward is
// not limited by the gather but by the per-sample attention arithmetic (40 counter gathers, six warp reductions, exp, divisions);
// removed again.)
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
                     const float4 *__restrict__ feats, int m, int nn, float *__restrict__ X, long ldx, float *__restrict__ att_w) {
    __shared__ int s_idx[kWarps][kMaxNN];
    __shared__ float s_w[kWarps][kMaxNN];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long q = (long)blockIdx.x * kWarps + wib;
    if (q >= m) return;
    const float var = attention(knn_idx + q * nn, counter, nn, lane, s_idx[wib], s_w[wib]);
    if (att_w)                                                        // kept for the run-length backward
        for (int n = lane; n < nn; n += 32) att_w[q * nn + n] = s_w[wib][n];
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
                     const float *__restrict__ gX, long ldg, int m, int nn, float *__restrict__ g_feats,
                     int copies, long copy_stride) {
    __shared__ int s_idx[kWarps][kMaxNN];
    __shared__ float s_w[kWarps][kMaxNN];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long q = (long)blockIdx.x * kWarps + wib;
    if (q >= m) return;
    // privatised accumulators: every row of the 6890 x 36 table takes ~4500 reductions per step and the L2 atomic units
    // serialise per address, so CTAs are spread over `copies` replicas (summed afterwards)
    g_feats += (long)(blockIdx.x % copies) * copy_stride;
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
// Run-length backward for samples ordered along rays.  Measured on the bench workload (tools/exp_knn_persist.py): the
// neighbour in a given slot (level, rank) of sample q+1 is the SAME vertex as in sample q in 69 % (level 0) to 86 %
// (level 3) of the cases.  Thread = (run of kRun consecutive samples, 4 neighbour slots, 4-column chunk): it walks the run
// with three vector loads per sample (4 ids, 4 attention weights from the forward pass, its gradient chunk), adds w * g
// per slot in registers and issues one vector reduction per RUN OF EQUAL VERTICES instead of one per sample: 2.6x fewer
// L2 reductions (ncu).  Same sums.
constexpr int kRun = 16;

__global__ void __launch_bounds__(256, 4)
aggregate_bwd_slot_kernel(const int32_t *__restrict__ knn_idx, const float *__restrict__ att_w, const float *__restrict__ gX,
                          long ldg, int m, int nn, float *__restrict__ g_feats, int copies, long copy_stride) {
    const long gid = (long)blockIdx.x * 256 + threadIdx.x;
    const int per_run = (nn / 4) * kRowF4;
    const long run = gid / per_run;
    const long q0 = run * kRun;
    if (q0 >= m) return;
    const int r = (int)(gid - run * per_run);
    const int n4 = r / kRowF4, col = r - n4 * kRowF4;
    g_feats += (long)(blockIdx.x % copies) * copy_stride;
    const long q1 = q0 + kRun < m ? q0 + kRun : m;
    int cur[4] = {-1, -1, -1, -1};
    float4 sum[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) sum[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush = [&](int k) {
        if (cur[k] >= 0 && (sum[k].x != 0.f || sum[k].y != 0.f || sum[k].z != 0.f || sum[k].w != 0.f))
            red_add_v4(g_feats + ((size_t)cur[k] * kRowF4 + col) * 4, sum[k].x, sum[k].y, sum[k].z, sum[k].w);
    };
#pragma unroll 2
    for (long q = q0; q < q1; ++q) {
        float4 g = __ldg(reinterpret_cast<const float4 *>(gX + q * ldg) + col);
        if (col == kRowF4 - 1) g.w = 0.f;   // column 35 is the variance slot, not a feature
        const int4 vi = __ldg(reinterpret_cast<const int4 *>(knn_idx + q * nn) + n4);
        const float4 wf = __ldg(reinterpret_cast<const float4 *>(att_w + q * nn) + n4);
        const int v[4] = {vi.x, vi.y, vi.z, vi.w};
        const float w[4] = {wf.x, wf.y, wf.z, wf.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (v[k] != cur[k]) {
                flush(k);
                cur[k] = v[k];
                sum[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            sum[k].x = fmaf(w[k], g.x, sum[k].x); sum[k].y = fmaf(w[k], g.y, sum[k].y);
            sum[k].z = fmaf(w[k], g.z, sum[k].z); sum[k].w = fmaf(w[k], g.w, sum[k].w);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) flush(k);
}

}  // namespace

extern "C" int occnerf_aggregate_forward(const int32_t *knn_idx, const float *point_counter, const float *feats, int m,
                                         int nn, float *X, int ldx, float *att_w, occnerf_stream_t stream) {
    OCC_CHECK_ARG(knn_idx && point_counter && feats && X, "aggregate_forward: null pointer");
    OCC_CHECK_ARG(nn >= 2 && nn <= kMaxNN, "aggregate_forward: nn=%d outside [2,%d]", nn, kMaxNN);
    OCC_CHECK_ARG(ldx >= 36 && ldx % 4 == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)feats & 15) == 0,
                  "aggregate_forward: X/feats must be 16-byte aligned with ldx %% 4 == 0 (ldx=%d)", ldx);
    if (m <= 0) return OCCNERF_OK;
    aggregate_fwd_kernel<<<occ_div_up(m, kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(
        knn_idx, point_counter, (const float4 *)feats, m, nn, X, ldx, att_w);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_aggregate_backward(const int32_t *knn_idx, const float *point_counter, const float *gX, int ldg,
                                          int m, int nn, float *g_feats, int V, int copies, const float *att_w,
                                          occnerf_stream_t stream) {
    OCC_CHECK_ARG(knn_idx && point_counter && gX && g_feats, "aggregate_backward: null pointer");
    OCC_CHECK_ARG(nn >= 2 && nn <= kMaxNN, "aggregate_backward: nn=%d outside [2,%d]", nn, kMaxNN);
    OCC_CHECK_ARG(ldg >= 36 && ldg % 4 == 0 && ((uintptr_t)gX & 15) == 0 && ((uintptr_t)g_feats & 15) == 0,
                  "aggregate_backward: gX/g_feats must be 16-byte aligned with ldg %% 4 == 0 (ldg=%d)", ldg);
    if (m <= 0) return OCCNERF_OK;
    OCC_CHECK_ARG(copies >= 1 && V >= 1, "aggregate_backward: copies=%d V=%d", copies, V);
    if (att_w && nn % 4 == 0 && ((uintptr_t)knn_idx & 15) == 0 && ((uintptr_t)att_w & 15) == 0) {
        const long threads = occ_div_up(m, kRun) * (long)(nn / 4) * kRowF4;
        aggregate_bwd_slot_kernel<<<occ_div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(knn_idx, att_w, gX, ldg, m, nn, g_feats,
                                                                                            copies, (long)V * 36);
    } else {
        aggregate_bwd_kernel<<<occ_div_up(m, kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(knn_idx, point_counter, gX,
                                                                                             ldg, m, nn, g_feats, copies, (long)V * 36);
    }
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
