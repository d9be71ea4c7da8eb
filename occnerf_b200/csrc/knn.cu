// Exact k-nearest-neighbour search, per-sample surface geometry and the visibility vote.
//
// Replaces the pykeops `Kmin_argKmin` reductions the reference calls through core/nets/occnerf/knn.py:33-85
// (network.py:236-255: 4 block-diagonal levels x k=10 per sample; :265 k=3 per vertex; :508 k=10 per
// terminating ray) and the no_grad geometry block of canonical_mlps/occnerf_mlp.py:146-167.
//
// v1 is tiled brute force on the fp32 CUDA cores (tensor cores are reserved for the MLP): the support set is
// streamed through shared memory in 2048-point tiles as float4 and read with warp-broadcast LDS.128; every
// thread owns one query and a sorted register top-k per level.  The ranking key is
// (dx*dx + dy*dy) + dz*dz with explicit round-to-nearest ops (no FMA contraction) and ties go to the lower
// support row, which makes the neighbour ids bit-exact with the oracle.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 2048;

template <int K>
__device__ __forceinline__ void topk_insert(float (&dk)[K], int (&ik)[K], float d, int idx) {
    bool ins = false;
    float cd = d;
    int ci = idx;
#pragma unroll
    for (int t = 0; t < K; ++t) {
        ins = ins || (cd < dk[t]);
        if (ins) {
            const float td = dk[t];
            const int ti = ik[t];
            dk[t] = cd; ik[t] = ci;
            cd = td; ci = ti;
        }
    }
}

template <int K>
__global__ void __launch_bounds__(kThreads)
knn_kernel(const float *__restrict__ queries, int m, const float4 *__restrict__ supports, const int32_t *__restrict__ gid,
           int lv0, int lv1, int lv2, int lv3, int lv4, int n_levels, const uint8_t *__restrict__ query_sel,
           int32_t *__restrict__ out) {
    __shared__ float4 tile[kTile];
    const int q = blockIdx.x * kThreads + threadIdx.x;
    bool active = q < m;
    if (active && query_sel) active = query_sel[q] != 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        qx = __ldg(queries + (size_t)q * 3 + 0);
        qy = __ldg(queries + (size_t)q * 3 + 1);
        qz = __ldg(queries + (size_t)q * 3 + 2);
    }
    const int begins[5] = {lv0, lv1, lv2, lv3, lv4};
#pragma unroll 1
    for (int lev = 0; lev < n_levels; ++lev) {
        const int s0 = begins[lev], s1 = begins[lev + 1];
        float dk[K];
        int ik[K];
#pragma unroll
        for (int t = 0; t < K; ++t) { dk[t] = INFINITY; ik[t] = -1; }
#pragma unroll 1
        for (int base = s0; base < s1; base += kTile) {
            const int n = min(kTile, s1 - base);
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += kThreads) tile[i] = __ldg(supports + base + i);
            __syncthreads();
            if (active) {
#pragma unroll 4
                for (int i = 0; i < n; ++i) {
                    const float4 s = tile[i];
                    const float dx = __fsub_rn(qx, s.x), dy = __fsub_rn(qy, s.y), dz = __fsub_rn(qz, s.z);
                    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    if (d < dk[K - 1]) topk_insert<K>(dk, ik, d, base + i);
                }
            }
        }
        if (active) {
            int32_t *o = out + ((size_t)q * n_levels + lev) * K;
#pragma unroll
            for (int t = 0; t < K; ++t) {
                int v = ik[t];
                if (v >= 0) v = gid ? __ldg(gid + v) : v - s0;
                o[t] = v;
            }
        }
    }
}

// Warp-per-query variant of the brute-force search for SMALL query sets (the per-vertex block: 6890 queries, k = 3; the
// visibility vote: <= one query per ray, k = 10).  One thread per query would leave most of the GPU idle there; here the 32
// lanes of a warp scan interleaved slices of the support set (coalesced 512-byte reads), each keeps its own sorted top-k,
// and k rounds of a lexicographic (distance, row) warp arg-min merge the 32 lists.  Same arithmetic and tie rule as
// knn_kernel, so the ids are identical.
template <int K>
__global__ void __launch_bounds__(kThreads)
knn_warp_kernel(const float *__restrict__ queries, int m, const float4 *__restrict__ supports, const int32_t *__restrict__ gid,
                int lv0, int lv1, int lv2, int lv3, int lv4, int n_levels, const uint8_t *__restrict__ query_sel,
                int32_t *__restrict__ out) {
    const int q = (blockIdx.x * kThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= m) return;
    if (query_sel && query_sel[q] == 0) return;
    const float qx = __ldg(queries + (size_t)q * 3 + 0), qy = __ldg(queries + (size_t)q * 3 + 1), qz = __ldg(queries + (size_t)q * 3 + 2);
    const int begins[5] = {lv0, lv1, lv2, lv3, lv4};
#pragma unroll 1
    for (int lev = 0; lev < n_levels; ++lev) {
        const int s0 = begins[lev], s1 = begins[lev + 1];
        float dk[K];
        int ik[K];
#pragma unroll
        for (int t = 0; t < K; ++t) { dk[t] = INFINITY; ik[t] = 0x7fffffff; }
#pragma unroll 2
        for (int i = s0 + lane; i < s1; i += 32) {
            const float4 s = __ldg(supports + i);
            const float dx = __fsub_rn(qx, s.x), dy = __fsub_rn(qy, s.y), dz = __fsub_rn(qz, s.z);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d < dk[K - 1]) topk_insert<K>(dk, ik, d, i);        // rows ascend within a lane: strict < keeps the lower row on ties
        }
        int mine = -1;                                               // lane t ends up with the t-th neighbour
#pragma unroll 1
        for (int t = 0; t < K; ++t) {
            float bd = dk[0];
            int bi = ik[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(OCC_FULL, bd, o);
                const int oi = __shfl_xor_sync(OCC_FULL, bi, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if (bi == ik[0] && bi != 0x7fffffff) {                   // this lane's head won: pop it
#pragma unroll
                for (int u = 0; u + 1 < K; ++u) { dk[u] = dk[u + 1]; ik[u] = ik[u + 1]; }
                dk[K - 1] = INFINITY; ik[K - 1] = 0x7fffffff;
            }
            if (lane == t) mine = bi;
        }
        if (lane < K) {
            int v = mine;
            v = (v == 0x7fffffff) ? -1 : (gid ? __ldg(gid + v) : v - s0);
            out[((size_t)q * n_levels + lev) * K + lane] = v;
        }
    }
}

// occnerf_mlp.py:146-167
__global__ void __launch_bounds__(kThreads)
sample_geometry_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ knn_idx, int knn_stride,
                       const float *__restrict__ base, const float *__restrict__ norms, float bound, int m,
                       float *__restrict__ enc_in, float *__restrict__ dist_out, int dist_stride) {
    const int q = blockIdx.x * kThreads + threadIdx.x;
    if (q >= m) return;
    const float x = __ldg(xyz + (size_t)q * 3), y = __ldg(xyz + (size_t)q * 3 + 1), z = __ldg(xyz + (size_t)q * 3 + 2);
    const int32_t *idx = knn_idx + (size_t)q * knn_stride;
    int inside = 0;
    float nsum = 0.f, asum = 0.f, px = 0.f, py = 0.f, pz = 0.f;
    const float inv2b = 2.0f * bound;
#pragma unroll
    for (int t = 0; t < 10; ++t) {
        const int v = __ldg(idx + t);
        const float bx = __ldg(base + (size_t)v * 3), by = __ldg(base + (size_t)v * 3 + 1), bz = __ldg(base + (size_t)v * 3 + 2);
        const float nx = __ldg(norms + (size_t)v * 3), ny = __ldg(norms + (size_t)v * 3 + 1), nz = __ldg(norms + (size_t)v * 3 + 2);
        const float dx = __fsub_rn(x, bx), dy = __fsub_rn(y, by), dz = __fsub_rn(z, bz);
        const double dot = ((double)dx * (double)nx + (double)dy * (double)ny) + (double)dz * (double)nz;
        inside += dot < 0.0;
        const float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        nsum = __fadd_rn(nsum, dn);
        if (t < 3) {
            // |cosine_similarity(dir, n)| with eps = 1e-8 on each norm, then the weighted mean of the
            // neighbour positions normalised to [0,1]^3
            const float nn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
            const float dd = fmaxf(dn, 1e-8f), nd_ = fmaxf(nn, 1e-8f);
            const float a = fabsf((dx / dd) * (nx / nd_) + (dy / dd) * (ny / nd_) + (dz / dd) * (nz / nd_));
            asum += a;
            px += a * (__fadd_rn(bx, bound) / inv2b);
            py += a * (__fadd_rn(by, bound) / inv2b);
            pz += a * (__fadd_rn(bz, bound) / inv2b);
        }
    }
    float dist = nsum / 10.0f;
    if (inside > 5) dist = -dist;
    const float nd = fminf(fmaxf(__fdiv_rn(__fadd_rn(dist, 0.2f), 0.5f), 0.0f), 1.0f);
    reinterpret_cast<float4 *>(enc_in)[q] = make_float4(px / asum, py / asum, pz / asum, nd);
    dist_out[(size_t)q * dist_stride] = dist;
}

// network.py:502-517
__global__ void vis_select_kernel(const float *__restrict__ depth, const int64_t *__restrict__ term,
                                  const float *__restrict__ x_skel, int N, int S, float thresh, int *__restrict__ count,
                                  uint8_t *__restrict__ sel, float *__restrict__ qpts) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const bool s = __ldg(depth + r) > thresh;
    sel[r] = s ? 1 : 0;
    long t = (long)__ldg(term + r);
    t = t < 0 ? 0 : (t >= S ? S - 1 : t);
    const float *p = x_skel + ((size_t)r * S + t) * 3;
    qpts[r * 3 + 0] = p[0]; qpts[r * 3 + 1] = p[1]; qpts[r * 3 + 2] = p[2];
    if (s) atomicAdd(count, 1);
}
__global__ void vis_mark_kernel(const int *__restrict__ count, const uint8_t *__restrict__ sel,
                                const int32_t *__restrict__ idx, int N, int k, float *__restrict__ hits) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N || *count <= 1 || !sel[r]) return;
    for (int t = 0; t < k; ++t) {
        const int v = idx[(size_t)r * k + t];
        if (v >= 0) hits[v] = 1.0f;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Cluster-pruned exact search over a two-level hierarchy (fine level clustered around the points of a coarse level).
// OccNeRF's multi-scale neighbourhoods are exactly that: level 2 (431 FPS points) are natural cluster centres for
// level 0 (6890 vertices), level 3 (108) for level 1 (1723).  Per query:
//   pass 1  distances to all centres  -> the coarse level's own k-NN (written out) and the nearest centre;
//   seed    scan the nearest centre's members -> an upper bound U on the k-th fine distance;
//   pass 2  every other cluster whose lower bound  |q-c| - r_c  can still beat U is scanned, the rest is skipped.
// Lanes of a warp are 32 consecutive RAYS at the same sample index (group_stride = samples per ray), a few cm apart,
// so they need the same handful of clusters: the loop over (cluster, member) is warp-uniform (broadcast LDS.128,
// per-lane predicate) and only ~1/6 of the brute-force distance evaluations remain.
// Exactness: member distances use the same arithmetic as the brute-force kernel; a cluster is skipped only if its
// lower bound exceeds sqrt(U) by a margin (1e-5 relative + 1e-6) far above fp32 rounding, radii are inflated the same
// way on the host, and ties are broken on (distance, original row) so the visiting order does not matter.
template <int K>
__device__ __forceinline__ void topk_insert_lex(float (&dk)[K], int (&ik)[K], float d, int idx) {
    bool ins = false;
    float cd = d;
    int ci = idx;
#pragma unroll
    for (int t = 0; t < K; ++t) {
        ins = ins || (cd < dk[t]) || (cd == dk[t] && ci < ik[t]);
        if (ins) {
            const float td = dk[t];
            const int ti = ik[t];
            dk[t] = cd; ik[t] = ci;
            cd = td; ci = ti;
        }
    }
}

__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, const float4 &s) {
    const float dx = __fsub_rn(qx, s.x), dy = __fsub_rn(qy, s.y), dz = __fsub_rn(qz, s.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// A cluster (centre c, radius r >= every member's distance to c, inflated on the host) can be skipped when
// |q-c| - r > sqrt(U), U = the current k-th squared distance.  `prune_reach` turns U into the reach sqrt(U) (+ margins far
// above fp32 rounding) once per bound update; the per-cluster test is then sqrt-free: |q-c|^2 > (reach + r)^2.
__device__ __forceinline__ float prune_reach(float worst_d2) { return sqrtf(worst_d2) * 1.00001f + 1e-6f; }
__device__ __forceinline__ bool cannot_prune(float d2_centre, float radius, float reach) {
    const float t = (reach + radius) * 1.000001f;
    return !(d2_centre > t * t);
}

constexpr int kHierThreads = 1024;   // one CTA per SM (the fine level fills most of shared memory) -> 32 warps/SM

template <int K>
__global__ void __launch_bounds__(kHierThreads)
knn_hier_kernel(const float *__restrict__ queries, int m, int group_stride, const float4 *__restrict__ fine4,
                const float4 *__restrict__ centers4, const int2 *__restrict__ ranges, int nf, int nc,
                const int32_t *__restrict__ fine_gid, const int32_t *__restrict__ center_gid, int32_t *__restrict__ out_fine,
                int32_t *__restrict__ out_center, int out_stride) {
    extern __shared__ __align__(16) unsigned char hier_smem[];
    float4 *s_fine = reinterpret_cast<float4 *>(hier_smem);
    float4 *s_cent = s_fine + nf;
    int2 *s_rng = reinterpret_cast<int2 *>(s_cent + nc);
    for (int i = threadIdx.x; i < nf; i += kHierThreads) s_fine[i] = __ldg(fine4 + i);
    for (int i = threadIdx.x; i < nc; i += kHierThreads) { s_cent[i] = __ldg(centers4 + i); s_rng[i] = __ldg(ranges + i); }
    __syncthreads();
    // warp -> (sample index j, block of 32 rays); lane -> ray
    const long warp_global = ((long)blockIdx.x * kHierThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long j = warp_global % group_stride, r0 = (warp_global / group_stride) * 32;
    const long q = (r0 + lane) * group_stride + j;
    const bool active = q < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { qx = __ldg(queries + q * 3); qy = __ldg(queries + q * 3 + 1); qz = __ldg(queries + q * 3 + 2); }

    float dk[K];
    int ik[K];
    // ---- pass 1: the coarse level itself
#pragma unroll
    for (int t = 0; t < K; ++t) { dk[t] = INFINITY; ik[t] = 0x7fffffff; }
#pragma unroll 4
    for (int c = 0; c < nc; ++c) {
        const float d = dist2_rn(qx, qy, qz, s_cent[c]);
        if (d < dk[K - 1]) topk_insert<K>(dk, ik, d, c);
    }
    const int seed = ik[0];
    if (active) {
        int32_t *o = out_center + q * out_stride;
#pragma unroll
        for (int t = 0; t < K; ++t) o[t] = (ik[t] == 0x7fffffff) ? -1 : (center_gid ? __ldg(center_gid + ik[t]) : ik[t]);
    }
    // ---- seed: members of the nearest centre (warp-uniform loop over the union of the lanes' seeds)
#pragma unroll
    for (int t = 0; t < K; ++t) { dk[t] = INFINITY; ik[t] = 0x7fffffff; }
    {
        unsigned todo = __ballot_sync(OCC_FULL, active);
        while (todo) {
            const int c = __shfl_sync(OCC_FULL, seed, __ffs(todo) - 1);
            const bool need = active && seed == c;
            const int2 rg = s_rng[c];
            for (int i = rg.x; i < rg.x + rg.y; ++i) {
                const float4 p = s_fine[i];
                const float d = dist2_rn(qx, qy, qz, p);
                const int row = __float_as_int(p.w);
                if (need && (d < dk[K - 1] || (d == dk[K - 1] && row < ik[K - 1]))) topk_insert_lex<K>(dk, ik, d, row);
            }
            todo &= ~__ballot_sync(OCC_FULL, need);
        }
    }
    // ---- pass 2: every other cluster that can still contain one of the k nearest
    float reach = prune_reach(dk[K - 1]);
#pragma unroll 2
    for (int c = 0; c < nc; ++c) {
        const float4 cc = s_cent[c];
        const float t = (reach + cc.w) * 1.000001f;
        const bool need = active && c != seed && !(dist2_rn(qx, qy, qz, cc) > t * t);
        if (!__any_sync(OCC_FULL, need)) continue;
        const int2 rg = s_rng[c];
        for (int i = rg.x; i < rg.x + rg.y; ++i) {
            const float4 p = s_fine[i];
            const float d = dist2_rn(qx, qy, qz, p);
            const int row = __float_as_int(p.w);
            if (need && (d < dk[K - 1] || (d == dk[K - 1] && row < ik[K - 1]))) topk_insert_lex<K>(dk, ik, d, row);
        }
        reach = prune_reach(dk[K - 1]);
    }
    if (active) {
        int32_t *o = out_fine + q * out_stride;
#pragma unroll
        for (int t = 0; t < K; ++t) o[t] = (ik[t] == 0x7fffffff) ? -1 : (fine_gid ? __ldg(fine_gid + ik[t]) : ik[t]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// All four levels in one kernel, through a three-level cluster tree: level 3 (108 points) clusters level 2 (431) and
// level 1 (1723); level 2 clusters level 0 (6890).  Level-0 search prunes twice: a level-3 cluster is skipped with the
// super-radius max(|child - c3| + r_child), otherwise its level-2 children are tested one by one.  ~0.8 k distance
// evaluations per query instead of 9152, all of shared memory holds the 9152 points (146 KB) plus the tables.
struct TreeArgs {
    const float4 *p0s;      // [n0] level-0 points sorted by level-2 cluster (P2s order), .w = vertex id
    const float4 *p1s;      // [n1] level-1 points sorted by level-3 cluster, .w = row in level 1
    const float4 *p2s;      // [n2] level-2 points sorted by level-3 cluster, .w = row in level 2
    const float4 *p3;       // [n3] level-3 points, .w unused
    const float4 *c2tab;    // [n2] per P2s entry: (r20 = radius of its level-0 cluster, begin0, count0, unused) as float bits
    const float4 *c3tab;    // [n3] (r32, r31, R30, unused)
    const int4 *c3rng;      // [n3] (begin2, count2, begin1, count1)
    const int32_t *gid1, *gid2, *gid3;   // level row -> vertex id
    const int32_t *inv2;    // [n2] level-2 row -> position in p2s
    int n0, n1, n2, n3;
};

template <int K>
__device__ __forceinline__ void scan_members(const float4 *__restrict__ pts, int begin, int count, bool need, float qx, float qy,
                                             float qz, float (&dk)[K], int (&ik)[K]) {
#pragma unroll 2
    for (int i = begin; i < begin + count; ++i) {
        const float4 p = pts[i];
        const float d = dist2_rn(qx, qy, qz, p);
        const int row = __float_as_int(p.w);
        if (need && (d < dk[K - 1] || (d == dk[K - 1] && row < ik[K - 1]))) topk_insert_lex<K>(dk, ik, d, row);
    }
}
template <int K>
__device__ __forceinline__ void reset_topk(float (&dk)[K], int (&ik)[K]) {
#pragma unroll
    for (int t = 0; t < K; ++t) { dk[t] = INFINITY; ik[t] = 0x7fffffff; }
}
template <int K>
__device__ __forceinline__ void write_topk(int32_t *__restrict__ o, const int (&ik)[K], const int32_t *__restrict__ gid) {
#pragma unroll
    for (int t = 0; t < K; ++t) o[t] = (ik[t] == 0x7fffffff) ? -1 : (gid ? __ldg(gid + ik[t]) : ik[t]);
}

template <int K>
__global__ void __launch_bounds__(kHierThreads)
knn_tree_kernel(const float *__restrict__ queries, int m, int group_stride, int lane_rays, const TreeArgs T,
                int32_t *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char hier_smem[];
    float4 *s0 = reinterpret_cast<float4 *>(hier_smem);
    float4 *s1 = s0 + T.n0, *s2 = s1 + T.n1, *s3 = s2 + T.n2, *sc2 = s3 + T.n3, *sc3 = sc2 + T.n2;
    int4 *sr3 = reinterpret_cast<int4 *>(sc3 + T.n3);
    for (int i = threadIdx.x; i < T.n0; i += kHierThreads) s0[i] = __ldg(T.p0s + i);
    for (int i = threadIdx.x; i < T.n1; i += kHierThreads) s1[i] = __ldg(T.p1s + i);
    for (int i = threadIdx.x; i < T.n2; i += kHierThreads) { s2[i] = __ldg(T.p2s + i); sc2[i] = __ldg(T.c2tab + i); }
    for (int i = threadIdx.x; i < T.n3; i += kHierThreads) { s3[i] = __ldg(T.p3 + i); sc3[i] = __ldg(T.c3tab + i); sr3[i] = __ldg(T.c3rng + i); }
    __syncthreads();
    // lane -> query: a warp covers `lane_rays` consecutive rays x (32 / lane_rays) consecutive samples, i.e. a compact patch
    // of space (adjacent pixels are millimetres apart, adjacent samples ~1.5 cm), so its lanes need the same few clusters
    const int lane = threadIdx.x & 31;
    const int lane_samples = 32 / lane_rays;
    const long jblocks = (group_stride + lane_samples - 1) / lane_samples;
    // consecutive warps of a CTA take consecutive sample depths of the same rays: they run the same phases of the search at
    // about the same time, which keeps the (large, unrolled) code of this kernel hot in the instruction cache.  Measured:
    // a persistent grid-stride version that gives every warp a pseudo-random sequence of items is 1.4-1.7x SLOWER.
    const long warp_global = ((long)blockIdx.x * kHierThreads + threadIdx.x) >> 5;
    {
    const long j = (warp_global % jblocks) * lane_samples + lane / lane_rays;
    const long q = ((warp_global / jblocks) * lane_rays + lane % lane_rays) * group_stride + j;
    const bool active = j < group_stride && q < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { qx = __ldg(queries + q * 3); qy = __ldg(queries + q * 3 + 1); qz = __ldg(queries + q * 3 + 2); }
    int32_t *o = out + (active ? q : 0) * 4 * K;
    float dk[K];
    int ik[K];

    // ---- level 3: brute force over the 108 roots
    reset_topk<K>(dk, ik);
#pragma unroll 4
    for (int c = 0; c < T.n3; ++c) {
        const float d = dist2_rn(qx, qy, qz, s3[c]);
        if (d < dk[K - 1]) topk_insert<K>(dk, ik, d, c);
    }
    const int seed3 = ik[0];
    if (active) write_topk<K>(o + 3 * K, ik, T.gid3);

    // ---- levels 2 and 1: clusters around the level-3 roots
    int seed2 = 0;                                                // nearest level-2 point, as a position in p2s
#pragma unroll 1
    for (int lev = 2; lev >= 1; --lev) {
        const float4 *pts = lev == 2 ? s2 : s1;
        reset_topk<K>(dk, ik);
        unsigned todo = __ballot_sync(OCC_FULL, active);
        while (todo) {                                            // seeds: every lane's nearest root first
            const int c = __shfl_sync(OCC_FULL, seed3, __ffs(todo) - 1);
            const bool need = active && seed3 == c;
            const int4 rg = sr3[c];
            scan_members<K>(pts, lev == 2 ? rg.x : rg.z, lev == 2 ? rg.y : rg.w, need, qx, qy, qz, dk, ik);
            todo &= ~__ballot_sync(OCC_FULL, need);
        }
        float reach = prune_reach(dk[K - 1]);
        for (int c = 0; c < T.n3; ++c) {
            const float4 tab = sc3[c];
            const bool need = active && c != seed3 && cannot_prune(dist2_rn(qx, qy, qz, s3[c]), lev == 2 ? tab.x : tab.y, reach);
            if (!__any_sync(OCC_FULL, need)) continue;
            const int4 rg = sr3[c];
            scan_members<K>(pts, lev == 2 ? rg.x : rg.z, lev == 2 ? rg.y : rg.w, need, qx, qy, qz, dk, ik);
            reach = prune_reach(dk[K - 1]);
        }
        if (active) write_topk<K>(o + lev * K, ik, lev == 2 ? T.gid2 : T.gid1);
        if (lev == 2 && active && ik[0] != 0x7fffffff) seed2 = __ldg(T.inv2 + ik[0]);
    }

    // ---- level 0: two-level descent, level-3 clusters -> their level-2 children -> level-0 members
    reset_topk<K>(dk, ik);
    {
        unsigned todo = __ballot_sync(OCC_FULL, active);
        while (todo) {                                            // seeds: the level-0 cluster of every lane's nearest level-2 point
            const int i2 = __shfl_sync(OCC_FULL, seed2, __ffs(todo) - 1);
            const bool need = active && seed2 == i2;
            const float4 ct = sc2[i2];
            scan_members<K>(s0, __float_as_int(ct.y), __float_as_int(ct.z), need, qx, qy, qz, dk, ik);
            todo &= ~__ballot_sync(OCC_FULL, need);
        }
    }
    float reach = prune_reach(dk[K - 1]);
    for (int c = 0; c < T.n3; ++c) {
        const bool need3 = active && cannot_prune(dist2_rn(qx, qy, qz, s3[c]), sc3[c].z, reach);
        if (!__any_sync(OCC_FULL, need3)) continue;
        const int4 rg = sr3[c];
        for (int i2 = rg.x; i2 < rg.x + rg.y; ++i2) {
            const float4 ct = sc2[i2];
            const bool need = need3 && i2 != seed2 && cannot_prune(dist2_rn(qx, qy, qz, s2[i2]), ct.x, reach);
            if (!__any_sync(OCC_FULL, need)) continue;
            scan_members<K>(s0, __float_as_int(ct.y), __float_as_int(ct.z), need, qx, qy, qz, dk, ik);
            reach = prune_reach(dk[K - 1]);
        }
    }
    if (active) write_topk<K>(o, ik, nullptr);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Candidate-list search over a static uniform grid.  The support sets never move (point_base is not trained), so for
// every cell of a grid over the canonical volume and every level the host side precomputes (once per subject,
// occnerf_b200.ops.build_knn_grid) the list of all points that can be among the k nearest of ANY query inside the cell:
//     L(cell) = { p : |p - centre| <= d_k(centre) + 2*rho + margin },   rho = half the cell diagonal
// (for q in the cell, d_k(q) <= d_k(centre) + rho, and a k-nearest p of q has |p - centre| <= |p - q| + rho), sorted by
// distance to the centre.  A query then scans only its cell's list -- a few hundred candidates for all levels together
// instead of ~1-1.4 k cluster-pruned (or 9152 brute-force) ones, with no pruning tests at all, and because the list is
// nearest-first the running top-k is final after a handful of insertions.  The ranking is the same (distance, row)
// lexicographic order on the same fp32 distances, so the ids are bit-identical to occnerf_knn.  Queries outside the grid
// fall back to an exact scan of the whole level inside the same kernel.
struct GridArgs {
    const float4 *p[4];      // level points in their own row order (.w unused)
    int n[4];
    const int32_t *gid[4];   // level row -> vertex id (gid[0] unused: rows ARE vertex ids)
    const int2 *cell_tab;    // [cells][4] (offset into lists, count) per level
    const uint16_t *lists;   // level-local rows
    float gmin[3], inv_h;
    int dims[3];
};

template <int K>
__global__ void __launch_bounds__(kHierThreads)
knn_grid_kernel(const float *__restrict__ queries, int m, int group_stride, int lane_rays, const GridArgs G,
                int32_t *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char hier_smem[];
    float4 *s0 = reinterpret_cast<float4 *>(hier_smem);
    float4 *s1 = s0 + G.n[0], *s2 = s1 + G.n[1], *s3 = s2 + G.n[2];
    for (int i = threadIdx.x; i < G.n[0]; i += kHierThreads) s0[i] = __ldg(G.p[0] + i);
    for (int i = threadIdx.x; i < G.n[1]; i += kHierThreads) s1[i] = __ldg(G.p[1] + i);
    for (int i = threadIdx.x; i < G.n[2]; i += kHierThreads) s2[i] = __ldg(G.p[2] + i);
    for (int i = threadIdx.x; i < G.n[3]; i += kHierThreads) s3[i] = __ldg(G.p[3] + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int lane_samples = 32 / lane_rays;
    const long jblocks = (group_stride + lane_samples - 1) / lane_samples;
    const long warp_global = ((long)blockIdx.x * kHierThreads + threadIdx.x) >> 5;
    const long j = (warp_global % jblocks) * lane_samples + lane / lane_rays;
    const long q = ((warp_global / jblocks) * lane_rays + lane % lane_rays) * group_stride + j;
    const bool active = j < group_stride && q < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { qx = __ldg(queries + q * 3); qy = __ldg(queries + q * 3 + 1); qz = __ldg(queries + q * 3 + 2); }
    int32_t *o = out + (active ? q : 0) * 4 * K;
    float dk[K];
    int ik[K];

    // ---- the query's cell
    const float fx = (qx - G.gmin[0]) * G.inv_h, fy = (qy - G.gmin[1]) * G.inv_h, fz = (qz - G.gmin[2]) * G.inv_h;
    const bool in_grid = active && fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)G.dims[0] && fy < (float)G.dims[1] &&
                         fz < (float)G.dims[2];
    const long cell = in_grid ? ((long)(int)fz * G.dims[1] + (int)fy) * G.dims[0] + (int)fx : 0;
    const bool stray = active && !in_grid;

#pragma unroll 1
    for (int lev = 3; lev >= 0; --lev) {
        const float4 *pts = lev == 0 ? s0 : (lev == 1 ? s1 : (lev == 2 ? s2 : s3));
        const int n_lev = lev == 0 ? G.n[0] : (lev == 1 ? G.n[1] : (lev == 2 ? G.n[2] : G.n[3]));
        reset_topk<K>(dk, ik);
        int2 oc = make_int2(0, 0);
        if (in_grid) oc = __ldg(G.cell_tab + cell * 4 + lev);
        const uint16_t *lst = G.lists + oc.x;
        const int cnt = oc.y;
        const int cmax = __reduce_max_sync(OCC_FULL, cnt);
        // the first K candidates always enter the (empty) top-k: take them as they come and sort them once with a 29-comparator
        // network (232 instructions) instead of K sorted insertions (~800)
        int first = 0;
        if (K == 10 && __all_sync(OCC_FULL, cnt == 0 || cnt >= K)) {
#pragma unroll
            for (int t = 0; t < K; ++t) {
                if (cnt > 0) {
                    const int row = (int)__ldg(lst + t);
                    dk[t] = dist2_rn(qx, qy, qz, pts[row]);
                    ik[t] = row;
                }
            }
            auto cex = [&](int a, int b) {       // compare-exchange on (distance, row)
                const bool swap = dk[b] < dk[a] || (dk[b] == dk[a] && ik[b] < ik[a]);
                const float da = dk[a], db = dk[b];
                const int ia = ik[a], ib = ik[b];
                dk[a] = swap ? db : da; dk[b] = swap ? da : db;
                ik[a] = swap ? ib : ia; ik[b] = swap ? ia : ib;
            };
            cex(0, 8); cex(1, 9); cex(2, 7); cex(3, 5); cex(4, 6);
            cex(0, 2); cex(1, 4); cex(5, 8); cex(7, 9);
            cex(0, 3); cex(2, 4); cex(5, 7); cex(6, 9);
            cex(0, 1); cex(3, 6); cex(8, 9);
            cex(1, 5); cex(2, 3); cex(4, 8); cex(6, 7);
            cex(1, 2); cex(3, 5); cex(4, 6); cex(7, 8);
            cex(2, 3); cex(4, 5); cex(6, 7);
            cex(3, 4); cex(5, 6);
            first = K;
        }
#pragma unroll 2
        for (int i = first; i < cmax; ++i) {
            if (i < cnt) {
                const int row = (int)__ldg(lst + i);
                const float d = dist2_rn(qx, qy, qz, pts[row]);
                if (d < dk[K - 1] || (d == dk[K - 1] && row < ik[K - 1])) topk_insert_lex<K>(dk, ik, d, row);
            }
        }
        if (__any_sync(OCC_FULL, stray)) {                  // outside the grid: exact scan of the whole level
            for (int c = 0; c < n_lev; ++c) {
                const float d = dist2_rn(qx, qy, qz, pts[c]);
                if (stray && d < dk[K - 1]) topk_insert<K>(dk, ik, d, c);
            }
        }
        if (active) write_topk<K>(o + lev * K, ik, lev == 0 ? nullptr : (lev == 1 ? G.gid[1] : (lev == 2 ? G.gid[2] : G.gid[3])));
    }
}

int launch_knn(const float *queries, int m, const float *supports4, const int32_t *gid, const int32_t *lb, int n_levels,
               int k, const uint8_t *query_sel, int32_t *out, cudaStream_t st) {
    int b[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i <= n_levels; ++i) b[i] = lb[i];
    if (m <= 32768) {                                                // few queries: one warp each
        const unsigned wgrid = occ_div_up((long)m * 32, kThreads);
        if (k == 10)
            knn_warp_kernel<10><<<wgrid, kThreads, 0, st>>>(queries, m, (const float4 *)supports4, gid, b[0], b[1], b[2], b[3], b[4],
                                                             n_levels, query_sel, out);
        else
            knn_warp_kernel<3><<<wgrid, kThreads, 0, st>>>(queries, m, (const float4 *)supports4, gid, b[0], b[1], b[2], b[3], b[4],
                                                            n_levels, query_sel, out);
        OCC_LAUNCH_CHECK();
        return OCCNERF_OK;
    }
    const unsigned grid = occ_div_up(m, kThreads);
    if (k == 10)
        knn_kernel<10><<<grid, kThreads, 0, st>>>(queries, m, (const float4 *)supports4, gid, b[0], b[1], b[2], b[3], b[4],
                                                  n_levels, query_sel, out);
    else
        knn_kernel<3><<<grid, kThreads, 0, st>>>(queries, m, (const float4 *)supports4, gid, b[0], b[1], b[2], b[3], b[4],
                                                 n_levels, query_sel, out);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

}  // namespace

extern "C" int occnerf_knn(const float *queries, int m, const float *supports4, const int32_t *support_gid,
                           const int32_t *level_begin_host, int n_levels, int k, const uint8_t *query_sel,
                           int32_t *out_idx, occnerf_stream_t stream) {
    OCC_CHECK_ARG(queries && supports4 && level_begin_host && out_idx, "knn: null pointer");
    OCC_CHECK_ARG(n_levels >= 1 && n_levels <= 4, "knn: n_levels=%d outside [1,4]", n_levels);
    OCC_CHECK_ARG(k == 3 || k == 10, "knn: k=%d (supported: 3, 10)", k);
    OCC_CHECK_ARG(((uintptr_t)supports4 & 15) == 0, "knn: supports4 must be 16-byte aligned");
    for (int i = 0; i < n_levels; ++i)
        OCC_CHECK_ARG(level_begin_host[i] <= level_begin_host[i + 1], "knn: level_begin not ascending");
    if (m <= 0) return OCCNERF_OK;
    return launch_knn(queries, m, supports4, support_gid, level_begin_host, n_levels, k, query_sel, out_idx,
                      (cudaStream_t)stream);
}

extern "C" int occnerf_sample_geometry(const float *xyz, const int32_t *knn_idx, int knn_stride, const float *point_base,
                                       const float *point_norms, float bound, int m, float *enc_in, float *dist,
                                       int dist_stride, occnerf_stream_t stream) {
    OCC_CHECK_ARG(xyz && knn_idx && point_base && point_norms && enc_in && dist, "sample_geometry: null pointer");
    OCC_CHECK_ARG(knn_stride >= 10 && bound > 0.f && dist_stride >= 1, "sample_geometry: knn_stride=%d bound=%f", knn_stride, bound);
    OCC_CHECK_ARG(((uintptr_t)enc_in & 15) == 0, "sample_geometry: enc_in must be 16-byte aligned");
    if (m <= 0) return OCCNERF_OK;
    sample_geometry_kernel<<<occ_div_up(m, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        xyz, knn_idx, knn_stride, point_base, point_norms, bound, m, enc_in, dist, dist_stride);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_knn_hier(const float *queries, int m, int group_stride, const float *fine4, const float *centers4,
                                const int32_t *cluster_ranges, int nf, int nc, const int32_t *fine_gid,
                                const int32_t *center_gid, int k, int32_t *out_fine, int32_t *out_center, int out_stride,
                                occnerf_stream_t stream) {
    if (m <= 0) return OCCNERF_OK;
    OCC_CHECK_ARG(queries && fine4 && centers4 && cluster_ranges && out_fine && out_center, "knn_hier: null pointer");
    OCC_CHECK_ARG(k == 10, "knn_hier: k=%d (supported: 10)", k);
    OCC_CHECK_ARG(group_stride >= 1 && nf >= 1 && nc >= 1 && out_stride >= k, "knn_hier: bad sizes");
    OCC_CHECK_ARG((((uintptr_t)fine4 | (uintptr_t)centers4) & 15) == 0 && ((uintptr_t)cluster_ranges & 7) == 0,
                  "knn_hier: fine4/centers4 must be 16-byte and cluster_ranges 8-byte aligned");
    const size_t smem = (size_t)(nf + nc) * 16 + (size_t)nc * 8;
    OCC_CHECK_ARG(smem <= 220 * 1024, "knn_hier: %d + %d support points do not fit in shared memory", nf, nc);
    OCC_CUDA(cudaFuncSetAttribute(knn_hier_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long rays = (m + group_stride - 1) / group_stride;
    const long warps = ((rays + 31) / 32) * group_stride;
    const unsigned grid = occ_div_up(warps * 32, kHierThreads);
    knn_hier_kernel<10><<<grid, kHierThreads, smem, (cudaStream_t)stream>>>(
        queries, m, group_stride, (const float4 *)fine4, (const float4 *)centers4, (const int2 *)cluster_ranges, nf, nc,
        fine_gid, center_gid, out_fine, out_center, out_stride);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_knn_tree(const float *queries, int m, int group_stride, int lane_rays, const float *p0s, const float *p1s,
                                const float *p2s, const float *p3, const float *c2tab, const float *c3tab, const int32_t *c3rng,
                                const int32_t *gid1, const int32_t *gid2, const int32_t *gid3, const int32_t *inv2, int n0,
                                int n1, int n2, int n3, int k, int32_t *out, occnerf_stream_t stream) {
    if (m <= 0) return OCCNERF_OK;
    OCC_CHECK_ARG(queries && p0s && p1s && p2s && p3 && c2tab && c3tab && c3rng && gid1 && gid2 && gid3 && inv2 && out,
                  "knn_tree: null pointer");
    OCC_CHECK_ARG(k == 10, "knn_tree: k=%d (supported: 10)", k);
    OCC_CHECK_ARG(group_stride >= 1 && n0 >= 1 && n1 >= 1 && n2 >= 1 && n3 >= 1, "knn_tree: bad sizes");
    OCC_CHECK_ARG(lane_rays == 1 || lane_rays == 2 || lane_rays == 4 || lane_rays == 8 || lane_rays == 16 || lane_rays == 32,
                  "knn_tree: lane_rays=%d (a power of two <= 32)", lane_rays);
    OCC_CHECK_ARG((((uintptr_t)p0s | (uintptr_t)p1s | (uintptr_t)p2s | (uintptr_t)p3 | (uintptr_t)c2tab | (uintptr_t)c3tab |
                    (uintptr_t)c3rng) & 15) == 0, "knn_tree: tables must be 16-byte aligned");
    const size_t smem = (size_t)(n0 + n1 + 2 * n2 + 3 * n3) * 16;
    OCC_CHECK_ARG(smem <= 220 * 1024, "knn_tree: %d support points do not fit in shared memory", n0 + n1 + n2 + n3);
    static size_t configured = 0;
    if (configured < smem) {
        OCC_CUDA(cudaFuncSetAttribute(knn_tree_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    TreeArgs T;
    T.p0s = (const float4 *)p0s; T.p1s = (const float4 *)p1s; T.p2s = (const float4 *)p2s; T.p3 = (const float4 *)p3;
    T.c2tab = (const float4 *)c2tab; T.c3tab = (const float4 *)c3tab; T.c3rng = (const int4 *)c3rng;
    T.gid1 = gid1; T.gid2 = gid2; T.gid3 = gid3; T.inv2 = inv2;
    T.n0 = n0; T.n1 = n1; T.n2 = n2; T.n3 = n3;
    const long rays = (m + group_stride - 1) / group_stride;
    const int lane_samples = 32 / lane_rays;
    const long warps = ((rays + lane_rays - 1) / lane_rays) * ((group_stride + lane_samples - 1) / lane_samples);
    knn_tree_kernel<10><<<occ_div_up(warps * 32, kHierThreads), kHierThreads, smem, (cudaStream_t)stream>>>(queries, m, group_stride,
                                                                                                      lane_rays, T, out);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_visibility_hits(const float *depth, const int64_t *term, const float *x_skel, int N, int S,
                                       float thresh, const float *cloud4, int V, int k, float *hits, void *scratch,
                                       occnerf_stream_t stream) {
    OCC_CHECK_ARG(depth && term && x_skel && cloud4 && hits && scratch, "visibility_hits: null pointer");
    OCC_CHECK_ARG(k == 3 || k == 10, "visibility_hits: k=%d", k);
    cudaStream_t st = (cudaStream_t)stream;
    OCC_CUDA(cudaMemsetAsync(hits, 0, sizeof(float) * (size_t)V, st));
    if (N <= 0) return OCCNERF_OK;
    int *count = (int *)scratch;
    uint8_t *sel = (uint8_t *)scratch + 16;
    float *qpts = (float *)((uint8_t *)scratch + 16 + (size_t)4 * N);
    int32_t *idx = (int32_t *)(qpts + (size_t)3 * N);
    OCC_CUDA(cudaMemsetAsync(count, 0, 16, st));
    vis_select_kernel<<<occ_div_up(N, 256), 256, 0, st>>>(depth, term, x_skel, N, S, thresh, count, sel, qpts);
    OCC_LAUNCH_CHECK();
    const int32_t lb[2] = {0, V};
    if (int e = launch_knn(qpts, N, cloud4, nullptr, lb, 1, k, sel, idx, st)) return e;
    vis_mark_kernel<<<occ_div_up(N, 256), 256, 0, st>>>(count, sel, idx, N, k, hits);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_knn_grid(const float *queries, int m, int group_stride, int lane_rays, const float *p0, const float *p1,
                                const float *p2, const float *p3, int n0, int n1, int n2, int n3, const int32_t *gid1,
                                const int32_t *gid2, const int32_t *gid3, const int32_t *cell_tab, const uint16_t *lists,
                                const float *grid_min_invh_host, const int32_t *grid_dims_host, int k, int32_t *out,
                                occnerf_stream_t stream) {
    if (m <= 0) return OCCNERF_OK;
    OCC_CHECK_ARG(queries && p0 && p1 && p2 && p3 && gid1 && gid2 && gid3 && cell_tab && lists && grid_min_invh_host &&
                  grid_dims_host && out, "knn_grid: null pointer");
    OCC_CHECK_ARG(k == 10, "knn_grid: k=%d (supported: 10)", k);
    OCC_CHECK_ARG(group_stride >= 1 && n0 >= 1 && n1 >= 1 && n2 >= 1 && n3 >= 1 && n0 <= 65535, "knn_grid: bad sizes");
    OCC_CHECK_ARG(lane_rays == 1 || lane_rays == 2 || lane_rays == 4 || lane_rays == 8 || lane_rays == 16 || lane_rays == 32,
                  "knn_grid: lane_rays=%d (a power of two <= 32)", lane_rays);
    OCC_CHECK_ARG((((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)p3) & 15) == 0 && ((uintptr_t)cell_tab & 7) == 0,
                  "knn_grid: point arrays must be 16-byte and cell_tab 8-byte aligned");
    OCC_CHECK_ARG(grid_dims_host[0] >= 1 && grid_dims_host[1] >= 1 && grid_dims_host[2] >= 1 && grid_min_invh_host[3] > 0.f,
                  "knn_grid: bad grid");
    const size_t smem = (size_t)(n0 + n1 + n2 + n3) * 16;
    OCC_CHECK_ARG(smem <= 227 * 1024, "knn_grid: %d support points do not fit in shared memory", n0 + n1 + n2 + n3);
    static size_t configured = 0;
    if (configured < smem) {
        OCC_CUDA(cudaFuncSetAttribute(knn_grid_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    GridArgs G;
    G.p[0] = (const float4 *)p0; G.p[1] = (const float4 *)p1; G.p[2] = (const float4 *)p2; G.p[3] = (const float4 *)p3;
    G.n[0] = n0; G.n[1] = n1; G.n[2] = n2; G.n[3] = n3;
    G.gid[0] = nullptr; G.gid[1] = gid1; G.gid[2] = gid2; G.gid[3] = gid3;
    G.cell_tab = (const int2 *)cell_tab;
    G.lists = lists;
    for (int a = 0; a < 3; ++a) { G.gmin[a] = grid_min_invh_host[a]; G.dims[a] = grid_dims_host[a]; }
    G.inv_h = grid_min_invh_host[3];
    const long rays = (m + group_stride - 1) / group_stride;
    const int lane_samples = 32 / lane_rays;
    const long warps = ((rays + lane_rays - 1) / lane_rays) * ((group_stride + lane_samples - 1) / lane_samples);
    knn_grid_kernel<10><<<occ_div_up(warps * 32, kHierThreads), kHierThreads, smem, (cudaStream_t)stream>>>(queries, m, group_stride,
                                                                                                      lane_rays, G, out);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
