// Per-vertex block of the canonical query (core/nets/occnerf/network.py:263-284 + canonical_mlps/occnerf_mlp.py:171-175):
// every vertex of the learnable cloud  pc = point_base + point_dist  is projected onto the base cloud through its 3
// nearest base vertices (|cos|-weighted mean), gets a signed mean distance (inside vote of the 3 normals) and from both
// the 4-D hash-grid input of the vertex.  V = 6890 points once per call -- tiny, but in the reference (and in a first
// version here) it was ~25 eager kernels forward and ~40 backward; here it is one kernel each way, with the analytic
// gradient to point_dist (the only trainable input: point_dist is (V,1), so d pc / d point_dist = (1,1,1)).
#include "common.cuh"

namespace {

struct VertexGeom {
    float d[3][3];      // pc - b_j
    float n[3][3];      // normals of the 3 neighbours
    float b[3][3];
    float dn[3], nd[3], c[3], a[3];
    float asum, kb[3], du;
    bool inside;
};

__device__ __forceinline__ void vertex_geometry(const float *__restrict__ base, const float *__restrict__ norms,
                                                const int32_t *__restrict__ kidx, float px, float py, float pz, VertexGeom &g) {
    int votes = 0;
    g.asum = 0.f;
    float sx = 0.f, sy = 0.f, sz = 0.f, dsum = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int v = __ldg(kidx + j);
#pragma unroll
        for (int a = 0; a < 3; ++a) { g.b[j][a] = __ldg(base + (size_t)v * 3 + a); g.n[j][a] = __ldg(norms + (size_t)v * 3 + a); }
        g.d[j][0] = px - g.b[j][0]; g.d[j][1] = py - g.b[j][1]; g.d[j][2] = pz - g.b[j][2];
        const float dot = g.d[j][0] * g.n[j][0] + g.d[j][1] * g.n[j][1] + g.d[j][2] * g.n[j][2];
        votes += dot < 0.f;
        g.dn[j] = sqrtf(g.d[j][0] * g.d[j][0] + g.d[j][1] * g.d[j][1] + g.d[j][2] * g.d[j][2]);
        const float nn = sqrtf(g.n[j][0] * g.n[j][0] + g.n[j][1] * g.n[j][1] + g.n[j][2] * g.n[j][2]);
        const float dd = fmaxf(g.dn[j], 1e-8f);
        g.nd[j] = fmaxf(nn, 1e-8f);
        // cosine_similarity with eps = 1e-8 on each norm
        g.c[j] = (g.d[j][0] / dd) * (g.n[j][0] / g.nd[j]) + (g.d[j][1] / dd) * (g.n[j][1] / g.nd[j]) + (g.d[j][2] / dd) * (g.n[j][2] / g.nd[j]);
        g.a[j] = fabsf(g.c[j]);
        g.asum += g.a[j];
        sx += g.a[j] * g.b[j][0]; sy += g.a[j] * g.b[j][1]; sz += g.a[j] * g.b[j][2];
        dsum += g.dn[j];
    }
    g.kb[0] = sx / g.asum; g.kb[1] = sy / g.asum; g.kb[2] = sz / g.asum;
    g.du = dsum / 3.0f;
    g.inside = votes > 1;                           // (sum > 1.5)
}

__global__ void vertex_fwd_kernel(const float *__restrict__ base, const float *__restrict__ dist, const float *__restrict__ norms,
                                  const int32_t *__restrict__ kidx, float bound, int V, float *__restrict__ v_in,
                                  float *__restrict__ feats_tail, int ld) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float pd = __ldg(dist + v);
    const float px = __ldg(base + (size_t)v * 3) + pd, py = __ldg(base + (size_t)v * 3 + 1) + pd, pz = __ldg(base + (size_t)v * 3 + 2) + pd;
    VertexGeom g;
    vertex_geometry(base, norms, kidx + (size_t)v * 3, px, py, pz, g);
    const float sd = g.inside ? -g.du : g.du;
    const float two_b = 2.0f * bound;
    reinterpret_cast<float4 *>(v_in)[v] = make_float4((g.kb[0] + bound) / two_b, (g.kb[1] + bound) / two_b, (g.kb[2] + bound) / two_b,
                                                      fminf(fmaxf((sd + 0.2f) / 0.8f, 0.0f), 1.0f));
    float *t = feats_tail + (size_t)v * ld;
    t[0] = px; t[1] = py; t[2] = pz; t[3] = 0.f;
}

// Backward in DOUBLE precision.  For the vertex's own base point |d_0| = |point_dist| sqrt(3) ~ 1e-4 and d_0 is (up to the
// rounding of pc = fl(base + dist)) parallel to (1,1,1), the direction in which point_dist moves pc: the three components of
// d loss / d pc are of order 1e3..1e4 and cancel to ~1e-4 of their size in the sum that is d loss / d point_dist (measured on the
// golden case: 3893.3 - 3048.7 - 844.2 = 0.39).  An fp32 evaluation of this closed form leaves 1e-4-relative errors in the
// components (cos rounding amplified by 1 / |d_0|), i.e. O(1) absolute errors in such sums -- the 1.9e-2 normwise outliers
// round 1 reported against the reference's gradient.  V = 6890 threads once per step: fp64 costs nothing here.
// The forward values (v_in) stay fp32: they select hash-grid cells and must follow the reference's arithmetic.
__global__ void vertex_bwd_kernel(const float *__restrict__ base, const float *__restrict__ dist, const float *__restrict__ norms,
                                  const int32_t *__restrict__ kidx, float bound, int V, const float *__restrict__ g_v_in,
                                  const float *__restrict__ g_tail, int ld, float *__restrict__ g_dist) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float pd = __ldg(dist + v);
    const float pf[3] = {__ldg(base + (size_t)v * 3) + pd, __ldg(base + (size_t)v * 3 + 1) + pd, __ldg(base + (size_t)v * 3 + 2) + pd};
    VertexGeom gf;
    vertex_geometry(base, norms, kidx + (size_t)v * 3, pf[0], pf[1], pf[2], gf);       // fp32 forward: inside vote, clamp decision
    const float sd = gf.inside ? -gf.du : gf.du;
    const float u = (sd + 0.2f) / 0.8f;
    const float4 gv = __ldg(reinterpret_cast<const float4 *>(g_v_in) + v);
    const double two_b = 2.0 * (double)bound;
    const double gkb[3] = {gv.x / two_b, gv.y / two_b, gv.z / two_b};
    double gsd = (u >= 0.0f && u <= 1.0f) ? (double)gv.w / 0.8 : 0.0;     // clamp passes the gradient on [min, max]
    if (gf.inside) gsd = -gsd;                                           // now d loss / d (mean |d_j|)
    // geometry again in double from the same fp32 inputs
    double d[3][3], n[3][3], b[3][3], dn[3], nd[3], c[3], a[3], asum = 0.0, kb[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int w = __ldg(kidx + (size_t)v * 3 + j);
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            b[j][x] = (double)__ldg(base + (size_t)w * 3 + x);
            n[j][x] = (double)__ldg(norms + (size_t)w * 3 + x);
            d[j][x] = (double)pf[x] - b[j][x];
        }
        dn[j] = sqrt(d[j][0] * d[j][0] + d[j][1] * d[j][1] + d[j][2] * d[j][2]);
        nd[j] = fmax(sqrt(n[j][0] * n[j][0] + n[j][1] * n[j][1] + n[j][2] * n[j][2]), 1e-8);
        const double dd = fmax(dn[j], 1e-8);
        c[j] = (d[j][0] * n[j][0] + d[j][1] * n[j][1] + d[j][2] * n[j][2]) / (dd * nd[j]);
        a[j] = fabs(c[j]);
        asum += a[j];
#pragma unroll
        for (int x = 0; x < 3; ++x) kb[x] += a[j] * b[j][x];
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) kb[x] /= asum;
    double gp[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        // kb = sum a_j b_j / A  ->  d kb / d a_j = (b_j - kb) / A
        const double ga = (gkb[0] * (b[j][0] - kb[0]) + gkb[1] * (b[j][1] - kb[1]) + gkb[2] * (b[j][2] - kb[2])) / asum;
        const double sgn = c[j] > 0.0 ? 1.0 : (c[j] < 0.0 ? -1.0 : 0.0);
        if (dn[j] > 1e-8) {
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const double dc = (n[j][x] / nd[j] - c[j] * d[j][x] / dn[j]) / dn[j];          // d cos / d d_x
                gp[x] += ga * sgn * dc + gsd * d[j][x] / (3.0 * dn[j]);
            }
        }
    }
    const float *t = g_tail + (size_t)v * ld;
    g_dist[v] = (float)((gp[0] + (double)__ldg(t + 0)) + (gp[1] + (double)__ldg(t + 1)) + (gp[2] + (double)__ldg(t + 2)));
}

}  // namespace

extern "C" int occnerf_vertex_block_forward(const float *point_base, const float *point_dist, const float *point_norms,
                                            const int32_t *kidx3, float bound, int V, float *v_in, float *feats_tail, int ld,
                                            occnerf_stream_t stream) {
    OCC_CHECK_ARG(point_base && point_dist && point_norms && kidx3 && v_in && feats_tail, "vertex_block_forward: null pointer");
    OCC_CHECK_ARG(bound > 0.f && ld >= 4 && ((uintptr_t)v_in & 15) == 0, "vertex_block_forward: bound=%f ld=%d", bound, ld);
    if (V <= 0) return OCCNERF_OK;
    vertex_fwd_kernel<<<occ_div_up(V, 128), 128, 0, (cudaStream_t)stream>>>(point_base, point_dist, point_norms, kidx3, bound, V, v_in,
                                                                           feats_tail, ld);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_vertex_block_backward(const float *point_base, const float *point_dist, const float *point_norms,
                                             const int32_t *kidx3, float bound, int V, const float *g_v_in, const float *g_feats_tail,
                                             int ld, float *g_point_dist, occnerf_stream_t stream) {
    OCC_CHECK_ARG(point_base && point_dist && point_norms && kidx3 && g_v_in && g_feats_tail && g_point_dist,
                  "vertex_block_backward: null pointer");
    OCC_CHECK_ARG(bound > 0.f && ld >= 4 && ((uintptr_t)g_v_in & 15) == 0, "vertex_block_backward: bound=%f ld=%d", bound, ld);
    if (V <= 0) return OCCNERF_OK;
    vertex_bwd_kernel<<<occ_div_up(V, 128), 128, 0, (cudaStream_t)stream>>>(point_base, point_dist, point_norms, kidx3, bound, V, g_v_in,
                                                                           g_feats_tail, ld, g_point_dist);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
