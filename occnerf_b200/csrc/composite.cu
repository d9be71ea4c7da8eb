// K4: alpha compositing as a warp-shuffle transmittance scan, backward recomputed (nothing stored).
//
// Replaces Network._raw2outputs (core/nets/occnerf/network.py:320-348: ~20 pointwise kernels and a
// torch.cumprod over a concatenated [N,S+1] tensor) and the training-only completeness term
// (network.py:486-499).  One warp owns one ray; lane l handles samples l, l+32, l+64, ... so that the
// mask / z loads are fully coalesced and the exclusive transmittance product is one 5-step
// __shfl_up_sync scan per 32-sample chunk with a scalar carry between chunks.
//
// Backward: pass 1 recomputes alpha_j and T_j into registers; pass 2 walks the chunks in reverse with a
// __shfl_down_sync suffix scan of c_k*w_k.  The suffix is summed directly (never as total - prefix), because
// it is divided by (1 - alpha_j + 1e-10), which can be 1e-10.
//
// HBM traffic (algorithmic): forward 28 B/sample in + 28 B/ray out; backward 28 B/sample in, 24 B/sample out.
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxChunks = 8;   // S <= 256

__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

struct SampleIn {
    float r[5];
    float mask, z, delta;
};

__device__ __forceinline__ SampleIn load_sample(const float *__restrict__ raw, const float *__restrict__ mask,
                                                const float *__restrict__ z, long base, int j, int S, float dnorm) {
    SampleIn s;
    const float *rp = raw + (base + j) * 5;
#pragma unroll
    for (int c = 0; c < 5; ++c) s.r[c] = __ldg(rp + c);
    s.mask = __ldg(mask + base + j);
    s.z = __ldg(z + base + j);
    const float dz = (j + 1 < S) ? __fsub_rn(__ldg(z + base + j + 1), s.z) : 1e10f;
    s.delta = __fmul_rn(dz, dnorm);
    return s;
}

__device__ __forceinline__ float ray_dnorm(const float *__restrict__ rays, long ray) {
    const float dx = __ldg(rays + ray * 8 + 3), dy = __ldg(rays + ray * 8 + 4), dz = __ldg(rays + ray * 8 + 5);
    return sqrtf(dx * dx + dy * dy + dz * dz);
}

// completeness term, network.py:486-499
__device__ __forceinline__ float comp_term(float sigma, float dist) {
    if (!(dist < 0.0f)) return 0.0f;
    const float s = dist > 0.3f ? 0.0f : sigma;
    return 10.0f * expf(fminf(fmaxf(-fmaxf(s, 0.0f), -10.0f), 0.0f));
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(const float *__restrict__ raw, const float *__restrict__ mask, const float *__restrict__ z,
                     const float *__restrict__ rays, const float *__restrict__ bg, int N, int S,
                     float *__restrict__ rgb_out, float *__restrict__ acc_out, float *__restrict__ depth_out,
                     int64_t *__restrict__ term_out, float *__restrict__ weights_out, float *__restrict__ comp_out) {
    const int lane = threadIdx.x & 31;
    const long ray = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (ray >= N) return;
    const long base = ray * S;
    const float dnorm = ray_dnorm(rays, ray);
    float carryT = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f, depth = 0.f;
    float best_a = -INFINITY;
    int best_j = 0x7fffffff;
    const int nchunks = (S + 31) >> 5;
    for (int c = 0; c < nchunks; ++c) {
        const int j = c * 32 + lane;
        const bool valid = j < S;
        float alpha = 0.f, zz = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (valid) {
            const SampleIn s = load_sample(raw, mask, z, base, j, S, dnorm);
            const float sp = softplus_f(s.r[3]);
            alpha = __fmul_rn(__fsub_rn(1.0f, expf(__fmul_rn(-sp, s.delta))), s.mask);
            zz = s.z;
            r0 = sigmoid_f(s.r[0]); r1 = sigmoid_f(s.r[1]); r2 = sigmoid_f(s.r[2]);
            if (comp_out) comp_out[base + j] = comp_term(s.r[3], s.r[4]);
            if (alpha > best_a) { best_a = alpha; best_j = j; }
        }
        const float f = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
        float P = f;   // inclusive product scan over the 32 lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float up = __shfl_up_sync(OCC_FULL, P, o);
            if (lane >= o) P *= up;
        }
        float E = __shfl_up_sync(OCC_FULL, P, 1);
        if (lane == 0) E = 1.0f;
        const float T = carryT * E;
        const float w = alpha * T;
        carryT *= __shfl_sync(OCC_FULL, P, 31);
        if (valid && weights_out) weights_out[base + j] = w;
        cr += w * r0; cg += w * r1; cb += w * r2;
        acc += w; depth += w * zz;
    }
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
    acc = warp_sum(acc); depth = warp_sum(depth);
    // argmax(alpha), first maximum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float oa = __shfl_xor_sync(OCC_FULL, best_a, o);
        const int oj = __shfl_xor_sync(OCC_FULL, best_j, o);
        if (oa > best_a || (oa == best_a && oj < best_j)) { best_a = oa; best_j = oj; }
    }
    if (lane == 0) {
        const float rest = 1.0f - acc;
        rgb_out[ray * 3 + 0] = cr + rest * __ldg(bg + 0) / 255.0f;
        rgb_out[ray * 3 + 1] = cg + rest * __ldg(bg + 1) / 255.0f;
        rgb_out[ray * 3 + 2] = cb + rest * __ldg(bg + 2) / 255.0f;
        acc_out[ray] = acc;
        depth_out[ray] = depth;
        term_out[ray] = best_j == 0x7fffffff ? 0 : best_j;
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(const float *__restrict__ raw, const float *__restrict__ mask, const float *__restrict__ z,
                     const float *__restrict__ rays, const float *__restrict__ bg, const float *__restrict__ g_rgb,
                     const float *__restrict__ g_acc, const float *__restrict__ g_depth,
                     const float *__restrict__ g_comp, int N, int S, float *__restrict__ g_raw,
                     float *__restrict__ g_mask) {
    const int lane = threadIdx.x & 31;
    const long ray = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (ray >= N) return;
    const long base = ray * S;
    const float dnorm = ray_dnorm(rays, ray);
    const int nchunks = (S + 31) >> 5;
    float alpha_s[kMaxChunks], T_s[kMaxChunks];
    // pass 1: alpha and transmittance
    float carryT = 1.0f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
        if (c < nchunks) {
            const int j = c * 32 + lane;
            const bool valid = j < S;
            float alpha = 0.f;
            if (valid) {
                const SampleIn s = load_sample(raw, mask, z, base, j, S, dnorm);
                alpha = __fmul_rn(__fsub_rn(1.0f, expf(__fmul_rn(-softplus_f(s.r[3]), s.delta))), s.mask);
            }
            const float f = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
            float P = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float up = __shfl_up_sync(OCC_FULL, P, o);
                if (lane >= o) P *= up;
            }
            float E = __shfl_up_sync(OCC_FULL, P, 1);
            if (lane == 0) E = 1.0f;
            alpha_s[c] = alpha;
            T_s[c] = carryT * E;
            carryT *= __shfl_sync(OCC_FULL, P, 31);
        }
    }
    const float gr = __ldg(g_rgb + ray * 3 + 0), gg = __ldg(g_rgb + ray * 3 + 1), gb = __ldg(g_rgb + ray * 3 + 2);
    const float ga = __ldg(g_acc + ray), gd = __ldg(g_depth + ray);
    const float b0 = __ldg(bg + 0) / 255.0f, b1 = __ldg(bg + 1) / 255.0f, b2 = __ldg(bg + 2) / 255.0f;
    // pass 2: reverse suffix scan
    float carryR = 0.f;
#pragma unroll
    for (int c = kMaxChunks - 1; c >= 0; --c) {
        if (c < nchunks) {
            const int j = c * 32 + lane;
            const bool valid = j < S;
            SampleIn s;
            float alpha = 0.f, T = 0.f, w = 0.f, cj = 0.f, e = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
            if (valid) {
                s = load_sample(raw, mask, z, base, j, S, dnorm);
                alpha = alpha_s[c];
                T = T_s[c];
                w = alpha * T;
                e = expf(__fmul_rn(-softplus_f(s.r[3]), s.delta));
                r0 = sigmoid_f(s.r[0]); r1 = sigmoid_f(s.r[1]); r2 = sigmoid_f(s.r[2]);
                cj = gr * (r0 - b0) + gg * (r1 - b1) + gb * (r2 - b2) + ga + gd * s.z;
            }
            const float term = cj * w;
            float Sfx = term;   // inclusive suffix sum over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float dn = __shfl_down_sync(OCC_FULL, Sfx, o);
                if (lane + o < 32) Sfx += dn;
            }
            float R = __shfl_down_sync(OCC_FULL, Sfx, 1);
            if (lane == 31) R = 0.f;
            R += carryR;
            carryR += __shfl_sync(OCC_FULL, Sfx, 0);
            if (valid) {
                const float dalpha = cj * T - R / __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
                const float dsp = s.r[3] > 20.0f ? 1.0f : sigmoid_f(s.r[3]);
                float gs = dalpha * s.mask * (e * s.delta) * dsp;
                if (g_comp) {
                    const float sg = s.r[3], dist = s.r[4];
                    if (dist < 0.0f && sg > 0.0f && sg <= 10.0f) gs -= __ldg(g_comp + base + j) * comp_term(sg, dist);
                }
                float *o = g_raw + (base + j) * 5;
                o[0] = gr * w * r0 * (1.0f - r0);
                o[1] = gg * w * r1 * (1.0f - r1);
                o[2] = gb * w * r2 * (1.0f - r2);
                o[3] = gs;
                o[4] = 0.0f;
                g_mask[base + j] = dalpha * (1.0f - e);
            }
        }
    }
}

}  // namespace

extern "C" int occnerf_composite_forward(const float *raw, const float *mask, const float *z, const float *rays,
                                         const float *bg, int N, int S, float *rgb, float *acc, float *depth,
                                         int64_t *term, float *weights, float *comp, occnerf_stream_t stream) {
    if (N == 0) return OCCNERF_OK;
    OCC_CHECK_ARG(raw && mask && z && rays && bg && rgb && acc && depth && term, "composite_forward: null pointer");
    OCC_CHECK_ARG(N >= 0 && S >= 1, "composite_forward: bad N=%d S=%d", N, S);
    if (N == 0) return OCCNERF_OK;
    composite_fwd_kernel<<<occ_div_up(N, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        raw, mask, z, rays, bg, N, S, rgb, acc, depth, term, weights, comp);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_composite_backward(const float *raw, const float *mask, const float *z, const float *rays,
                                          const float *bg, const float *g_rgb, const float *g_acc,
                                          const float *g_depth, const float *g_comp, int N, int S, float *g_raw,
                                          float *g_mask, occnerf_stream_t stream) {
    if (N == 0) return OCCNERF_OK;
    OCC_CHECK_ARG(raw && mask && z && rays && bg && g_rgb && g_acc && g_depth && g_raw && g_mask,
                  "composite_backward: null pointer");
    OCC_CHECK_ARG(N >= 0 && S >= 1 && S <= 32 * kMaxChunks, "composite_backward: S=%d outside [1,%d]", S,
                  32 * kMaxChunks);
    if (N == 0) return OCCNERF_OK;
    composite_bwd_kernel<<<occ_div_up(N, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        raw, mask, z, rays, bg, g_rgb, g_acc, g_depth, g_comp, N, S, g_raw, g_mask);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
