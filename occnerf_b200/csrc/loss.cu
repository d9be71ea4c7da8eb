// Loss epilogue of the training step without the LPIPS term:
// core/train/trainers/occnerf/trainer.py:31-41 (_unpack_imgs: scatter of the rendered rays into N_patch background-filled P x P images),
// :24 / :96-97 (img2mse over those images) and :172-189 (get_loss: + mean(comp_loss), weighted sum), forward AND gradients in one call:
//   patch_loss_kernel  block = patch: in-patch ranks of the hit pixels (ballot scan) -> ray of each pixel; writes the patch image,
//                      accumulates the squared error (double) and writes d loss / d rgb for its rays
//   comp_sum_kernel    sum of comp_loss (double)
//   loss_final_kernel  loss = w_mse * sse / (N P P 3) + w_comp * csum / numel; the constant d loss / d comp_loss
// The perceptual term (third_parties/lpips, VGG weights that are not in this image) consumes `patch_imgs` in the host framework.
#include "common.cuh"

namespace {

__device__ __forceinline__ int block_rank1024(bool flag, int *total, int *s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(OCC_FULL, flag);
    const int in_warp = __popc(b & ((1u << lane) - 1u));
    __syncthreads();
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    if (warp == 0) {
        int v = s_warp[lane], x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(OCC_FULL, x, o);
            if (lane >= o) x += y;
        }
        s_warp[lane] = x - v;
        if (lane == 31) s_warp[32] = x;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + in_warp;
}

__global__ void __launch_bounds__(1024) patch_loss_kernel(const float *__restrict__ rgb, const uint8_t *__restrict__ masks,
                                                          const int *__restrict__ div, const float *__restrict__ bg,
                                                          const float *__restrict__ targets, int P, float g_scale, float *__restrict__ imgs,
                                                          double *__restrict__ acc, float *__restrict__ g_rgb) {
    __shared__ int s_warp[33];
    __shared__ double s_sum[32];
    const int k = blockIdx.x, PP = P * P;
    const int first = div[k], last = div[k + 1];
    const float b0 = __ldg(bg), b1 = __ldg(bg + 1), b2 = __ldg(bg + 2);
    double sse = 0.0;
    int base = 0;
    for (int i0 = 0; i0 < PP; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool in = i < PP;
        const bool hit = in && masks[(long)k * PP + i] != 0;
        int t;
        const int r = first + base + block_rank1024(hit, &t, s_warp);
        base += t;
        if (in) {
            const long o = ((long)k * PP + i) * 3;
            float c0 = b0, c1 = b1, c2 = b2;
            const bool ok = hit && r < last;                  // (a mask with more hits than div allots would be a caller error)
            if (ok) { c0 = __ldg(rgb + (long)r * 3); c1 = __ldg(rgb + (long)r * 3 + 1); c2 = __ldg(rgb + (long)r * 3 + 2); }
            const float d0 = c0 - __ldg(targets + o), d1 = c1 - __ldg(targets + o + 1), d2 = c2 - __ldg(targets + o + 2);
            if (imgs) { imgs[o] = c0; imgs[o + 1] = c1; imgs[o + 2] = c2; }
            sse += (double)(d0 * d0) + (double)(d1 * d1) + (double)(d2 * d2);
            if (ok) { g_rgb[(long)r * 3] = g_scale * d0; g_rgb[(long)r * 3 + 1] = g_scale * d1; g_rgb[(long)r * 3 + 2] = g_scale * d2; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(OCC_FULL, sse, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sse;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += s_sum[w];
        atomicAdd(acc, t);
    }
}

__global__ void __launch_bounds__(256) comp_sum_kernel(const float *__restrict__ comp, long n, double *__restrict__ acc) {
    double s = 0.0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) s += (double)__ldg(comp + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(OCC_FULL, s, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += part[w];
        atomicAdd(acc + 1, t);
    }
}

__global__ void loss_final_kernel(const double *__restrict__ acc, double img_numel, double comp_numel, float w_mse, float w_comp,
                                  float *__restrict__ out) {
    const double mse = acc[0] / img_numel, comp = comp_numel > 0 ? acc[1] / comp_numel : 0.0;
    out[1] = (float)(w_mse * mse);
    out[2] = (float)(w_comp * comp);
    out[0] = (float)(w_mse * mse + w_comp * comp);
    out[3] = comp_numel > 0 ? (float)(w_comp / comp_numel) : 0.f;       // d loss / d comp_loss[i]
}

}  // namespace

// rgb [n, 3] (n = div[N]); masks [N, P, P] bytes; div [N + 1] i32; bgcolor [3] (already divided by 255 as the trainer passes it);
// targets [N, P, P, 3]; comp [comp_numel] or NULL.  Outputs: imgs [N, P, P, 3] or NULL (the unpacked patch images), out [4] =
// (loss, w_mse * mse, w_comp * mean(comp), d loss / d comp_i), g_rgb [n, 3] = d loss / d rgb; acc: 2 doubles of scratch.
extern "C" int occnerf_patch_loss(const float *rgb, const uint8_t *masks, const int32_t *div, const float *bgcolor, const float *targets,
                                  const float *comp, long comp_numel, int N, int P, float w_mse, float w_comp, float *imgs, float *out,
                                  float *g_rgb, void *acc, occnerf_stream_t stream) {
    OCC_CHECK_ARG(rgb && masks && div && bgcolor && targets && out && g_rgb && acc, "patch_loss: null pointer");
    OCC_CHECK_ARG(N >= 1 && P >= 1 && comp_numel >= 0 && (comp || comp_numel == 0), "patch_loss: N=%d P=%d comp_numel=%ld", N, P, comp_numel);
    cudaStream_t st = (cudaStream_t)stream;
    OCC_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
    const double img_numel = (double)N * P * P * 3;
    patch_loss_kernel<<<N, 1024, 0, st>>>(rgb, masks, div, bgcolor, targets, P, (float)(2.0 * w_mse / img_numel), imgs, (double *)acc, g_rgb);
    OCC_LAUNCH_CHECK();
    if (comp_numel > 0) {
        comp_sum_kernel<<<(unsigned)min((comp_numel + 255) / 256, 1184L), 256, 0, st>>>(comp, comp_numel, (double *)acc);
        OCC_LAUNCH_CHECK();
    }
    loss_final_kernel<<<1, 1, 0, st>>>((const double *)acc, img_numel, (double)comp_numel, w_mse, w_comp, out);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
