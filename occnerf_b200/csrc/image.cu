// Image assembly behind the path: run.py:39-66 (unpack_alpha_map / unpack_to_image) + image_util.py:19-20,34-35
// (to_8b_image) on the device.  The reference copies rgb [n,3] and alpha [n] to the host, scatters them into a
// background-filled float32 frame through the boolean ray_mask and converts to 8 bit in numpy; here the frame is filled
// and scattered in 8 bit directly (4 B per pixel written instead of 16 B per valid ray read back + host work):
//   rgb8[p]   = uint8(255.f * clip(v, 0, 1))   v = rgb of the ray at pixel p, else bgcolor      (float32 product, truncation)
//   alpha8[p] = uint8(255.f * clip(a, 0, 1))   a = alpha of the ray at pixel p, else 0
// Rays are addressed through pixel_index (what occnerf_generate_rays returns), so one rank of a sharded render can
// scatter just its contiguous ray range; the fill pass is optional for the same reason.
#include "common.cuh"

__device__ __forceinline__ uint8_t to_8b(float v) {
    v = fminf(fmaxf(v, 0.0f), 1.0f);            // np.clip (NaN would propagate in numpy; fmaxf maps it to 0 -- not produced by the path)
    return (uint8_t)(int)__fmul_rn(255.0f, v);   // .astype(np.uint8): truncation
}

__global__ void image_fill_kernel(uint8_t *__restrict__ rgb8, uint8_t *__restrict__ alpha8, long P, uint8_t b0, uint8_t b1, uint8_t b2) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    rgb8[3 * p + 0] = b0;
    rgb8[3 * p + 1] = b1;
    rgb8[3 * p + 2] = b2;
    if (alpha8) alpha8[p] = 0;
}

__global__ void image_scatter_kernel(const float *__restrict__ rgb, const float *__restrict__ alpha, const int *__restrict__ pixel_index,
                                     int n, long P, uint8_t *__restrict__ rgb8, uint8_t *__restrict__ alpha8, int *__restrict__ bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const long p = pixel_index[r];
    if (p < 0 || p >= P) {
        atomicAdd(bad, 1);
        return;
    }
    rgb8[3 * p + 0] = to_8b(rgb[3 * (long)r + 0]);
    rgb8[3 * p + 1] = to_8b(rgb[3 * (long)r + 1]);
    rgb8[3 * p + 2] = to_8b(rgb[3 * (long)r + 2]);
    if (alpha8 && alpha) alpha8[p] = to_8b(alpha[r]);
}

static inline uint8_t host_to_8b(float v) {
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    return (uint8_t)(int)(255.0f * v);
}

extern "C" int occnerf_unpack_image(const float *rgb, const float *alpha, const int *pixel_index, int n, int H, int W,
                                    const float *bgcolor_host, int fill, uint8_t *rgb8, uint8_t *alpha8, int *bad,
                                    occnerf_stream_t stream) {
    OCC_CHECK_ARG(H > 0 && W > 0 && n >= 0, "occnerf_unpack_image: bad sizes n=%d H=%d W=%d", n, H, W);
    OCC_CHECK_ARG(rgb8 && bad, "occnerf_unpack_image: NULL rgb8 / bad");
    OCC_CHECK_ARG(n == 0 || (rgb && pixel_index), "occnerf_unpack_image: NULL rgb / pixel_index with n=%d", n);
    OCC_CHECK_ARG(!fill || bgcolor_host, "occnerf_unpack_image: fill requested without a background colour");
    cudaStream_t s = (cudaStream_t)stream;
    const long P = (long)H * W;
    if (fill) {
        image_fill_kernel<<<occ_div_up(P, 256), 256, 0, s>>>(rgb8, alpha8, P, host_to_8b(bgcolor_host[0]), host_to_8b(bgcolor_host[1]),
                                                            host_to_8b(bgcolor_host[2]));
        OCC_LAUNCH_CHECK();
    }
    if (n > 0) {
        image_scatter_kernel<<<occ_div_up(n, 256), 256, 0, s>>>(rgb, alpha, pixel_index, n, P, rgb8, alpha8, bad);
        OCC_LAUNCH_CHECK();
    }
    return OCCNERF_OK;
}
