// Ray generation + ray / bounding-box intersection + order-preserving compaction of the rays that hit the box
// (the step in front of the ray path; the reference runs it in numpy inside the DataLoader).
//
// Follows core/utils/camera_util.py:133-160 (get_rays_from_KRT) and :163-212 (rays_intersect_3d_bbox) and the
// masking that every dataset applies right after them (freeview.py:208-219, train.py:440-461):
//   pixel_camera = [i, j, 1] . inv(K)^T           (float32 when K is float32 -- ZJU pickles -- else float64)
//   rays_d       = (pixel_camera - T) . R - rays_o,   rays_o = -R^T T     (float64; float32 when K, R and T all are --
//                                                       tpose.py:66-84 --, then |d| is a float32 norm too)
//   rays_d[|rays_d| < 1e-5] = 1e-5                    (in place: the rays handed to the network carry it)
//   six plane hits of the box grown by 1 cm, kept when inside the box (+-1e-6); a ray is valid when exactly two are
//   near/far = min/max of |p - o| / |d|,   cast to float32 together with o, d
// The arithmetic is float64 with the roundings numpy makes (dot products as a k = 0,1,2 FMA chain like the BLAS
// micro-kernels, everything else un-fused), so the hit mask -- an integer decision -- matches the reference.
//
// Three launches, all HBM-trivial (33 B written per hit pixel, 1 B per pixel):
//   rays_count  : one block per 1024 pixels -> mask [H*W] + hits per block
//   rays_scan   : one block, exclusive scan of the block counts (+ total)
//   rays_emit   : recomputes the ray, ranks it inside its block (ballot + warp prefix) and writes [o3,d3,near,far]
#include "common.cuh"

struct RayCam {
    double kinv[9];   // inv(K), row major
    double R[9];      // row major
    double T[3];
    double o[3];      // -R^T T
    double lo[3], hi[3];   // box grown by 1 cm
    int H, W, k_f32, all_f32;
};

struct RayOut {
    double d[3];
    float near, far;
    bool hit;
};

#define RAYS_BLOCK 1024

__device__ __forceinline__ double dot3_chain(double a0, double a1, double a2, double b0, double b1, double b2) {
    double acc = __dmul_rn(a0, b0);
    acc = __fma_rn(a1, b1, acc);
    return __fma_rn(a2, b2, acc);
}
__device__ __forceinline__ float dot3_chain_f(float a0, float a1, float a2, float b0, float b1, float b2) {
    float acc = __fmul_rn(a0, b0);
    acc = __fmaf_rn(a1, b1, acc);
    return __fmaf_rn(a2, b2, acc);
}

// float32 length-3 dot as numpy's N-D np.dot evaluates it through OpenBLAS' sdot (the build behind the fixtures): k = 0,1 paired
// -- fma(a0 b0, a1 b1) -- then the odd k = 2 product added.  Equal to the chain whenever a1 b1 = 0 (a K without skew); any other
// BLAS differs by <= 1 float32 ulp.
__device__ __forceinline__ float dot3_sdot_f(float a0, float a1, float a2, float b0, float b1, float b2) {
    return __fadd_rn(__fmaf_rn(a0, b0, __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

__device__ __forceinline__ RayOut ray_for_pixel(const RayCam &c, int pix) {
    RayOut r;
    const int j = pix / c.W, i = pix - j * c.W;
    double cam[3];
    if (c.k_f32) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
            // (every K of the reference has zero skew: the middle product is 0 and any summation order rounds alike; with a skewed
            //  float32 K numpy's strided BLAS dot was seen to round differently in the last ulp -- not reproduced)
            cam[a] = (double)dot3_chain_f((float)i, (float)j, 1.0f, (float)c.kinv[3 * a], (float)c.kinv[3 * a + 1], (float)c.kinv[3 * a + 2]);
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a)
            cam[a] = dot3_chain((double)i, (double)j, 1.0, c.kinv[3 * a], c.kinv[3 * a + 1], c.kinv[3 * a + 2]);
    }
    if (c.all_f32) {
        // every operand is float32 (numpy keeps float32 throughout get_rays_from_KRT); the clamp compares and assigns float32(1e-5)
        const float q0 = __fsub_rn((float)cam[0], (float)c.T[0]), q1 = __fsub_rn((float)cam[1], (float)c.T[1]),
                    q2 = __fsub_rn((float)cam[2], (float)c.T[2]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float w = dot3_sdot_f(q0, q1, q2, (float)c.R[a], (float)c.R[3 + a], (float)c.R[6 + a]);
            float d = __fsub_rn(w, (float)c.o[a]);
            if (fabsf(d) < 1e-5f) d = 1e-5f;
            r.d[a] = (double)d;
        }
    } else {
        const double q0 = __dsub_rn(cam[0], c.T[0]), q1 = __dsub_rn(cam[1], c.T[1]), q2 = __dsub_rn(cam[2], c.T[2]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double w = dot3_chain(q0, q1, q2, c.R[a], c.R[3 + a], c.R[6 + a]);
            double d = __dsub_rn(w, c.o[a]);
            if (fabs(d) < 1e-5) d = 1e-5;
            r.d[a] = d;
        }
    }
    // six plane hits in the reference's order: (min x, min y, min z, max x, max y, max z)
    const double eps = 1e-6;
    int n_in = 0;
    double dist[2] = {0.0, 0.0};
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const int a = p % 3;
        const double bound = p < 3 ? c.lo[a] : c.hi[a];
        const double t = __ddiv_rn(__dsub_rn(bound, c.o[a]), r.d[a]);
        double pt[3];
        bool in = true;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            pt[b] = __dadd_rn(__dmul_rn(t, r.d[b]), c.o[b]);
            in = in && pt[b] >= __dsub_rn(c.lo[b], eps) && pt[b] <= __dadd_rn(c.hi[b], eps);
        }
        if (in) {
            if (n_in < 2) {
                const double e0 = __dsub_rn(pt[0], c.o[0]), e1 = __dsub_rn(pt[1], c.o[1]), e2 = __dsub_rn(pt[2], c.o[2]);
                dist[n_in] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(e0, e0), __dmul_rn(e1, e1)), __dmul_rn(e2, e2)));
            }
            ++n_in;
        }
    }
    r.hit = n_in == 2;
    double nd;
    if (c.all_f32) {                                   // np.linalg.norm of a float32 array stays float32
        const float x = (float)r.d[0], y = (float)r.d[1], z = (float)r.d[2];
        nd = (double)__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    } else {
        nd = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(r.d[0], r.d[0]), __dmul_rn(r.d[1], r.d[1])), __dmul_rn(r.d[2], r.d[2])));
    }
    const double d0 = __ddiv_rn(dist[0], nd), d1 = __ddiv_rn(dist[1], nd);
    r.near = (float)fmin(d0, d1);
    r.far = (float)fmax(d0, d1);
    return r;
}

// hits of this thread's pixel ranked inside the block; returns the rank of this thread (valid if `hit`) and the block total
__device__ __forceinline__ int block_rank(bool hit, int *total) {
    __shared__ int warp_count[RAYS_BLOCK / 32];
    const unsigned ballot = __ballot_sync(OCC_FULL, hit);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_count[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        int v = warp_count[lane], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(OCC_FULL, incl, o);
            if (lane >= o) incl += n;
        }
        warp_count[lane] = incl - v;                    // exclusive
        if (lane == 31) *total = incl;
    }
    __syncthreads();
    return warp_count[warp] + __popc(ballot & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(RAYS_BLOCK) rays_count_kernel(RayCam c, uint8_t *__restrict__ mask, int *__restrict__ block_count) {
    __shared__ int total;
    const long pix = (long)blockIdx.x * RAYS_BLOCK + threadIdx.x;
    const long P = (long)c.H * c.W;
    bool hit = false;
    if (pix < P) {
        hit = ray_for_pixel(c, (int)pix).hit;
        mask[pix] = hit ? 1 : 0;
    }
    block_rank(hit, &total);
    if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}

// exclusive scan of block_count[0..nb) in place; count[0] = number of hit rays
__global__ void __launch_bounds__(1024) rays_scan_kernel(int *__restrict__ block_count, int nb, int *__restrict__ count) {
    __shared__ int warp_sum_s[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < nb ? block_count[idx] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(OCC_FULL, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_sum_s[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sum_s[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int n = __shfl_up_sync(OCC_FULL, wi, o);
                if (lane >= o) wi += n;
            }
            warp_sum_s[lane] = wi - w;
        }
        __syncthreads();
        const int excl = carry + warp_sum_s[warp] + incl - v;
        if (idx < nb) block_count[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) count[0] = carry;
}

__global__ void __launch_bounds__(RAYS_BLOCK) rays_emit_kernel(RayCam c, const int *__restrict__ block_offset, int capacity,
                                                               float *__restrict__ rays, int *__restrict__ pixel_index) {
    __shared__ int total;
    const long pix = (long)blockIdx.x * RAYS_BLOCK + threadIdx.x;
    const long P = (long)c.H * c.W;
    RayOut r;
    r.hit = false;
    if (pix < P) r = ray_for_pixel(c, (int)pix);
    const int slot = block_offset[blockIdx.x] + block_rank(r.hit, &total);
    if (r.hit && slot < capacity) {
        float4 *dst = reinterpret_cast<float4 *>(rays + (size_t)slot * 8);
        dst[0] = make_float4((float)c.o[0], (float)c.o[1], (float)c.o[2], (float)r.d[0]);
        dst[1] = make_float4((float)r.d[1], (float)r.d[2], r.near, r.far);
        if (pixel_index) pixel_index[slot] = (int)pix;
    }
}

extern "C" long occnerf_rays_scratch_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    return (long)occ_div_up((long)H * W, RAYS_BLOCK) * (long)sizeof(int);
}

extern "C" int occnerf_generate_rays(const double *kinv_host, int k_is_f32, const double *R_host, const double *T_host,
                                     const double *bbox_min_host, const double *bbox_max_host, int H, int W, int capacity,
                                     float *rays, uint8_t *mask, int *pixel_index, int *count, void *scratch,
                                     occnerf_stream_t stream) {
    OCC_CHECK_ARG(kinv_host && R_host && T_host && bbox_min_host && bbox_max_host, "occnerf_generate_rays: NULL camera / box");
    OCC_CHECK_ARG(k_is_f32 >= 0 && k_is_f32 <= OCCNERF_RAYS_ALL_F32, "occnerf_generate_rays: camera dtype mode %d", k_is_f32);
    OCC_CHECK_ARG(H > 0 && W > 0 && (long)H * W < (1L << 31), "occnerf_generate_rays: H=%d W=%d out of range", H, W);
    OCC_CHECK_ARG(capacity >= 0 && (capacity == 0 || rays), "occnerf_generate_rays: rays is NULL with capacity %d", capacity);
    OCC_CHECK_ARG(mask && count && scratch, "occnerf_generate_rays: NULL mask / count / scratch");
    OCC_CHECK_ARG(((uintptr_t)rays & 15) == 0, "occnerf_generate_rays: rays must be 16-byte aligned");
    RayCam c;
    for (int a = 0; a < 9; ++a) {
        c.kinv[a] = kinv_host[a];
        c.R[a] = R_host[a];
    }
    for (int a = 0; a < 3; ++a) {
        c.T[a] = T_host[a];
        // bounds + [-0.01, 0.01] (camera_util.py:180), float64
        c.lo[a] = bbox_min_host[a] + -0.01;
        c.hi[a] = bbox_max_host[a] + 0.01;
    }
    // rays_o = -np.dot(R.T, T): o[a] = -sum_k R[k][a] T[k]  (same k = 0,1,2 chain as on the device)
    const bool all_f32 = k_is_f32 == OCCNERF_RAYS_ALL_F32;
    for (int a = 0; a < 3; ++a) {
        if (all_f32) {
            float acc = (float)R_host[a] * (float)T_host[0];
            acc = __builtin_fmaf((float)R_host[3 + a], (float)T_host[1], acc);
            acc = __builtin_fmaf((float)R_host[6 + a], (float)T_host[2], acc);
            c.o[a] = (double)-acc;
        } else {
            double acc = R_host[a] * T_host[0];
            acc = __builtin_fma(R_host[3 + a], T_host[1], acc);
            acc = __builtin_fma(R_host[6 + a], T_host[2], acc);
            c.o[a] = -acc;
        }
    }
    c.H = H;
    c.W = W;
    c.k_f32 = k_is_f32 ? 1 : 0;
    c.all_f32 = all_f32 ? 1 : 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = occ_div_up((long)H * W, RAYS_BLOCK);
    int *block_count = (int *)scratch;
    rays_count_kernel<<<nb, RAYS_BLOCK, 0, s>>>(c, mask, block_count);
    OCC_LAUNCH_CHECK();
    rays_scan_kernel<<<1, 1024, 0, s>>>(block_count, (int)nb, count);
    OCC_LAUNCH_CHECK();
    if (capacity > 0) {
        rays_emit_kernel<<<nb, RAYS_BLOCK, 0, s>>>(c, block_count, capacity, rays, pixel_index);
        OCC_LAUNCH_CHECK();
    }
    return OCCNERF_OK;
}
