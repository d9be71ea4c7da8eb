// Training-patch selection on the device (the part that was still on the host):
// core/data/occnerf/train.py:167-222 (get_patch_ray_indices) + :225-273 (_get_patch_ray_indices) + the gathers of :160-165
// (sample_patch_rays).  For each of n_patch patches the reference picks the candidate region (subject, or bbox minus subject), takes the
// select_idx-th candidate pixel in row-major order as the centre, clips a P x P window into the image and returns, for the window's
// pixels that hit the bounding box, their ranks in the compacted ray list (`select_inds`), the window's hit mask, its corners and
// the running totals (`patch_div_indices`).  The two random draws per patch stay with the caller (numpy's RandomState, so that a
// seeded run selects the same patches as the reference); everything else -- three mask scans, the rank select, the window
// compaction and the gather of the selected rays -- is integer work done here, bit-exact against the reference's own code
// (tests/golden/patches.npz).
//
//   patch_count  : per 1024-pixel block: in-block exclusive ranks of the ray mask (u16 per pixel) and the block totals of the three masks
//   patch_scan   : one block, exclusive scans of the three arrays of block totals
//   patch_select : one block per patch: rank select of the centre (binary search over the block totals + one in-block scan),
//                  window clip, hit mask, in-window ranks, ranks in the ray list
//   patch_emit   : one block: patch_div_indices, compaction of the per-patch slots, gather of the selected rays
#include "common.cuh"

namespace {

constexpr int PB = 1024;

// exclusive scan of one 0/1 flag per thread over a 1024-thread block; returns the thread's rank, *total = number of flags
__device__ __forceinline__ int block_rank(bool flag, int *total, int *s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(OCC_FULL, flag);
    const int in_warp = __popc(b & ((1u << lane) - 1u));
    __syncthreads();                                   // (s_warp may still be read from a previous call)
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    if (warp == 0) {
        int v = s_warp[lane], x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(OCC_FULL, x, o);
            if (lane >= o) x += y;
        }
        s_warp[lane] = x - v;                          // exclusive prefix of the warp totals
        if (lane == 31) s_warp[32] = x;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + in_warp;
}

__global__ void __launch_bounds__(PB) patch_count_kernel(const uint8_t *__restrict__ ray_mask, const uint8_t *__restrict__ subject,
                                                         const uint8_t *__restrict__ bbox, long HW, int nb, uint16_t *__restrict__ rank_local,
                                                         int *__restrict__ tot) {
    __shared__ int s_warp[33];
    const long p = (long)blockIdx.x * PB + threadIdx.x;
    const bool in = p < HW;
    const bool r = in && ray_mask[p] != 0, s = in && subject[p] != 0, e = in && bbox[p] != 0 && subject[p] == 0;
    int t0, t1, t2;
    const int rk = block_rank(r, &t0, s_warp);
    block_rank(s, &t1, s_warp);
    block_rank(e, &t2, s_warp);
    if (in) rank_local[p] = (uint16_t)rk;
    if (threadIdx.x == 0) { tot[blockIdx.x] = t0; tot[nb + blockIdx.x] = t1; tot[2 * nb + blockIdx.x] = t2; }
}

// exclusive scans of tot[a][0..nb) in place, a = 0..2; totals[a] = sum
__global__ void __launch_bounds__(PB) patch_scan_kernel(int *__restrict__ tot, int nb, int *__restrict__ totals) {
    __shared__ int s[PB];
    __shared__ int carry;
    for (int a = 0; a < 3; ++a) {
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int base = 0; base < nb; base += PB) {
            const int i = base + threadIdx.x;
            const int v = i < nb ? tot[a * nb + i] : 0;
            s[threadIdx.x] = v;
            __syncthreads();
            for (int o = 1; o < PB; o <<= 1) {
                const int y = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
                __syncthreads();
                s[threadIdx.x] += y;
                __syncthreads();
            }
            if (i < nb) tot[a * nb + i] = carry + s[threadIdx.x] - v;
            __syncthreads();
            if (threadIdx.x == PB - 1) carry += s[PB - 1];
            __syncthreads();
        }
        if (threadIdx.x == 0) totals[a] = carry;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(PB) patch_select_kernel(const uint8_t *__restrict__ ray_mask, const uint8_t *__restrict__ subject,
                                                          const uint8_t *__restrict__ bbox, int H, int W, int P, int nb,
                                                          const uint16_t *__restrict__ rank_local, const int *__restrict__ tot,
                                                          const int *__restrict__ totals, const uint8_t *__restrict__ use_subject,
                                                          const int *__restrict__ select_idx, int *__restrict__ slot_inds,
                                                          int *__restrict__ counts, uint8_t *__restrict__ patch_masks, int *__restrict__ xy_min,
                                                          int *__restrict__ xy_max, int *__restrict__ status) {
    __shared__ int s_warp[33];
    __shared__ int s_block, s_target, s_center;
    const int k = blockIdx.x;
    const long HW = (long)H * W;
    const int which = use_subject[k] ? 1 : 2;
    if (threadIdx.x == 0) {
        int idx = select_idx[k];
        const int n = totals[which];
        if (idx < 0 || idx >= n) { atomicExch(status, 1); idx = n > 0 ? min(max(idx, 0), n - 1) : 0; }
        const int *pre = tot + which * nb;
        int lo = 0, hi = nb - 1;                         // last block whose exclusive prefix is <= idx
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (pre[mid] <= idx) lo = mid; else hi = mid - 1;
        }
        s_block = lo; s_target = idx - pre[lo]; s_center = 0;
    }
    __syncthreads();
    {   // the s_target-th candidate pixel of block s_block
        const long p = (long)s_block * PB + threadIdx.x;
        const bool c = p < HW && (which == 1 ? subject[p] != 0 : (bbox[p] != 0 && subject[p] == 0));
        int t;
        const int rk = block_rank(c, &t, s_warp);
        if (c && rk == s_target) s_center = (int)p;
        __syncthreads();
    }
    const int cy = s_center / W, cx = s_center - cy * W;
    const int half = P / 2;
    const int x0 = min(max(cx - half, 0), W - P), y0 = min(max(cy - half, 0), H - P);       // np.clip(a, 0, W - P)
    if (threadIdx.x == 0) { xy_min[2 * k] = x0; xy_min[2 * k + 1] = y0; xy_max[2 * k] = x0 + P; xy_max[2 * k + 1] = y0 + P; }
    int base = 0;
    for (int i0 = 0; i0 < P * P; i0 += PB) {            // window pixels in row-major order = increasing flat image index
        const int i = i0 + threadIdx.x;
        const bool in = i < P * P;
        const int y = y0 + (in ? i / P : 0), x = x0 + (in ? i % P : 0);
        const long p = (long)y * W + x;
        const bool hit = in && ray_mask[p] != 0;
        int t;
        const int pos = block_rank(hit, &t, s_warp);
        if (in) patch_masks[(long)k * P * P + i] = hit ? 1 : 0;
        if (hit) slot_inds[(long)k * P * P + base + pos] = tot[p / PB] + (int)rank_local[p];    // cumsum(ray_mask)[p] - 1
        base += t;
    }
    if (threadIdx.x == 0) counts[k] = base;
}

__global__ void __launch_bounds__(PB) patch_emit_kernel(const int *__restrict__ slot_inds, const int *__restrict__ counts, int n_patch, int PP,
                                                        int *__restrict__ select_inds, int *__restrict__ patch_div,
                                                        const float *__restrict__ rays, float *__restrict__ rays_out) {
    __shared__ int s_div[1025];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < n_patch; ++k) { s_div[k] = acc; acc += counts[k]; }
        s_div[n_patch] = acc;
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= n_patch; k += PB) patch_div[k] = s_div[k];
    for (int k = 0; k < n_patch; ++k) {
        const int n = s_div[k + 1] - s_div[k];
        for (int j = threadIdx.x; j < n; j += PB) {
            const int r = slot_inds[(long)k * PP + j];
            select_inds[s_div[k] + j] = r;
            if (rays && rays_out) {
                const float4 *src = reinterpret_cast<const float4 *>(rays + (long)r * 8);
                float4 *dst = reinterpret_cast<float4 *>(rays_out + (long)(s_div[k] + j) * 8);
                dst[0] = __ldg(src); dst[1] = __ldg(src + 1);
            }
        }
    }
}

long scratch_layout(long HW, int n_patch, int P, long *off_tot, long *off_totals, long *off_slots, long *off_counts) {
    const long nb = (HW + PB - 1) / PB;
    long o = (HW * 2 + 255) / 256 * 256;                // rank_local u16
    *off_tot = o; o += (3 * nb * 4 + 255) / 256 * 256;
    *off_totals = o; o += 256;
    *off_slots = o; o += ((long)n_patch * P * P * 4 + 255) / 256 * 256;
    *off_counts = o; o += ((long)n_patch * 4 + 255) / 256 * 256;
    return o;
}

}  // namespace

extern "C" long occnerf_patches_scratch_bytes(int H, int W, int n_patch, int patch) {
    if (H < 1 || W < 1 || n_patch < 1 || patch < 1) return -1;
    long a, b, c, d;
    return scratch_layout((long)H * W, n_patch, patch, &a, &b, &c, &d);
}

// ray_mask / subject_mask / bbox_mask: [H*W] bytes (0 / non-zero).  use_subject [n_patch] bytes and select_idx [n_patch] i32: the
// caller's two draws per patch (device memory).  Outputs: select_inds [n_patch * patch^2] i32 (the first patch_div[n_patch] entries
// are valid), patch_div [n_patch + 1] i32, patch_masks [n_patch, patch, patch] bytes, xy_min / xy_max [n_patch, 2] i32 (x, y),
// status [1] i32 (caller-zeroed; set to 1 when a select_idx is outside its candidate list -- np.random.choice could not have drawn it).
// rays [n_rays, 8] / rays_out [n_patch * patch^2, 8]: optional gather of the selected rays (both NULL to skip).
extern "C" int occnerf_sample_patches(const uint8_t *ray_mask, const uint8_t *subject_mask, const uint8_t *bbox_mask, int H, int W, int patch,
                                      int n_patch, const uint8_t *use_subject, const int32_t *select_idx, const float *rays, float *rays_out,
                                      int32_t *select_inds, int32_t *patch_div, uint8_t *patch_masks, int32_t *xy_min, int32_t *xy_max,
                                      int32_t *status, void *scratch, occnerf_stream_t stream) {
    OCC_CHECK_ARG(ray_mask && subject_mask && bbox_mask && use_subject && select_idx && select_inds && patch_div && patch_masks && xy_min &&
                  xy_max && status && scratch, "sample_patches: null pointer");
    OCC_CHECK_ARG(H >= 1 && W >= 1 && patch >= 1 && patch <= H && patch <= W && n_patch >= 1 && n_patch <= 1024,
                  "sample_patches: H=%d W=%d patch=%d n_patch=%d (patch must fit the image, at most 1024 patches)", H, W, patch, n_patch);
    OCC_CHECK_ARG((rays == nullptr) == (rays_out == nullptr), "sample_patches: rays and rays_out go together");
    const long HW = (long)H * W;
    OCC_CHECK_ARG(HW <= (1L << 30), "sample_patches: image too large");
    const int nb = (int)((HW + PB - 1) / PB);
    long o_tot, o_totals, o_slots, o_counts;
    scratch_layout(HW, n_patch, patch, &o_tot, &o_totals, &o_slots, &o_counts);
    unsigned char *ws = (unsigned char *)scratch;
    uint16_t *rank_local = (uint16_t *)ws;
    int *tot = (int *)(ws + o_tot), *totals = (int *)(ws + o_totals), *slots = (int *)(ws + o_slots), *counts = (int *)(ws + o_counts);
    cudaStream_t st = (cudaStream_t)stream;
    patch_count_kernel<<<nb, PB, 0, st>>>(ray_mask, subject_mask, bbox_mask, HW, nb, rank_local, tot);
    OCC_LAUNCH_CHECK();
    patch_scan_kernel<<<1, PB, 0, st>>>(tot, nb, totals);
    OCC_LAUNCH_CHECK();
    patch_select_kernel<<<n_patch, PB, 0, st>>>(ray_mask, subject_mask, bbox_mask, H, W, patch, nb, rank_local, tot, totals, use_subject,
                                                select_idx, slots, counts, patch_masks, xy_min, xy_max, status);
    OCC_LAUNCH_CHECK();
    patch_emit_kernel<<<1, PB, 0, st>>>(slots, counts, n_patch, patch * patch, select_inds, patch_div, rays, rays_out);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
