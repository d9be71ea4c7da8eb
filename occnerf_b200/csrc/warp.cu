// K1: fused ray sampling + 24-bone inverse-LBS warp through the motion-weight volume.
//
// Replaces the reference's _get_samples_along_ray / _stratified_sampling / `pts = o + d*z` /
// _sample_motion_fields chain (core/nets/occnerf/network.py:416-432,456,351-402): 48 small GEMMs, 24
// F.grid_sample launches and ~300 MB of stacked intermediates per 786k samples become one kernel that
// reads 32 B per ray and writes 20 B per sample.
//
// Arithmetic contract (tests/test_warp_gpu.py checks the integer voxel bins bit for bit against the
// oracle): every step is rounded exactly where eager PyTorch rounds it -- explicit __fmul_rn/__fadd_rn
// so that nvcc cannot contract across what are separate kernels upstream; the K=3 affine uses the
// mul, fma, fma order of the reference's sgemm; ATen's align_corners=True un-normalisation is
// ((g+1)/2)*(size-1).
//
// Mapping: one thread per sample, consecutive lanes = consecutive samples of one ray, so the 8-corner
// gathers of a warp fall into a handful of neighbouring voxels (sample spacing ~1.5 cm, voxel ~3-7 cm)
// and are served by L1; the 3 MiB volume itself is L2-resident.
#include "common.cuh"

namespace {

constexpr int kMaxBones = 32;
constexpr int kThreads = 128;

struct BoneSmem {
    float R[kMaxBones][9];
    float T[kMaxBones][3];
    float bmin[3], bscale[3];
};

__device__ __forceinline__ void load_bones(BoneSmem &sm, const float *Rs, const float *Ts, const float *bmin,
                                           const float *bscale, int nb) {
    for (int i = threadIdx.x; i < nb * 9; i += blockDim.x) sm.R[i / 9][i % 9] = __ldg(Rs + i);
    for (int i = threadIdx.x; i < nb * 3; i += blockDim.x) sm.T[i / 3][i % 3] = __ldg(Ts + i);
    if (threadIdx.x < 3) {
        sm.bmin[threadIdx.x] = __ldg(bmin + threadIdx.x);
        sm.bscale[threadIdx.x] = __ldg(bscale + threadIdx.x);
    }
    __syncthreads();
}

// network.py:416-432: z = near*(1-t) + far*t, optional stratified jitter.
__device__ __forceinline__ float lin_z(float nr, float fr, float t) {
    return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, t)), __fmul_rn(fr, t));
}
__device__ __forceinline__ float sample_z(float nr, float fr, const float *__restrict__ t_lin,
                                          const float *__restrict__ t_rand, long ray, int j, int S) {
    const float zj = lin_z(nr, fr, __ldg(t_lin + j));
    if (t_rand == nullptr) return zj;
    float upper = zj, lower = zj;
    if (j + 1 < S) upper = __fmul_rn(0.5f, __fadd_rn(lin_z(nr, fr, __ldg(t_lin + j + 1)), zj));
    if (j > 0) lower = __fmul_rn(0.5f, __fadd_rn(zj, lin_z(nr, fr, __ldg(t_lin + j - 1))));
    const float u = __ldg(t_rand + ray * S + j);
    return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
}

struct Cell {
    float q[3];      // bone-space position R.p + T
    float f0[3];     // weight of the floor corner per axis
    float f1[3];     // weight of the +1 corner per axis
    int c0[3];       // floor voxel (x,y,z)
};

__device__ __forceinline__ void locate(const BoneSmem &sm, int i, float px, float py, float pz, int vd, int vh, int vw,
                                       Cell &c) {
    const float *R = sm.R[i];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(R[r * 3 + 0], px);
        acc = __fmaf_rn(R[r * 3 + 1], py, acc);
        acc = __fmaf_rn(R[r * 3 + 2], pz, acc);
        c.q[r] = __fadd_rn(acc, sm.T[i][r]);
    }
    const float size_m1[3] = {(float)(vw - 1), (float)(vh - 1), (float)(vd - 1)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float g = __fsub_rn(__fmul_rn(__fsub_rn(c.q[a], sm.bmin[a]), sm.bscale[a]), 1.0f);
        const float ix = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), size_m1[a]);
        const float fl = floorf(ix);
        c.f1[a] = __fsub_rn(ix, fl);
        c.f0[a] = __fsub_rn(__fadd_rn(fl, 1.0f), ix);
        // clamp before the float->int conversion so that far-away / non-finite coordinates stay defined
        c.c0[a] = (int)fminf(fmaxf(fl, -1.0e9f), 1.0e9f);
    }
}

__global__ void __launch_bounds__(kThreads)
warp_fwd_kernel(const float *__restrict__ rays, const float *__restrict__ t_lin, const float *__restrict__ t_rand,
                const float *__restrict__ Rs, const float *__restrict__ Ts, const float *__restrict__ vol,
                const float *__restrict__ bmin, const float *__restrict__ bscale, long M, int S, int nb, int vd, int vh,
                int vw, float *__restrict__ z_out, float *__restrict__ x_skel, float *__restrict__ mask_out,
                int32_t *__restrict__ bins) {
    __shared__ BoneSmem sm;
    load_bones(sm, Rs, Ts, bmin, bscale, nb);
    const long m = (long)blockIdx.x * kThreads + threadIdx.x;
    if (m >= M) return;
    const long ray = m / S;
    const int j = (int)(m - ray * S);
    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays) + ray * 2);
    const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays) + ray * 2 + 1);
    // rays row = (ox,oy,oz,dx | dy,dz,near,far)
    const float z = sample_z(r1.z, r1.w, t_lin, t_rand, ray, j, S);
    const float px = __fadd_rn(r0.x, __fmul_rn(r0.w, z));
    const float py = __fadd_rn(r0.y, __fmul_rn(r1.x, z));
    const float pz = __fadd_rn(r0.z, __fmul_rn(r1.y, z));

    const long plane = (long)vh * vw, cube = (long)vd * plane;
    float total = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
    for (int i = 0; i < nb; ++i) {
        Cell c;
        locate(sm, i, px, py, pz, vd, vh, vw, c);
        if (bins) {
            int32_t *b = bins + (m * nb + i) * 3;
            b[0] = c.c0[0]; b[1] = c.c0[1]; b[2] = c.c0[2];
        }
        const int x0 = c.c0[0], y0 = c.c0[1], z0 = c.c0[2];
        if (x0 < -1 || x0 >= vw || y0 < -1 || y0 >= vh || z0 < -1 || z0 >= vd) continue;   // all 8 corners outside
        const float *v = vol + (long)i * cube;
        float w = 0.f;
#pragma unroll
        for (int dz = 0; dz < 2; ++dz) {
            const int zi = z0 + dz;
            const float wz = dz ? c.f1[2] : c.f0[2];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int yi = y0 + dy;
                const float wy = dy ? c.f1[1] : c.f0[1];
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int xi = x0 + dx;
                    const float wx = dx ? c.f1[0] : c.f0[0];
                    const bool ok = xi >= 0 && xi < vw && yi >= 0 && yi < vh && zi >= 0 && zi < vd;
                    if (ok) {
                        const float val = __ldg(v + zi * plane + (long)yi * vw + xi);
                        w = __fadd_rn(w, __fmul_rn(val, __fmul_rn(__fmul_rn(wx, wy), wz)));
                    }
                }
            }
        }
        total = __fadd_rn(total, w);
        sx = __fadd_rn(sx, __fmul_rn(w, c.q[0]));
        sy = __fadd_rn(sy, __fmul_rn(w, c.q[1]));
        sz = __fadd_rn(sz, __fmul_rn(w, c.q[2]));
    }
    const float den = fmaxf(total, 1e-4f);
    z_out[m] = z;
    mask_out[m] = total;
    x_skel[m * 3 + 0] = __fdiv_rn(sx, den);
    x_skel[m * 3 + 1] = __fdiv_rn(sy, den);
    x_skel[m * 3 + 2] = __fdiv_rn(sz, den);
}

// d(mask)/d(vol): every in-range corner of every bone receives g_mask * trilinear weight.
//
// Mapping: thread = (segment of kSeg consecutive samples of one ray, bone), bone fastest.  Consecutive samples (~1.5 cm
// apart) stay in the same voxel (~7 cm) of a bone's volume for several steps, so the thread sums the 8 corner
// contributions in registers and issues the 8 reductions only when the floor voxel changes: ~3x fewer L2 atomics than
// one-thread-per-sample, and the lanes of a warp (different bones) never hit the same address at the same time.
constexpr int kSeg = 16;

__global__ void __launch_bounds__(kThreads)
warp_bwd_kernel(const float *__restrict__ rays, const float *__restrict__ t_lin, const float *__restrict__ t_rand,
                const float *__restrict__ Rs, const float *__restrict__ Ts, const float *__restrict__ bmin,
                const float *__restrict__ bscale, const float *__restrict__ g_mask, long N, int S, int nb, int vd,
                int vh, int vw, float *__restrict__ g_vol) {
    __shared__ BoneSmem sm;
    load_bones(sm, Rs, Ts, bmin, bscale, nb);
    const int segs = (S + kSeg - 1) / kSeg;
    const long gid = (long)blockIdx.x * kThreads + threadIdx.x;
    if (gid >= N * segs * nb) return;
    const int i = (int)(gid % nb);
    const long rs = gid / nb;
    const long ray = rs / segs;
    const int j0 = (int)(rs - ray * segs) * kSeg, j1 = min(S, j0 + kSeg);
    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays) + ray * 2);
    const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays) + ray * 2 + 1);
    const long plane = (long)vh * vw, cube = (long)vd * plane;
    float *v = g_vol + (long)i * cube;
    int cx = 0, cy = 0, cz = 0;
    bool have = false;
    float acc[8];
    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int xi = cx + (k & 1), yi = cy + ((k >> 1) & 1), zi = cz + (k >> 2);
            if (acc[k] != 0.f && xi >= 0 && xi < vw && yi >= 0 && yi < vh && zi >= 0 && zi < vd)
                atomicAdd(v + zi * plane + (long)yi * vw + xi, acc[k]);
        }
    };
    for (int j = j0; j < j1; ++j) {
        const float gm = __ldg(g_mask + ray * S + j);
        if (gm == 0.f) continue;
        const float z = sample_z(r1.z, r1.w, t_lin, t_rand, ray, j, S);
        const float px = __fadd_rn(r0.x, __fmul_rn(r0.w, z));
        const float py = __fadd_rn(r0.y, __fmul_rn(r1.x, z));
        const float pz = __fadd_rn(r0.z, __fmul_rn(r1.y, z));
        Cell c;
        locate(sm, i, px, py, pz, vd, vh, vw, c);
        const int x0 = c.c0[0], y0 = c.c0[1], z0 = c.c0[2];
        if (x0 < -1 || x0 >= vw || y0 < -1 || y0 >= vh || z0 < -1 || z0 >= vd) continue;   // all 8 corners outside
        if (!have || x0 != cx || y0 != cy || z0 != cz) {
            if (have) flush();
            cx = x0; cy = y0; cz = z0;
            have = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float wx = (k & 1) ? c.f1[0] : c.f0[0], wy = (k & 2) ? c.f1[1] : c.f0[1], wz = (k & 4) ? c.f1[2] : c.f0[2];
            acc[k] += gm * ((wx * wy) * wz);
        }
    }
    if (have) flush();
}

int check_common(const void *rays, const void *t_lin, const void *Rs, const void *Ts, int N, int S, int nb, int vd,
                 int vh, int vw) {
    OCC_CHECK_ARG(rays && t_lin && Rs && Ts, "warp: null input pointer");
    OCC_CHECK_ARG(N >= 0 && S >= 1 && S <= 4096, "warp: bad N=%d S=%d", N, S);
    OCC_CHECK_ARG(nb >= 1 && nb <= kMaxBones, "warp: nb=%d outside [1,%d]", nb, kMaxBones);
    OCC_CHECK_ARG(vd >= 2 && vh >= 2 && vw >= 2, "warp: volume %dx%dx%d too small", vd, vh, vw);
    OCC_CHECK_ARG(((uintptr_t)rays & 15) == 0, "warp: rays must be 16-byte aligned");
    return 0;
}

}  // namespace

extern "C" int occnerf_warp_forward(const float *rays, const float *t_lin, const float *t_rand, const float *Rs,
                                    const float *Ts, const float *vol, const float *bbox_min, const float *bbox_scale,
                                    int N, int S, int nb, int vd, int vh, int vw, float *z, float *x_skel, float *mask,
                                    int32_t *bins, occnerf_stream_t stream) {
    if (N == 0) return OCCNERF_OK;
    if (int e = check_common(rays, t_lin, Rs, Ts, N, S, nb, vd, vh, vw)) return e;
    OCC_CHECK_ARG(vol && bbox_min && bbox_scale && z && x_skel && mask, "warp_forward: null pointer");
    const long M = (long)N * S;
    if (M == 0) return OCCNERF_OK;
    warp_fwd_kernel<<<occ_div_up(M, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        rays, t_lin, t_rand, Rs, Ts, vol, bbox_min, bbox_scale, M, S, nb, vd, vh, vw, z, x_skel, mask, bins);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_warp_backward(const float *rays, const float *t_lin, const float *t_rand, const float *Rs,
                                     const float *Ts, const float *bbox_min, const float *bbox_scale,
                                     const float *g_mask, int N, int S, int nb, int vd, int vh, int vw, float *g_vol,
                                     occnerf_stream_t stream) {
    if (N == 0) return OCCNERF_OK;
    if (int e = check_common(rays, t_lin, Rs, Ts, N, S, nb, vd, vh, vw)) return e;
    OCC_CHECK_ARG(bbox_min && bbox_scale && g_mask && g_vol, "warp_backward: null pointer");
    const long M = (long)N * S;
    if (M == 0) return OCCNERF_OK;
    const long threads = (long)N * ((S + kSeg - 1) / kSeg) * nb;
    warp_bwd_kernel<<<occ_div_up(threads, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        rays, t_lin, t_rand, Rs, Ts, bbox_min, bbox_scale, g_mask, N, S, nb, vd, vh, vw, g_vol);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
