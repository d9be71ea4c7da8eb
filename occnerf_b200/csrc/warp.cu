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


// =====================================================================================================================
// Corner-packed variant (the default path of ops.warp_forward / warp_backward).
//
// The 24 bones do NOT share a voxel: bone i is sampled at ITS OWN position q_i = R_i p + T_i, so a bone-interleaved
// [z][y][x][bone] volume would still cost one gather per (bone, corner).  What the 8 corners of one (sample, bone) lookup
// do share is the floor voxel, so the per-frame volume is re-laid out once (occnerf_warp_pack_volume, ~10 us) as
//     vol8 [bone][z0+1][y0+1][x0+1][8 corners],   z0,y0,x0 in [-1, size-1]  (zero padding baked in)
// = one aligned 32-byte sector per lookup: 2 x LDG.128 instead of 8 x LDG.32 with per-corner bounds tests, 24 sectors per
// sample instead of ~100.  27.6 MB for 24 x 32^3 -- L2-resident on B200 (126 MB).  The corner order k = dz*4 + dy*2 + dx
// and the mul/add sequence are those of the scalar kernel above, so z / bins are bit-identical and mask / x_skel agree
// to the last bit as well (adding an exact 0 for an outside corner changes nothing).
//
// Staging: a block owns 128 consecutive samples (one ray at S = 128).  Its inputs -- the bone table (Rs, Ts), its rays and
// its tile of the jitter tensor -- arrive in shared memory through TMA bulk copies (cp.async.bulk, one elected thread,
// one mbarrier) instead of per-thread loads + __syncthreads.
//
// Backward: same (16-sample segment, bone) walk as warp_bwd_kernel, but a run of samples in one floor voxel ends with two
// red.global.add.v4.f32 into g_vol8 (the packed layout) instead of 8 scalar REDs; occnerf_warp_unpack_grad folds g_vol8
// back into the reference layout.  With g_Rs/g_Ts requested it also produces d mask / d (R_i, T_i) -- the gradient that
// F.grid_sample gives the reference w.r.t. its grid (network.py:367-370), needed once the pose decoder trains.
constexpr int kTile = kThreads;          // samples per block

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct __align__(16) FwdSmem {
    BoneSmem bones;                      // R[32][9] and T[32][3] are flat images of Rs / Ts (offsets 0 and 1152: 16-byte aligned)
    __align__(16) float t_rand[kTile];
    __align__(16) float rays[(kTile + 1) * 8];   // the rays touched by this block (at most kTile / S + 1 <= kTile + 1)
    __align__(8) unsigned long long bar;
};

__device__ __forceinline__ long cell_index(int i, int x0, int y0, int z0, int vd, int vh, int vw) {
    return ((((long)i * (vd + 1) + (z0 + 1)) * (vh + 1) + (y0 + 1)) * (vw + 1) + (x0 + 1)) * 8;
}

__global__ void __launch_bounds__(256)
pack_volume_kernel(const float *__restrict__ vol, int nb, int vd, int vh, int vw, float *__restrict__ vol8) {
    const long cells = (long)nb * (vd + 1) * (vh + 1) * (vw + 1);
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const int x0 = (int)(c % (vw + 1)) - 1;
    long r = c / (vw + 1);
    const int y0 = (int)(r % (vh + 1)) - 1;
    r /= (vh + 1);
    const int z0 = (int)(r % (vd + 1)) - 1;
    const int i = (int)(r / (vd + 1));
    const float *v = vol + (long)i * vd * vh * vw;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int xi = x0 + (k & 1), yi = y0 + ((k >> 1) & 1), zi = z0 + (k >> 2);
        const bool ok = xi >= 0 && xi < vw && yi >= 0 && yi < vh && zi >= 0 && zi < vd;
        o[k] = ok ? __ldg(v + ((long)zi * vh + yi) * vw + xi) : 0.f;
    }
    float4 *dst = reinterpret_cast<float4 *>(vol8 + c * 8);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// g_vol [channels][vd][vh][vw] = fold of g_vol8; channels >= nb (the background channel) get zero
__global__ void __launch_bounds__(256)
unpack_grad_kernel(const float *__restrict__ g_vol8, int nb, int channels, int vd, int vh, int vw, float *__restrict__ g_vol) {
    const long total = (long)channels * vd * vh * vw;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int x = (int)(e % vw);
    long r = e / vw;
    const int y = (int)(r % vh);
    r /= vh;
    const int z = (int)(r % vd);
    const int i = (int)(r / vd);
    float s = 0.f;
    if (i < nb) {
#pragma unroll
        for (int k = 0; k < 8; ++k)      // corner k of floor voxel (x - dx, y - dy, z - dz) is this voxel
            s += __ldg(g_vol8 + cell_index(i, x - (k & 1), y - ((k >> 1) & 1), z - (k >> 2), vd, vh, vw) + k);
    }
    g_vol[e] = s;
}

template <bool BULK>
__global__ void __launch_bounds__(kThreads)
warp_fwd_packed_kernel(const float *__restrict__ rays, const float *__restrict__ t_lin, const float *__restrict__ t_rand,
                       const float *__restrict__ Rs, const float *__restrict__ Ts, const float *__restrict__ vol8,
                       const float *__restrict__ bmin, const float *__restrict__ bscale, long M, int S, int nb, int vd, int vh,
                       int vw, float *__restrict__ z_out, float *__restrict__ x_skel, float *__restrict__ mask_out) {
    __shared__ FwdSmem sm;
    const long m0 = (long)blockIdx.x * kTile;
    const int n_here = (int)min((long)kTile, M - m0);
    const long ray0 = m0 / S, ray1 = (m0 + n_here - 1) / S;
    const int n_rays = (int)(ray1 - ray0 + 1);
    if (BULK) {
        const uint32_t bar = smem_addr(&sm.bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t bytes = (uint32_t)nb * 48u + (uint32_t)n_rays * 32u + (t_rand ? (uint32_t)n_here * 4u : 0u);
            asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
            bulk_load(&sm.bones.R[0][0], Rs, (uint32_t)nb * 36u, bar);
            bulk_load(&sm.bones.T[0][0], Ts, (uint32_t)nb * 12u, bar);
            bulk_load(sm.rays, rays + ray0 * 8, (uint32_t)n_rays * 32u, bar);
            if (t_rand) bulk_load(sm.t_rand, t_rand + m0, (uint32_t)n_here * 4u, bar);
        }
        if (threadIdx.x < 3) {
            sm.bones.bmin[threadIdx.x] = __ldg(bmin + threadIdx.x);
            sm.bones.bscale[threadIdx.x] = __ldg(bscale + threadIdx.x);
        }
        __syncthreads();                                   // barrier initialised, bbox visible
        uint32_t ok = 0, spins = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar) : "memory");
            if (!ok && ++spins > (1u << 26)) __trap();       // never hang the GPU
        }
    } else {
        for (int i = threadIdx.x; i < nb * 9; i += blockDim.x) sm.bones.R[i / 9][i % 9] = __ldg(Rs + i);
        for (int i = threadIdx.x; i < nb * 3; i += blockDim.x) sm.bones.T[i / 3][i % 3] = __ldg(Ts + i);
        for (int i = threadIdx.x; i < n_rays * 8; i += blockDim.x) sm.rays[i] = __ldg(rays + ray0 * 8 + i);
        if (t_rand && threadIdx.x < n_here) sm.t_rand[threadIdx.x] = __ldg(t_rand + m0 + threadIdx.x);
        if (threadIdx.x < 3) {
            sm.bones.bmin[threadIdx.x] = __ldg(bmin + threadIdx.x);
            sm.bones.bscale[threadIdx.x] = __ldg(bscale + threadIdx.x);
        }
        __syncthreads();
    }
    if ((int)threadIdx.x >= n_here) return;
    const long m = m0 + threadIdx.x;
    const long ray = m / S;
    const int j = (int)(m - ray * S);
    const float *rr = sm.rays + (ray - ray0) * 8;            // (ox,oy,oz,dx,dy,dz,near,far)
    const float nr = rr[6], fr = rr[7];
    float z = lin_z(nr, fr, __ldg(t_lin + j));
    if (t_rand != nullptr) {
        float upper = z, lower = z;
        if (j + 1 < S) upper = __fmul_rn(0.5f, __fadd_rn(lin_z(nr, fr, __ldg(t_lin + j + 1)), z));
        if (j > 0) lower = __fmul_rn(0.5f, __fadd_rn(z, lin_z(nr, fr, __ldg(t_lin + j - 1))));
        z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), sm.t_rand[threadIdx.x]));
    }
    const float px = __fadd_rn(rr[0], __fmul_rn(rr[3], z));
    const float py = __fadd_rn(rr[1], __fmul_rn(rr[4], z));
    const float pz = __fadd_rn(rr[2], __fmul_rn(rr[5], z));

    float total = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
    for (int i = 0; i < nb; ++i) {
        Cell c;
        locate(sm.bones, i, px, py, pz, vd, vh, vw, c);
        const int x0 = c.c0[0], y0 = c.c0[1], z0 = c.c0[2];
        if (x0 < -1 || x0 >= vw || y0 < -1 || y0 >= vh || z0 < -1 || z0 >= vd) continue;   // all 8 corners outside
        const float4 *cell = reinterpret_cast<const float4 *>(vol8 + cell_index(i, x0, y0, z0, vd, vh, vw));
        const float4 lo = __ldg(cell), hi = __ldg(cell + 1);
        const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        float w = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float wx = (k & 1) ? c.f1[0] : c.f0[0], wy = (k & 2) ? c.f1[1] : c.f0[1], wz = (k & 4) ? c.f1[2] : c.f0[2];
            w = __fadd_rn(w, __fmul_rn(v[k], __fmul_rn(__fmul_rn(wx, wy), wz)));
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

template <bool POSE>
__global__ void __launch_bounds__(kThreads)
warp_bwd_packed_kernel(const float *__restrict__ rays, const float *__restrict__ t_lin, const float *__restrict__ t_rand,
                       const float *__restrict__ Rs, const float *__restrict__ Ts, const float *__restrict__ vol8,
                       const float *__restrict__ bmin, const float *__restrict__ bscale, const float *__restrict__ g_mask, long N,
                       int S, int nb, int vd, int vh, int vw, float *__restrict__ g_vol8, float *__restrict__ g_Rs,
                       float *__restrict__ g_Ts) {
    __shared__ BoneSmem sm;
    __shared__ float pose_acc[POSE ? kMaxBones * 12 : 1];
    if (POSE) for (int i = threadIdx.x; i < kMaxBones * 12; i += blockDim.x) pose_acc[i] = 0.f;
    load_bones(sm, Rs, Ts, bmin, bscale, nb);
    const int segs = (S + kSeg - 1) / kSeg;
    const long gid = (long)blockIdx.x * kThreads + threadIdx.x;
    const bool active = gid < N * segs * nb;
    const int i = (int)(gid % nb);
    float pg[12];
    if (POSE) {
#pragma unroll
        for (int k = 0; k < 12; ++k) pg[k] = 0.f;
    }
    if (active) {
        const long rs = gid / nb;
        const long ray = rs / segs;
        const int j0 = (int)(rs - ray * segs) * kSeg, j1 = min(S, j0 + kSeg);
        const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays) + ray * 2);
        const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays) + ray * 2 + 1);
        int cx = 0, cy = 0, cz = 0;
        bool have = false;
        float acc[8], val[8];
        auto flush = [&]() {
            float *cell = g_vol8 + cell_index(i, cx, cy, cz, vd, vh, vw);
            if (acc[0] != 0.f || acc[1] != 0.f || acc[2] != 0.f || acc[3] != 0.f) red_add_v4(cell, acc[0], acc[1], acc[2], acc[3]);
            if (acc[4] != 0.f || acc[5] != 0.f || acc[6] != 0.f || acc[7] != 0.f) red_add_v4(cell + 4, acc[4], acc[5], acc[6], acc[7]);
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
                if (POSE) {
                    const float4 *cell = reinterpret_cast<const float4 *>(vol8 + cell_index(i, x0, y0, z0, vd, vh, vw));
                    const float4 lo = __ldg(cell), hi = __ldg(cell + 1);
                    val[0] = lo.x; val[1] = lo.y; val[2] = lo.z; val[3] = lo.w; val[4] = hi.x; val[5] = hi.y; val[6] = hi.z; val[7] = hi.w;
                }
            }
            float dwx = 0.f, dwy = 0.f, dwz = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float wx = (k & 1) ? c.f1[0] : c.f0[0], wy = (k & 2) ? c.f1[1] : c.f0[1], wz = (k & 4) ? c.f1[2] : c.f0[2];
                acc[k] += gm * ((wx * wy) * wz);
                if (POSE) {      // d w / d (ix, iy, iz): the corner's weight with one factor replaced by -1 / +1
                    dwx += val[k] * ((k & 1) ? 1.f : -1.f) * (wy * wz);
                    dwy += val[k] * ((k & 2) ? 1.f : -1.f) * (wx * wz);
                    dwz += val[k] * ((k & 4) ? 1.f : -1.f) * (wx * wy);
                }
            }
            if (POSE) {
                // ix = ((g + 1) / 2) (size - 1), g = (q - bmin) bscale - 1  ->  d ix / d q = bscale (size - 1) / 2
                const float gq[3] = {gm * dwx * sm.bscale[0] * 0.5f * (float)(vw - 1), gm * dwy * sm.bscale[1] * 0.5f * (float)(vh - 1),
                                     gm * dwz * sm.bscale[2] * 0.5f * (float)(vd - 1)};
                const float p[3] = {px, py, pz};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    pg[9 + r] += gq[r];
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) pg[r * 3 + cc] += gq[r] * p[cc];
                }
            }
        }
        if (have) flush();
    }
    if (POSE) {
        if (active) {
#pragma unroll
            for (int k = 0; k < 12; ++k)
                if (pg[k] != 0.f) atomicAdd(&pose_acc[i * 12 + k], pg[k]);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nb * 12; e += blockDim.x) {
            const float v = pose_acc[e];
            if (v != 0.f) {
                const int b = e / 12, k = e % 12;
                atomicAdd(k < 9 ? g_Rs + b * 9 + k : g_Ts + b * 3 + (k - 9), v);
            }
        }
    }
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

extern "C" long occnerf_warp_packed_floats(int nb, int vd, int vh, int vw) {
    if (nb < 1 || vd < 2 || vh < 2 || vw < 2) return -1;
    return (long)nb * (vd + 1) * (vh + 1) * (vw + 1) * 8;
}

extern "C" int occnerf_warp_pack_volume(const float *vol, int nb, int vd, int vh, int vw, float *vol8, occnerf_stream_t stream) {
    OCC_CHECK_ARG(vol && vol8, "warp_pack_volume: null pointer");
    OCC_CHECK_ARG(nb >= 1 && nb <= kMaxBones && vd >= 2 && vh >= 2 && vw >= 2, "warp_pack_volume: nb=%d volume %dx%dx%d", nb, vd, vh, vw);
    OCC_CHECK_ARG(((uintptr_t)vol8 & 15) == 0, "warp_pack_volume: vol8 must be 16-byte aligned");
    const long cells = (long)nb * (vd + 1) * (vh + 1) * (vw + 1);
    pack_volume_kernel<<<occ_div_up(cells, 256), 256, 0, (cudaStream_t)stream>>>(vol, nb, vd, vh, vw, vol8);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_warp_unpack_grad(const float *g_vol8, int nb, int channels, int vd, int vh, int vw, float *g_vol,
                                        occnerf_stream_t stream) {
    OCC_CHECK_ARG(g_vol8 && g_vol, "warp_unpack_grad: null pointer");
    OCC_CHECK_ARG(nb >= 1 && nb <= kMaxBones && channels >= nb && vd >= 2 && vh >= 2 && vw >= 2, "warp_unpack_grad: nb=%d channels=%d", nb, channels);
    const long total = (long)channels * vd * vh * vw;
    unpack_grad_kernel<<<occ_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(g_vol8, nb, channels, vd, vh, vw, g_vol);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_warp_forward_packed(const float *rays, const float *t_lin, const float *t_rand, const float *Rs,
                                           const float *Ts, const float *vol8, const float *bbox_min, const float *bbox_scale,
                                           int N, int S, int nb, int vd, int vh, int vw, float *z, float *x_skel, float *mask,
                                           occnerf_stream_t stream) {
    if (N == 0) return OCCNERF_OK;
    if (int e = check_common(rays, t_lin, Rs, Ts, N, S, nb, vd, vh, vw)) return e;
    OCC_CHECK_ARG(vol8 && bbox_min && bbox_scale && z && x_skel && mask, "warp_forward_packed: null pointer");
    OCC_CHECK_ARG(((uintptr_t)vol8 & 15) == 0, "warp_forward_packed: vol8 must be 16-byte aligned");
    const long M = (long)N * S;
    // TMA bulk copies need 16-byte aligned sources and sizes that are multiples of 16 bytes
    const bool bulk = nb % 4 == 0 && S % 4 == 0 && (((uintptr_t)Rs | (uintptr_t)Ts | (uintptr_t)t_rand) & 15) == 0;
    const unsigned grid = occ_div_up(M, kTile);
    if (bulk)
        warp_fwd_packed_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(rays, t_lin, t_rand, Rs, Ts, vol8, bbox_min, bbox_scale,
                                                                                  M, S, nb, vd, vh, vw, z, x_skel, mask);
    else
        warp_fwd_packed_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(rays, t_lin, t_rand, Rs, Ts, vol8, bbox_min, bbox_scale,
                                                                                   M, S, nb, vd, vh, vw, z, x_skel, mask);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_warp_backward_packed(const float *rays, const float *t_lin, const float *t_rand, const float *Rs,
                                            const float *Ts, const float *vol8, const float *bbox_min, const float *bbox_scale,
                                            const float *g_mask, int N, int S, int nb, int vd, int vh, int vw, float *g_vol8,
                                            float *g_Rs, float *g_Ts, occnerf_stream_t stream) {
    if (N == 0) return OCCNERF_OK;
    if (int e = check_common(rays, t_lin, Rs, Ts, N, S, nb, vd, vh, vw)) return e;
    OCC_CHECK_ARG(bbox_min && bbox_scale && g_mask && g_vol8, "warp_backward_packed: null pointer");
    OCC_CHECK_ARG((g_Rs == nullptr) == (g_Ts == nullptr), "warp_backward_packed: g_Rs and g_Ts come together");
    OCC_CHECK_ARG(!g_Rs || vol8, "warp_backward_packed: pose gradients need the packed volume");
    OCC_CHECK_ARG(((uintptr_t)g_vol8 & 15) == 0 && ((uintptr_t)vol8 & 15) == 0, "warp_backward_packed: packed buffers must be 16-byte aligned");
    const long threads = (long)N * ((S + kSeg - 1) / kSeg) * nb;
    const unsigned grid = occ_div_up(threads, kThreads);
    if (g_Rs)
        warp_bwd_packed_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(rays, t_lin, t_rand, Rs, Ts, vol8, bbox_min, bbox_scale,
                                                                                  g_mask, N, S, nb, vd, vh, vw, g_vol8, g_Rs, g_Ts);
    else
        warp_bwd_packed_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(rays, t_lin, t_rand, Rs, Ts, vol8, bbox_min, bbox_scale,
                                                                                   g_mask, N, S, nb, vd, vh, vw, g_vol8, g_Rs, g_Ts);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
