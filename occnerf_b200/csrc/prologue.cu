// Per-frame prologue of Network.forward (core/nets/occnerf/network.py:556-597): the small stages in front
// of the ray path, each as ONE kernel instead of the reference's ~60 eager launches per frame:
//
//   occnerf_pose_refine        BodyPoseRefiner (pose_decoders/mlp_delta_body_pose.py:35-41): 69 -> 256 x4 -> 69 MLP, Rodrigues
//                              (network_util.py:98-127), dst_Rs[1:] <- dst_Rs[1:] . R_delta  (network.py:558-570).  One block.
//   occnerf_motion_basis       MotionBasisComputer (network_util.py:138-200): 24-bone forward-kinematics chain along SMPL_PARENT,
//                              affine inverse, cnl_gtfms . inverse -> motion_scale_Rs, motion_Ts.  One warp, bone per lane.
//   occnerf_weight_volume_fwd  MotionWeightVolumeDecoder's tail (deconv_vol_decoder.py:25-33): softmax over the 25 channels of
//                              (decoder logits + log prior) per voxel; _bwd is its gradient to the logits.
//
// The decoder's five ConvTranspose3d stay with the library (cuDNN): 4.5 GMAC forward per frame, i.e. tensor-core work at batch 1
// for which a hand-written tcgen05 implicit GEMM is the right tool -- DESIGN.md lists it as the open part of this row.
// Forward only for the first two (their inputs carry no gradient unless the pose decoder trains; then the caller keeps the
// differentiable library path, occnerf_b200/prologue.py).
#include "common.cuh"

namespace {

constexpr int kBones = 24;
__constant__ int c_parent[kBones] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};
__constant__ int c_depth[kBones] = {0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8};

struct M34 { float m[12]; };       // rows 0..2 of an affine 4x4 (last row 0 0 0 1)

__device__ __forceinline__ M34 mul(const M34 &a, const M34 &b) {        // a . b with the k = 0..3 order of a 4x4 matmul
    M34 c;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int col = 0; col < 4; ++col) {
            float acc = a.m[r * 4 + 0] * b.m[0 * 4 + col];
            acc = fmaf(a.m[r * 4 + 1], b.m[1 * 4 + col], acc);
            acc = fmaf(a.m[r * 4 + 2], b.m[2 * 4 + col], acc);
            if (col == 3) acc += a.m[r * 4 + 3];
            c.m[r * 4 + col] = acc;
        }
    }
    return c;
}

// [A t; 0 1]^-1 = [A^-1, -A^-1 t] by cofactors (the reference calls torch.inverse; same result to fp32 rounding)
__device__ __forceinline__ M34 affine_inverse(const M34 &g) {
    const float a0 = g.m[0], a1 = g.m[1], a2 = g.m[2], b0 = g.m[4], b1 = g.m[5], b2 = g.m[6], c0 = g.m[8], c1 = g.m[9], c2 = g.m[10];
    const float i00 = b1 * c2 - b2 * c1, i01 = a2 * c1 - a1 * c2, i02 = a1 * b2 - a2 * b1;
    const float i10 = b2 * c0 - b0 * c2, i11 = a0 * c2 - a2 * c0, i12 = a2 * b0 - a0 * b2;
    const float i20 = b0 * c1 - b1 * c0, i21 = a1 * c0 - a0 * c1, i22 = a0 * b1 - a1 * b0;
    const float det = a0 * i00 + a1 * i10 + a2 * i20, s = 1.0f / det;
    M34 r;
    r.m[0] = i00 * s; r.m[1] = i01 * s; r.m[2] = i02 * s;
    r.m[4] = i10 * s; r.m[5] = i11 * s; r.m[6] = i12 * s;
    r.m[8] = i20 * s; r.m[9] = i21 * s; r.m[10] = i22 * s;
    const float tx = g.m[3], ty = g.m[7], tz = g.m[11];
    r.m[3] = -(r.m[0] * tx + r.m[1] * ty + r.m[2] * tz);
    r.m[7] = -(r.m[4] * tx + r.m[5] * ty + r.m[6] * tz);
    r.m[11] = -(r.m[8] * tx + r.m[9] * ty + r.m[10] * tz);
    return r;
}

__global__ void __launch_bounds__(32) motion_basis_kernel(const float *__restrict__ dst_Rs, const float *__restrict__ dst_Ts,
                                                          const float *__restrict__ cnl_gtfms, float *__restrict__ Rs_out,
                                                          float *__restrict__ Ts_out) {
    __shared__ M34 glob[kBones];
    const int i = threadIdx.x;
    M34 local;
    if (i < kBones) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) local.m[r * 4 + c] = __ldg(dst_Rs + i * 9 + r * 3 + c);
            local.m[r * 4 + 3] = __ldg(dst_Ts + i * 3 + r);
        }
        if (i == 0) glob[0] = local;
    }
    __syncwarp();
    // the chain in the reference's association order: glob[i] = glob[parent(i)] . local[i], one tree level at a time
    for (int d = 1; d <= 8; ++d) {
        if (i < kBones && c_depth[i] == d) glob[i] = mul(glob[c_parent[i]], local);
        __syncwarp();
    }
    if (i >= kBones) return;
    const M34 inv = affine_inverse(glob[i]);
    M34 cnl;
#pragma unroll
    for (int k = 0; k < 12; ++k) cnl.m[k] = __ldg(cnl_gtfms + i * 16 + k);
    const M34 f = mul(cnl, inv);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Rs_out[i * 9 + r * 3 + c] = f.m[r * 4 + c];
        Ts_out[i * 3 + r] = f.m[r * 4 + 3];
    }
}

struct RefinerParams { const float *w[5]; const float *b[5]; };     // 69->256, 256->256 x3, 256->69 ([out,in] row-major)

// one block of 256 threads; warp w computes output rows w, w+8, ... of every layer (lanes stride over the inputs: coalesced rows)
__global__ void __launch_bounds__(256) pose_refine_kernel(RefinerParams P, const float *__restrict__ posevec,
                                                          const float *__restrict__ dst_Rs, float *__restrict__ Rs_out) {
    __shared__ float act[2][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < 69; k += blockDim.x) act[0][k] = __ldg(posevec + k);
    __syncthreads();
    int cur = 0;
    for (int l = 0; l < 5; ++l) {
        const int in = l == 0 ? 69 : 256, out = l == 4 ? 69 : 256;
        for (int n = warp; n < out; n += 8) {
            const float *wr = P.w[l] + (size_t)n * in;
            float acc = 0.f;
            for (int k = lane; k < in; k += 32) acc = fmaf(__ldg(wr + k), act[cur][k], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                acc += __ldg(P.b[l] + n);
                act[1 - cur][n] = l < 4 ? fmaxf(acc, 0.f) : acc;
            }
        }
        __syncthreads();
        cur = 1 - cur;
    }
    // Rodrigues (network_util.py:98-127) and dst_Rs[1 + j] . R_delta[j]; the root keeps its rotation (network.py:562-570)
    if (threadIdx.x < 9) Rs_out[threadIdx.x] = __ldg(dst_Rs + threadIdx.x);
    if (threadIdx.x < 23) {
        const int j = threadIdx.x;
        float x = act[cur][3 * j], y = act[cur][3 * j + 1], z = act[cur][3 * j + 2];
        const float theta = sqrtf(1e-5f + (x * x + y * y + z * z));
        x /= theta; y /= theta; z /= theta;
        const float c = cosf(theta), s = sinf(theta);
        const float Rd[9] = {x * x + (1.f - x * x) * c, x * y * (1.f - c) - z * s, x * z * (1.f - c) + y * s,
                             x * y * (1.f - c) + z * s, y * y + (1.f - y * y) * c, y * z * (1.f - c) - x * s,
                             x * z * (1.f - c) - y * s, y * z * (1.f - c) + x * s, z * z + (1.f - z * z) * c};
        const float *R0 = dst_Rs + (j + 1) * 9;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int col = 0; col < 3; ++col) {
                float acc = __ldg(R0 + r * 3) * Rd[col];
                acc = fmaf(__ldg(R0 + r * 3 + 1), Rd[3 + col], acc);
                acc = fmaf(__ldg(R0 + r * 3 + 2), Rd[6 + col], acc);
                Rs_out[(j + 1) * 9 + r * 3 + col] = acc;
            }
    }
}

// softmax over `channels` of logits + log(prior), voxel per thread (channel stride = voxels: coalesced)
__global__ void __launch_bounds__(256) weight_volume_fwd_kernel(const float *__restrict__ logits, const float *__restrict__ priors,
                                                                int channels, long voxels, float *__restrict__ vol) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= voxels) return;
    float x[32];
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < channels; ++c) {
        x[c] = __ldg(logits + c * voxels + v) + logf(__ldg(priors + c * voxels + v));
        mx = fmaxf(mx, x[c]);
    }
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < channels; ++c) { x[c] = expf(x[c] - mx); sum += x[c]; }
#pragma unroll 1
    for (int c = 0; c < channels; ++c) vol[c * voxels + v] = x[c] / sum;
}

// d logits_c = vol_c (g_c - sum_j g_j vol_j)
__global__ void __launch_bounds__(256) weight_volume_bwd_kernel(const float *__restrict__ vol, const float *__restrict__ g_vol, int channels,
                                                                long voxels, float *__restrict__ g_logits) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= voxels) return;
    float dot = 0.f;
#pragma unroll 1
    for (int c = 0; c < channels; ++c) dot = fmaf(__ldg(g_vol + c * voxels + v), __ldg(vol + c * voxels + v), dot);
#pragma unroll 1
    for (int c = 0; c < channels; ++c) {
        const float p = __ldg(vol + c * voxels + v);
        g_logits[c * voxels + v] = p * (__ldg(g_vol + c * voxels + v) - dot);
    }
}

}  // namespace

extern "C" int occnerf_motion_basis(const float *dst_Rs, const float *dst_Ts, const float *cnl_gtfms, int n_bones, float *Rs_out,
                                    float *Ts_out, occnerf_stream_t stream) {
    OCC_CHECK_ARG(dst_Rs && dst_Ts && cnl_gtfms && Rs_out && Ts_out, "motion_basis: null pointer");
    OCC_CHECK_ARG(n_bones == kBones, "motion_basis: n_bones=%d (the SMPL tree has %d)", n_bones, kBones);
    motion_basis_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(dst_Rs, dst_Ts, cnl_gtfms, Rs_out, Ts_out);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_pose_refine(const void *const *w5_host, const void *const *b5_host, const float *posevec69, const float *dst_Rs,
                                   int n_bones, float *Rs_out, occnerf_stream_t stream) {
    OCC_CHECK_ARG(w5_host && b5_host && posevec69 && dst_Rs && Rs_out, "pose_refine: null pointer");
    OCC_CHECK_ARG(n_bones == kBones, "pose_refine: n_bones=%d (the SMPL tree has %d)", n_bones, kBones);
    RefinerParams P;
    for (int l = 0; l < 5; ++l) {
        OCC_CHECK_ARG(w5_host[l] && b5_host[l], "pose_refine: layer %d has a null pointer", l);
        P.w[l] = (const float *)w5_host[l];
        P.b[l] = (const float *)b5_host[l];
    }
    pose_refine_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(P, posevec69, dst_Rs, Rs_out);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_weight_volume_forward(const float *logits, const float *priors, int channels, long voxels, float *vol,
                                             occnerf_stream_t stream) {
    OCC_CHECK_ARG(logits && priors && vol, "weight_volume_forward: null pointer");
    OCC_CHECK_ARG(channels >= 1 && channels <= 32 && voxels >= 0, "weight_volume_forward: channels=%d voxels=%ld", channels, voxels);
    if (voxels == 0) return OCCNERF_OK;
    weight_volume_fwd_kernel<<<occ_div_up(voxels, 256), 256, 0, (cudaStream_t)stream>>>(logits, priors, channels, voxels, vol);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_weight_volume_backward(const float *vol, const float *g_vol, int channels, long voxels, float *g_logits,
                                              occnerf_stream_t stream) {
    OCC_CHECK_ARG(vol && g_vol && g_logits, "weight_volume_backward: null pointer");
    OCC_CHECK_ARG(channels >= 1 && channels <= 32 && voxels >= 0, "weight_volume_backward: channels=%d voxels=%ld", channels, voxels);
    if (voxels == 0) return OCCNERF_OK;
    weight_volume_bwd_kernel<<<occ_div_up(voxels, 256), 256, 0, (cudaStream_t)stream>>>(vol, g_vol, channels, voxels, g_logits);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
