// Loss epilogue of one training iteration: global gradient-norm clipping + Adam over ALL trainable
// tensors of the model in two multi-tensor launches.
//
// Replaces  torch.nn.utils.clip_grad_norm_(network.parameters(), 1.0); optimizer.step()  of the reference's trainer
// (core/train/trainers/occnerf/trainer.py:248-249) with torch.optim.Adam(betas=(0.9, 0.999), eps=1e-8, no weight decay, one
// learning rate per parameter tensor: core/train/optimizers/occnerf/optimizer.py:12-43).  The library path walks the 59 MiB hash
// table gradient four times (per-tensor norms, the in-place `grad *= clip_coef`, Adam) in ~40 launches; here
//   pass 1  occnerf_grad_sumsq   reads every gradient once, one double atomic per block      (4 B / parameter)
//   pass 2  occnerf_clip_adam    applies the clip coefficient ON THE FLY (gradients are not rewritten), updates exp_avg,
//                                exp_avg_sq and the parameter                                  (28 B / parameter)
// = 32 B per parameter, the minimum for dense Adam.  Entries whose gradient and both moments are exactly zero -- hash-table
// slots no sample has ever touched, the 7/8 of the first ConvTranspose3d's taps that see no input voxel -- are skipped after
// reading them (their update is exactly 0 in the reference as well), which saves their 16 B of parameter read + three writes.
// Everything the kernels need beyond the tensors (the squared norm, the step counter) lives in device memory, so the step is
// CUDA-graph capturable and needs no host synchronisation.
//
// The tensor list travels as a kernel PARAMETER (<= kMaxTensors entries, 40 B each): gradient buffers are re-allocated by
// autograd every iteration, so a table cached in device memory would go stale.
#include "common.cuh"

namespace {

constexpr int kMaxTensors = 64;          // 64 x 52 B + prefix table = 3.6 KB of kernel parameters
constexpr int kChunk = 4096;             // elements per block

struct TensorList {
    float *p[kMaxTensors];
    const float *g[kMaxTensors];
    float *m[kMaxTensors];
    float *v[kMaxTensors];
    float *step[kMaxTensors];            // per-tensor step count (torch keeps one per parameter: a tensor without gradient in
                                         // some iteration is skipped and its bias correction lags behind)
    int chunk_begin[kMaxTensors + 1];    // prefix sums of ceil(numel / kChunk)
    int numel[kMaxTensors];
    float lr[kMaxTensors];
    int n;
};

__device__ __forceinline__ int find_tensor(const TensorList &L, int chunk) {
    int lo = 0, hi = L.n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (L.chunk_begin[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const __grid_constant__ TensorList L, double *sumsq) {
    const int t = find_tensor(L, blockIdx.x);
    const int base = (blockIdx.x - L.chunk_begin[t]) * kChunk;
    const int n = min(kChunk, L.numel[t] - base);
    const float *g = L.g[t] + base;
    float acc = 0.f;
    if ((((uintptr_t)g) & 15) == 0) {
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
        for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
            const float4 x = __ldg(g4 + i);
            acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
        for (int i = (n / 4) * 4 + threadIdx.x; i < n; i += blockDim.x) { const float x = __ldg(g + i); acc += x * x; }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) { const float x = __ldg(g + i); acc += x * x; }
    }
    double d = (double)warp_sum(acc);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += part[w];
        if (s != 0.0) atomicAdd(sumsq, s);
    }
}

struct AdamHyper { float beta1, beta2, eps, max_norm; };

__global__ void __launch_bounds__(256) clip_adam_kernel(const __grid_constant__ TensorList L, const double *__restrict__ sumsq, AdamHyper h) {
    const int t = find_tensor(L, blockIdx.x);
    const int base = (blockIdx.x - L.chunk_begin[t]) * kChunk;
    const int n = min(kChunk, L.numel[t] - base);
    // clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
    const float norm = (float)sqrt(*sumsq);
    const float coef = h.max_norm > 0.f ? fminf(h.max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
    // torch.optim.Adam (single-tensor formulation): step_size = lr / (1 - beta1^t); denom = sqrt(v) / sqrt(1 - beta2^t) + eps
    const double tstep = (double)*L.step[t];
    const float bc1 = 1.0f - (float)pow((double)h.beta1, tstep);
    const float bc2_sqrt = sqrtf(1.0f - (float)pow((double)h.beta2, tstep));
    const float step_size = L.lr[t] / bc1;
    float *p = L.p[t] + base, *m = L.m[t] + base, *v = L.v[t] + base;
    const float *g = L.g[t] + base;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float gi = __ldg(g + i) * coef;
        float mi = m[i], vi = v[i];
        if (gi == 0.f && mi == 0.f && vi == 0.f) continue;          // exact no-op in the reference as well
        mi = mi + (1.0f - h.beta1) * (gi - mi);                     // torch: exp_avg.lerp_(grad, 1 - beta1)
        vi = vi * h.beta2 + (1.0f - h.beta2) * gi * gi;             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        m[i] = mi;
        v[i] = vi;
        p[i] = p[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + h.eps));
    }
}

}  // namespace

// One optimisation step over n tensors (host arrays of device pointers).  sumsq: one double in device memory (squared gradient
// norm of this step); every listed tensor's step count is incremented FIRST, as torch does.  max_norm <= 0 disables the clipping.
__global__ void adam_count_kernel(const __grid_constant__ TensorList L, double *sumsq, int clear) {
    if (clear && threadIdx.x == 0) *sumsq = 0.0;
    if ((int)threadIdx.x < L.n) *L.step[threadIdx.x] += 1.0f;
}

extern "C" int occnerf_clip_adam_step(void *const *params, const void *const *grads, void *const *exp_avg, void *const *exp_avg_sq,
                                      void *const *steps, const long *numel, const float *lr, int n, float beta1, float beta2, float eps,
                                      float max_norm, double *sumsq, occnerf_stream_t stream) {
    OCC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && steps && numel && lr && sumsq, "clip_adam_step: null pointer");
    OCC_CHECK_ARG(n >= 0, "clip_adam_step: n=%d", n);
    cudaStream_t st = (cudaStream_t)stream;
    double *state2 = sumsq;
    AdamHyper h = {beta1, beta2, eps, max_norm};
    bool cleared = false;
    // (the squared norm spans ALL tensors, so every batch of the list runs pass 1 before any batch runs pass 2)
    for (int pass = -1; pass < 2; ++pass) {                 // -1: step counters (+ clear the norm), 0: norm, 1: update
        for (int t0 = 0; t0 < n; t0 += kMaxTensors) {
            TensorList L;
            L.n = 0;
            int chunks = 0;
            for (int t = t0; t < n && t < t0 + kMaxTensors; ++t) {
                if (numel[t] == 0) continue;
                OCC_CHECK_ARG(params[t] && grads[t] && exp_avg[t] && exp_avg_sq[t], "clip_adam_step: tensor %d has a null pointer", t);
                OCC_CHECK_ARG(numel[t] > 0 && numel[t] < (1l << 31), "clip_adam_step: tensor %d has %ld elements", t, numel[t]);
                const int k = L.n++;
                L.p[k] = (float *)params[t]; L.g[k] = (const float *)grads[t]; L.m[k] = (float *)exp_avg[t]; L.v[k] = (float *)exp_avg_sq[t];
                OCC_CHECK_ARG(steps[t], "clip_adam_step: tensor %d has no step counter", t);
                L.step[k] = (float *)steps[t];
                L.numel[k] = (int)numel[t]; L.lr[k] = lr[t];
                L.chunk_begin[k] = chunks;
                chunks += (int)((numel[t] + kChunk - 1) / kChunk);
            }
            L.chunk_begin[L.n] = chunks;
            if (chunks == 0) continue;
            if (pass < 0) { adam_count_kernel<<<1, kMaxTensors, 0, st>>>(L, state2, cleared ? 0 : 1); cleared = true; }
            else if (pass == 0) grad_sumsq_kernel<<<chunks, 256, 0, st>>>(L, state2);
            else clip_adam_kernel<<<chunks, 256, 0, st>>>(L, state2, h);
            OCC_LAUNCH_CHECK();
        }
    }
    return OCCNERF_OK;
}
