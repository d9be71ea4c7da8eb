// K2: multi-resolution hash-grid encode, scatter-add backward and input backward.
//
// Drop-in for the reference's `_gridencoder` operator (core/nets/occnerf/gridencoder/src/gridencoder.h:12-15,
// kernels gridencoder.cu:87-369) for gridtype=hash, align_corners=False, linear interpolation -- the only
// configuration OccNeRF instantiates (canonical_mlps/occnerf_mlp.py:45: D=4, L=16, C=2, H=16, T=2^19).
// Table slots are bit-exact with the reference: same uint32 prime/XOR hash, same dense-stride walk, same
// modulo, same FMA-contracted `x*scale + 0.5`, and the per-level scale table is evaluated on the device
// with the reference's own expression (occnerf_hashgrid_level_scales).
//
// What is different, by design:
//  * thread = (sample, level) with the level fastest, so a warp writes 2 x 128 B fully coalesced rows of
//    the [B, L*C] output directly (the reference writes [L,B,C] and pays a transpose copy, grid.py:58,76);
//    the row stride `ld` lets the caller encode straight into the MLP input buffer.
//  * backward uses red.global.add.v2.f32: one L2 reduction per corner instead of C scalar REDs
//    (the reference compiled for sm_100a emits 32 scalar RED.E.ADD.F32 per thread).
//  * the whole 59 MiB table (and its gradient) is L2-resident on B200 (126 MB L2); no level-major scheduling
//    is needed to keep gathers out of HBM.
// Algorithmic traffic per sample (D=4,L=16,C=2): forward 16 B in + 128 B out (HBM), 2048 B gathered (L2);
// backward 128 B in, 256 x 8 B reductions (L2).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kOob = 0xFFFFFFFFu;

template <uint32_t D>
__device__ __forceinline__ uint32_t hash_coords(const uint32_t (&g)[D]) {
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) r ^= g[d] * primes[d];
    return r;
}

// slot of a grid vertex inside its level's table (gridencoder.cu:66-84 with gridtype=0, align_corners=false)
template <uint32_t D>
__device__ __forceinline__ uint32_t cell_slot(const uint32_t (&g)[D], uint32_t hashmap_size, uint32_t resolution) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if (stride <= hashmap_size) {
            index += g[d] * stride;
            stride *= resolution + 1;
        }
    }
    if (stride > hashmap_size) index = hash_coords<D>(g);
    return (hashmap_size & (hashmap_size - 1)) == 0 ? (index & (hashmap_size - 1)) : (index % hashmap_size);
}

template <uint32_t D>
struct Located {
    float frac[D];
    uint32_t g[D];
    bool inside;
};

template <uint32_t D>
__device__ __forceinline__ Located<D> locate(const float *__restrict__ x, float scale) {
    Located<D> r;
    r.inside = true;
    float xv[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        xv[d] = __ldg(x + d);
        if (xv[d] < 0.0f || xv[d] > 1.0f) r.inside = false;
    }
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const float p = __fmaf_rn(xv[d], scale, 0.5f);
        const float fl = floorf(p);
        r.g[d] = (uint32_t)fl;
        r.frac[d] = p - (float)r.g[d];
    }
    return r;
}

__global__ void level_scales_kernel(float S, uint32_t H, uint32_t L, float *__restrict__ out) {
    const uint32_t level = blockIdx.x * blockDim.x + threadIdx.x;
    if (level >= L) return;
    // the reference's expression, gridencoder.cu:138 (same operand types, same contraction opportunities)
    const float scale = exp2f(level * S) * H - 1.0f;
    out[level] = scale;
}

template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(kThreads)
hashgrid_fwd_kernel(const float *__restrict__ inputs, const float *__restrict__ emb, const int32_t *__restrict__ offsets,
                    const float *__restrict__ scales, float *__restrict__ outputs, int layout, long ld, uint32_t B,
                    uint32_t L, float *__restrict__ dy_dx, uint32_t *__restrict__ cells, uint32_t *__restrict__ slots) {
    const unsigned long long gid = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    if (gid >= (unsigned long long)B * L) return;
    const uint32_t b = (uint32_t)(gid / L), level = (uint32_t)(gid - (unsigned long long)b * L);
    const uint32_t off = (uint32_t)__ldg(offsets + level);
    const uint32_t hashmap_size = (uint32_t)__ldg(offsets + level + 1) - off;
    const float scale = __ldg(scales + level);
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
    const float *grid = emb + (size_t)off * C;
    float *out = layout == OCCNERF_LAYOUT_LBC ? outputs + ((size_t)level * B + b) * C : outputs + (size_t)b * ld + level * C;
    constexpr uint32_t NC = 1u << D;
    const Located<D> loc = locate<D>(inputs + (size_t)b * D, scale);
    if (!loc.inside) {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) out[c] = 0.0f;
        if (dy_dx) {
            float *dd = dy_dx + ((size_t)b * L + level) * D * C;
            for (uint32_t i = 0; i < D * C; ++i) dd[i] = 0.0f;
        }
        if (cells) for (uint32_t d = 0; d < D; ++d) cells[((size_t)b * L + level) * D + d] = kOob;
        if (slots) for (uint32_t k = 0; k < NC; ++k) slots[((size_t)b * L + level) * NC + k] = kOob;
        return;
    }
    if (cells) for (uint32_t d = 0; d < D; ++d) cells[((size_t)b * L + level) * D + d] = loc.g[d];
    float acc[C];
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) acc[c] = 0.0f;
#pragma unroll
    for (uint32_t k = 0; k < NC; ++k) {
        float w = 1.0f;
        uint32_t gl[D];
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            if ((k & (1u << d)) == 0) { w *= 1.0f - loc.frac[d]; gl[d] = loc.g[d]; }
            else                      { w *= loc.frac[d];        gl[d] = loc.g[d] + 1; }
        }
        const uint32_t slot = cell_slot<D>(gl, hashmap_size, resolution);
        if (slots) slots[((size_t)b * L + level) * NC + k] = slot;
        const float *e = grid + (size_t)slot * C;
        if constexpr (C == 2) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(e));
            acc[0] = __fmaf_rn(w, v.x, acc[0]);
            acc[1] = __fmaf_rn(w, v.y, acc[1]);
        } else if constexpr (C == 4 || C == 8) {
#pragma unroll
            for (uint32_t c4 = 0; c4 < C; c4 += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(e + c4));
                acc[c4 + 0] = __fmaf_rn(w, v.x, acc[c4 + 0]);
                acc[c4 + 1] = __fmaf_rn(w, v.y, acc[c4 + 1]);
                acc[c4 + 2] = __fmaf_rn(w, v.z, acc[c4 + 2]);
                acc[c4 + 3] = __fmaf_rn(w, v.w, acc[c4 + 3]);
            }
        } else {
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) acc[c] = __fmaf_rn(w, __ldg(e + c), acc[c]);
        }
    }
    if constexpr (C == 2) {
        *reinterpret_cast<float2 *>(out) = make_float2(acc[0], acc[1]);
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) out[c] = acc[c];
    }
    if (!dy_dx) return;
    // d out / d x  (gridencoder.cu:201-244): finite difference along gd, multilinear in the other dims
    float *dd = dy_dx + ((size_t)b * L + level) * D * C;
#pragma unroll
    for (uint32_t gd = 0; gd < D; ++gd) {
        float ga[C];
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) ga[c] = 0.0f;
#pragma unroll
        for (uint32_t k = 0; k < (1u << (D - 1)); ++k) {
            float w = scale;
            uint32_t gl[D];
#pragma unroll
            for (uint32_t nd = 0; nd < D - 1; ++nd) {
                const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                if ((k & (1u << nd)) == 0) { w *= 1.0f - loc.frac[d]; gl[d] = loc.g[d]; }
                else                       { w *= loc.frac[d];        gl[d] = loc.g[d] + 1; }
            }
            gl[gd] = loc.g[gd];
            const uint32_t sl = cell_slot<D>(gl, hashmap_size, resolution);
            gl[gd] = loc.g[gd] + 1;
            const uint32_t sr = cell_slot<D>(gl, hashmap_size, resolution);
#pragma unroll
            for (uint32_t c = 0; c < C; ++c)
                ga[c] += w * (__ldg(grid + (size_t)sr * C + c) - __ldg(grid + (size_t)sl * C + c));
        }
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) dd[gd * C + c] = ga[c];
    }
}

template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(kThreads)
hashgrid_bwd_kernel(const float *__restrict__ grad, int layout, long ld, const float *__restrict__ inputs,
                    const int32_t *__restrict__ offsets, const float *__restrict__ scales, float *__restrict__ g_emb,
                    uint32_t B, uint32_t L) {
    const unsigned long long gid = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    if (gid >= (unsigned long long)B * L) return;
    const uint32_t b = (uint32_t)(gid / L), level = (uint32_t)(gid - (unsigned long long)b * L);
    const uint32_t off = (uint32_t)__ldg(offsets + level);
    const uint32_t hashmap_size = (uint32_t)__ldg(offsets + level + 1) - off;
    const float scale = __ldg(scales + level);
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
    const Located<D> loc = locate<D>(inputs + (size_t)b * D, scale);
    if (!loc.inside) return;
    const float *gp = layout == OCCNERF_LAYOUT_LBC ? grad + ((size_t)level * B + b) * C : grad + (size_t)b * ld + level * C;
    float g[C];
    bool any = false;
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) { g[c] = __ldg(gp + c); any |= g[c] != 0.0f; }
    if (!any) return;
    float *gg = g_emb + (size_t)off * C;
    constexpr uint32_t NC = 1u << D;
#pragma unroll
    for (uint32_t k = 0; k < NC; ++k) {
        float w = 1.0f;
        uint32_t gl[D];
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            if ((k & (1u << d)) == 0) { w *= 1.0f - loc.frac[d]; gl[d] = loc.g[d]; }
            else                      { w *= loc.frac[d];        gl[d] = loc.g[d] + 1; }
        }
        float *dst = gg + (size_t)cell_slot<D>(gl, hashmap_size, resolution) * C;
        if constexpr (C == 2) {
            red_add_v2(dst, w * g[0], w * g[1]);
        } else if constexpr (C == 4 || C == 8) {
#pragma unroll
            for (uint32_t c4 = 0; c4 < C; c4 += 4) red_add_v4(dst + c4, w * g[c4], w * g[c4 + 1], w * g[c4 + 2], w * g[c4 + 3]);
        } else {
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) atomicAdd(dst + c, w * g[c]);
        }
    }
}

// Run-length variant of the gather for inputs that are ordered along rays: thread = (segment of SEG consecutive samples,
// level), level fastest (a warp still writes 2 x 128 B coalesced row pieces per step).  The 2^D corner vectors of the current
// cell stay in registers and are re-fetched only when the sample moves to another cell: 1.5-4.5x fewer L2 gathers.
// Bitwise the same outputs as hashgrid_fwd_kernel (same slots, same FMA order).
template <uint32_t D, uint32_t C, uint32_t SEG>
__global__ void __launch_bounds__(kThreads)
hashgrid_fwd_runs_kernel(const float *__restrict__ inputs, const float *__restrict__ emb, const int32_t *__restrict__ offsets,
                         const float *__restrict__ scales, float *__restrict__ outputs, int layout, long ld, uint32_t B, uint32_t L) {
    const unsigned long long gid = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    const uint32_t nseg = (B + SEG - 1) / SEG;
    if (gid >= (unsigned long long)nseg * L) return;
    const uint32_t seg = (uint32_t)(gid / L), level = (uint32_t)(gid - (unsigned long long)seg * L);
    const uint32_t off = (uint32_t)__ldg(offsets + level);
    const uint32_t hashmap_size = (uint32_t)__ldg(offsets + level + 1) - off;
    const float scale = __ldg(scales + level);
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
    const float *grid = emb + (size_t)off * C;
    constexpr uint32_t NC = 1u << D;
    uint32_t cur[D];
    float val[NC][C];
    bool have = false;
    const uint32_t b0 = seg * SEG, b1 = min(B, b0 + SEG);
    for (uint32_t b = b0; b < b1; ++b) {
        float *out = layout == OCCNERF_LAYOUT_LBC ? outputs + ((size_t)level * B + b) * C : outputs + (size_t)b * ld + level * C;
        const Located<D> loc = locate<D>(inputs + (size_t)b * D, scale);
        float acc[C];
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) acc[c] = 0.0f;
        if (loc.inside) {
            bool same = have;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) same = same && (loc.g[d] == cur[d]);
            if (!same) {
                have = true;
#pragma unroll
                for (uint32_t d = 0; d < D; ++d) cur[d] = loc.g[d];
#pragma unroll
                for (uint32_t k = 0; k < NC; ++k) {
                    uint32_t gl[D];
#pragma unroll
                    for (uint32_t d = 0; d < D; ++d) gl[d] = cur[d] + ((k >> d) & 1u);
                    const float *e = grid + (size_t)cell_slot<D>(gl, hashmap_size, resolution) * C;
                    if constexpr (C == 2) {
                        const float2 v = __ldg(reinterpret_cast<const float2 *>(e));
                        val[k][0] = v.x; val[k][1] = v.y;
                    } else {
#pragma unroll
                        for (uint32_t c = 0; c < C; ++c) val[k][c] = __ldg(e + c);
                    }
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < NC; ++k) {
                float w = 1.0f;
#pragma unroll
                for (uint32_t d = 0; d < D; ++d) w *= (k & (1u << d)) ? loc.frac[d] : 1.0f - loc.frac[d];
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) acc[c] = __fmaf_rn(w, val[k][c], acc[c]);
            }
        }
        if constexpr (C == 2) {
            *reinterpret_cast<float2 *>(out) = make_float2(acc[0], acc[1]);
        } else {
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) out[c] = acc[c];
        }
    }
}

// Run-length variant of the scatter for inputs that are ordered along rays (what the render path produces): thread =
// (segment of SEG consecutive samples, level), level fastest.  Neighbouring samples fall into the same grid cell at the
// coarse levels (and ~1/3 of the time even at the finest), so the thread accumulates the 2^D corner gradients of the
// current cell in registers and issues its reductions only when the cell changes: 1.5-4.5x fewer L2 atomics on the
// addresses that are contended the most.  Mathematically the same sum as hashgrid_bwd_kernel (fp32 addition order differs,
// as it does between any two runs of an atomic scatter).
template <uint32_t D, uint32_t C, uint32_t SEG>
__global__ void __launch_bounds__(kThreads)
hashgrid_bwd_runs_kernel(const float *__restrict__ grad, int layout, long ld, const float *__restrict__ inputs,
                         const int32_t *__restrict__ offsets, const float *__restrict__ scales, float *__restrict__ g_emb,
                         uint32_t B, uint32_t L) {
    const unsigned long long gid = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    const uint32_t nseg = (B + SEG - 1) / SEG;
    if (gid >= (unsigned long long)nseg * L) return;
    const uint32_t seg = (uint32_t)(gid / L), level = (uint32_t)(gid - (unsigned long long)seg * L);
    const uint32_t off = (uint32_t)__ldg(offsets + level);
    const uint32_t hashmap_size = (uint32_t)__ldg(offsets + level + 1) - off;
    const float scale = __ldg(scales + level);
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
    float *gg = g_emb + (size_t)off * C;
    constexpr uint32_t NC = 1u << D;
    uint32_t cur[D];
    float acc[NC][C];
    bool have = false;
    auto flush = [&]() {
#pragma unroll
        for (uint32_t k = 0; k < NC; ++k) {
            uint32_t gl[D];
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) gl[d] = cur[d] + ((k >> d) & 1u);
            float *dst = gg + (size_t)cell_slot<D>(gl, hashmap_size, resolution) * C;
            if constexpr (C == 2) {
                red_add_v2(dst, acc[k][0], acc[k][1]);
            } else {
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) atomicAdd(dst + c, acc[k][c]);
            }
        }
    };
    const uint32_t b0 = seg * SEG, b1 = min(B, b0 + SEG);
    for (uint32_t b = b0; b < b1; ++b) {
        const Located<D> loc = locate<D>(inputs + (size_t)b * D, scale);
        if (!loc.inside) continue;
        const float *gp = layout == OCCNERF_LAYOUT_LBC ? grad + ((size_t)level * B + b) * C : grad + (size_t)b * ld + level * C;
        float g[C];
        bool any = false;
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) { g[c] = __ldg(gp + c); any |= g[c] != 0.0f; }
        if (!any) continue;
        bool same = have;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) same = same && (loc.g[d] == cur[d]);
        if (!same) {
            if (have) flush();
            have = true;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) cur[d] = loc.g[d];
#pragma unroll
            for (uint32_t k = 0; k < NC; ++k)
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) acc[k][c] = 0.0f;
        }
#pragma unroll
        for (uint32_t k = 0; k < NC; ++k) {
            float w = 1.0f;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) w *= (k & (1u << d)) ? loc.frac[d] : 1.0f - loc.frac[d];
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) acc[k][c] = __fmaf_rn(w, g[c], acc[k][c]);
        }
    }
    if (have) flush();
}

__global__ void __launch_bounds__(kThreads)
hashgrid_input_bwd_kernel(const float *__restrict__ grad, int layout, long ld, const float *__restrict__ dy_dx,
                          float *__restrict__ g_in, uint32_t B, uint32_t D, uint32_t C, uint32_t L) {
    const unsigned long long t = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= (unsigned long long)B * D) return;
    const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (unsigned long long)b * D);
    float r = 0.0f;
    for (uint32_t l = 0; l < L; ++l)
        for (uint32_t c = 0; c < C; ++c) {
            const float go = layout == OCCNERF_LAYOUT_LBC ? __ldg(grad + ((size_t)l * B + b) * C + c)
                                                         : __ldg(grad + (size_t)b * ld + l * C + c);
            r += go * __ldg(dy_dx + (((size_t)b * L + l) * D + d) * C + c);
        }
    g_in[(size_t)b * D + d] = r;
}

template <uint32_t D, uint32_t C>
int launch_fwd(const float *inputs, const float *emb, const int32_t *offsets, const float *scales, float *outputs,
               int layout, long ld, uint32_t B, uint32_t L, float *dy_dx, uint32_t *cells, uint32_t *slots, int run_length,
               cudaStream_t st) {
    if (run_length == 16 && !dy_dx && !cells && !slots) {
        const long nseg = ((long)B + 15) / 16;
        hashgrid_fwd_runs_kernel<D, C, 16><<<occ_div_up(nseg * L, kThreads), kThreads, 0, st>>>(inputs, emb, offsets, scales, outputs,
                                                                                              layout, ld, B, L);
        OCC_LAUNCH_CHECK();
        return OCCNERF_OK;
    }
    hashgrid_fwd_kernel<D, C><<<occ_div_up((long)B * L, kThreads), kThreads, 0, st>>>(
        inputs, emb, offsets, scales, outputs, layout, ld, B, L, dy_dx, cells, slots);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
template <uint32_t D, uint32_t C>
int launch_bwd(const float *grad, int layout, long ld, const float *inputs, const int32_t *offsets,
               const float *scales, float *g_emb, uint32_t B, uint32_t L, int run_length, cudaStream_t st) {
    if (run_length == 8 || run_length == 16) {
        const long nseg = ((long)B + run_length - 1) / run_length;
        if (run_length == 8)
            hashgrid_bwd_runs_kernel<D, C, 8><<<occ_div_up(nseg * L, kThreads), kThreads, 0, st>>>(grad, layout, ld, inputs, offsets, scales, g_emb, B, L);
        else
            hashgrid_bwd_runs_kernel<D, C, 16><<<occ_div_up(nseg * L, kThreads), kThreads, 0, st>>>(grad, layout, ld, inputs, offsets, scales, g_emb, B, L);
        OCC_LAUNCH_CHECK();
        return OCCNERF_OK;
    }
    hashgrid_bwd_kernel<D, C><<<occ_div_up((long)B * L, kThreads), kThreads, 0, st>>>(grad, layout, ld, inputs, offsets,
                                                                                     scales, g_emb, B, L);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

#define OCC_DISPATCH_DC(D, C, CALL)                                                     \
    switch ((D) * 16 + (C)) {                                                            \
        case 2 * 16 + 1: return CALL(2, 1);                                              \
        case 2 * 16 + 2: return CALL(2, 2);                                              \
        case 2 * 16 + 4: return CALL(2, 4);                                              \
        case 2 * 16 + 8: return CALL(2, 8);                                              \
        case 3 * 16 + 1: return CALL(3, 1);                                              \
        case 3 * 16 + 2: return CALL(3, 2);                                              \
        case 3 * 16 + 4: return CALL(3, 4);                                              \
        case 3 * 16 + 8: return CALL(3, 8);                                              \
        case 4 * 16 + 1: return CALL(4, 1);                                              \
        case 4 * 16 + 2: return CALL(4, 2);                                              \
        case 4 * 16 + 4: return CALL(4, 4);                                              \
        case 4 * 16 + 8: return CALL(4, 8);                                              \
        default:                                                                         \
            occnerf_set_error("hashgrid: unsupported D=%u C=%u (D in {2,3,4}, C in {1,2,4,8})", (D), (C)); \
            return OCCNERF_EINVAL;                                                       \
    }

int check_layout(int layout, int ld, uint32_t L, uint32_t C) {
    OCC_CHECK_ARG(layout == OCCNERF_LAYOUT_BLC || layout == OCCNERF_LAYOUT_LBC, "hashgrid: bad layout %d", layout);
    OCC_CHECK_ARG(layout == OCCNERF_LAYOUT_LBC || ld >= (int)(L * C), "hashgrid: ld=%d < L*C=%u", ld, L * C);
    return 0;
}

}  // namespace

extern "C" int occnerf_hashgrid_level_scales(float S, uint32_t H, uint32_t L, float *level_scales,
                                             occnerf_stream_t stream) {
    OCC_CHECK_ARG(level_scales && L >= 1 && L <= 64, "hashgrid_level_scales: bad arguments");
    level_scales_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(S, H, L, level_scales);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}

extern "C" int occnerf_hashgrid_forward(const float *inputs, const float *embeddings, const int32_t *offsets,
                                        const float *level_scales, float *outputs, int layout, int ld, uint32_t B,
                                        uint32_t D, uint32_t C, uint32_t L, float *dy_dx, uint32_t *cells,
                                        uint32_t *slots, int run_length, occnerf_stream_t stream) {
    if (B == 0 && D >= 2 && D <= 4) return OCCNERF_OK;
    OCC_CHECK_ARG(D >= 2 && D <= 4, "hashgrid: unsupported D=%u C=%u (D in {2,3,4}, C in {1,2,4,8})", D, C);
    OCC_CHECK_ARG(inputs && embeddings && offsets && level_scales && outputs, "hashgrid_forward: null pointer");
    OCC_CHECK_ARG(run_length == 0 || run_length == 16, "hashgrid_forward: run_length=%d (0 or 16)", run_length);
    if (int e = check_layout(layout, ld, L, C)) return e;
    OCC_CHECK_ARG(C != 2 || layout == OCCNERF_LAYOUT_LBC || (ld % 2 == 0 && ((uintptr_t)outputs & 7) == 0),
                  "hashgrid_forward: C=2 output rows must be 8-byte aligned");
    if (B == 0) return OCCNERF_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD, CC) launch_fwd<DD, CC>(inputs, embeddings, offsets, level_scales, outputs, layout, ld, B, L, dy_dx, cells, slots, run_length, st)
    OCC_DISPATCH_DC(D, C, CALL)
#undef CALL
}

extern "C" int occnerf_hashgrid_backward(const float *grad, int layout, int ld, const float *inputs,
                                         const int32_t *offsets, const float *level_scales, float *grad_embeddings,
                                         uint32_t B, uint32_t D, uint32_t C, uint32_t L, int run_length,
                                         occnerf_stream_t stream) {
    if (B == 0) return OCCNERF_OK;
    OCC_CHECK_ARG(grad && inputs && offsets && level_scales && grad_embeddings, "hashgrid_backward: null pointer");
    OCC_CHECK_ARG(run_length == 0 || run_length == 8 || run_length == 16, "hashgrid_backward: run_length=%d (0, 8 or 16)", run_length);
    if (int e = check_layout(layout, ld, L, C)) return e;
    if (B == 0) return OCCNERF_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD, CC) launch_bwd<DD, CC>(grad, layout, ld, inputs, offsets, level_scales, grad_embeddings, B, L, run_length, st)
    OCC_DISPATCH_DC(D, C, CALL)
#undef CALL
}

extern "C" int occnerf_hashgrid_input_backward(const float *grad, int layout, int ld, const float *dy_dx,
                                               float *grad_inputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                                               occnerf_stream_t stream) {
    OCC_CHECK_ARG(grad && dy_dx && grad_inputs, "hashgrid_input_backward: null pointer");
    if (int e = check_layout(layout, ld, L, C)) return e;
    if (B == 0) return OCCNERF_OK;
    hashgrid_input_bwd_kernel<<<occ_div_up((long)B * D, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        grad, layout, ld, dy_dx, grad_inputs, B, D, C, L);
    OCC_LAUNCH_CHECK();
    return OCCNERF_OK;
}
