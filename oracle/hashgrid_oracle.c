/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * CPU restatement of the reference's multi-resolution hash-grid operator
 *   core/nets/occnerf/gridencoder/src/gridencoder.cu
 *     :50-63   fast_hash            (uint32 prime multiply + XOR)
 *     :66-84   get_grid_index       (dense stride walk, hash fallback, modulo)
 *     :99-197  kernel_grid          (forward gather, 2^D corners)
 *     :201-244 kernel_grid dy_dx    (d output / d input)
 *     :249-340 kernel_grid_backward (scatter-add into grad table)
 *     :343-369 kernel_input_backward
 * The reference has no CPU path (CHECK_CUDA at :449-452); this file is what the
 * CUDA kernels in occnerf_b200/csrc/hashgrid.cu are checked against.
 *
 * Arithmetic notes (SURVEY.md section 7 "hard parts"):
 *  - pos = x*scale + 0.5 is contracted to one FMA by nvcc, so it is computed
 *    here with fmaf() (single rounding).
 *  - scale = exp2f(level*S)*H - 1 uses the *device* exp2f in the reference; a
 *    caller may pass the 16-entry table read back from the device
 *    (level_scales != NULL) so that integer indices can be compared bit for bit.
 *    With level_scales == NULL the host exp2f is used.
 *  - weights are accumulated exactly in the reference's order
 *    (w = 1; for d: w *= (bit ? f : 1-f)), result += w*value via FMA.
 *
 * Parity pinning: the reference ships no golden vectors for this operator
 * (SURVEY.md section 8c).  Pinned instead against the reference kernel itself,
 * compiled unmodified into oracle/_ref/ and run on the GPU box
 * (tests/test_hashgrid_gpu.py::test_oracle_vs_compiled_reference).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <pthread.h>

/* threads used by the b-loops below; 1 = deterministic sample order (tests), >1 = CPU-baseline timing.
 * Plain pthreads: the process already hosts torch's OpenMP runtime and a second one is best avoided. */
static int g_threads = 1;
void hg_set_threads(int n) { g_threads = n > 0 ? (n > 256 ? 256 : n) : 1; }

typedef void (*range_fn)(uint32_t b0, uint32_t b1, void *ctx);
typedef struct { range_fn fn; uint32_t b0, b1; void *ctx; } range_job;
static void *range_tramp(void *p) { range_job *j = (range_job *)p; j->fn(j->b0, j->b1, j->ctx); return NULL; }
static void parallel_ranges(uint32_t B, range_fn fn, void *ctx) {
    int nt = g_threads;
    if (nt <= 1 || B < 1024) { fn(0, B, ctx); return; }
    pthread_t th[256];
    range_job jobs[256];
    uint32_t per = (B + (uint32_t)nt - 1) / (uint32_t)nt;
    int started = 0;
    for (int t = 0; t < nt; ++t) {
        uint32_t b0 = (uint32_t)t * per, b1 = b0 + per > B ? B : b0 + per;
        if (b0 >= B) break;
        jobs[t].fn = fn; jobs[t].b0 = b0; jobs[t].b1 = b1; jobs[t].ctx = ctx;
        pthread_create(&th[t], NULL, range_tramp, &jobs[t]);
        ++started;
    }
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

static void atomic_add_f32(float *addr, float v) {
    if (g_threads <= 1) { *addr += v; return; }
    uint32_t *ua = (uint32_t *)addr;
    uint32_t old = __atomic_load_n(ua, __ATOMIC_RELAXED), neu;
    do { float f; memcpy(&f, &old, 4); f += v; memcpy(&neu, &f, 4);
    } while (!__atomic_compare_exchange_n(ua, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
static void atomic_add_f64(double *addr, double v) {
    if (g_threads <= 1) { *addr += v; return; }
    uint64_t *ua = (uint64_t *)addr;
    uint64_t old = __atomic_load_n(ua, __ATOMIC_RELAXED), neu;
    do { double f; memcpy(&f, &old, 8); f += v; memcpy(&neu, &f, 8);
    } while (!__atomic_compare_exchange_n(ua, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

#define MAX_D 5

static const uint32_t PRIMES[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                   2097192037u, 1434869437u, 2165219737u};

static uint32_t hash_cell(uint32_t D, const uint32_t *g) {
    uint32_t r = 0;
    for (uint32_t d = 0; d < D; ++d) r ^= g[d] * PRIMES[d];
    return r;
}

/* element index (already multiplied by C, channel 0) of one grid cell */
static uint32_t cell_index(uint32_t D, uint32_t C, uint32_t gridtype, int align_corners,
                           uint32_t hashmap_size, uint32_t resolution, const uint32_t *g) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; ++d) {
        index += g[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) index = hash_cell(D, g);
    return (index % hashmap_size) * C;
}

static float level_scale(uint32_t level, float S, uint32_t H, const float *level_scales) {
    if (level_scales) return level_scales[level];
    return exp2f((float)level * S) * (float)H - 1.0f;
}

void hg_host_level_scales(float S, uint32_t H, uint32_t L, float *out) {
    for (uint32_t l = 0; l < L; ++l) out[l] = level_scale(l, S, H, NULL);
}

static int locate(uint32_t D, const float *x, float scale, int align_corners, int interp,
                  float *frac, float *deriv, uint32_t *g) {
    for (uint32_t d = 0; d < D; ++d)
        if (x[d] < 0.0f || x[d] > 1.0f) return 0;
    for (uint32_t d = 0; d < D; ++d) {
        float p = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
        float fl = floorf(p);
        g[d] = (uint32_t)fl;
        p -= (float)g[d];
        if (interp == 1) {
            deriv[d] = 6.0f * p * (1.0f - p);
            p = p * p * (3.0f - 2.0f * p);
        } else {
            deriv[d] = 1.0f;
        }
        frac[d] = p;
    }
    return 1;
}


typedef struct {
    const float *inputs, *emb, *level_scales, *grad;
    const int32_t *offsets;
    float *outputs, *dy_dx, *grad_emb, *grad_inputs;
    double *grad_emb_f64;
    uint32_t *cell_out, *idx_out;
    uint32_t B, D, C, L, H, gridtype, level;
    float S;
    int align_corners, interp, lbc;
} hg_ctx;

/* forward for one level, samples [b0,b1)  (gridencoder.cu:99-244) */
static void fwd_range(uint32_t b0, uint32_t b1, void *vp) {
    const hg_ctx *q = (const hg_ctx *)vp;
    const uint32_t D = q->D, C = q->C, L = q->L, B = q->B, level = q->level, NC = 1u << D;
    const float *grid = q->emb + (size_t)(uint32_t)q->offsets[level] * C;
    const uint32_t hashmap_size = (uint32_t)(q->offsets[level + 1] - q->offsets[level]);
    const float scale = level_scale(level, q->S, q->H, q->level_scales);
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
    for (uint32_t b = b0; b < b1; ++b) {
        const float *x = q->inputs + (size_t)b * D;
        float *out = q->lbc ? q->outputs + ((size_t)level * B + b) * C
                            : q->outputs + (size_t)b * L * C + (size_t)level * C;
        float frac[MAX_D], deriv[MAX_D];
        uint32_t g[MAX_D];
        float *dd = q->dy_dx ? q->dy_dx + ((size_t)b * L + level) * D * C : NULL;
        uint32_t *co = q->cell_out ? q->cell_out + ((size_t)b * L + level) * D : NULL;
        uint32_t *io = q->idx_out ? q->idx_out + ((size_t)b * L + level) * NC : NULL;
        if (!locate(D, x, scale, q->align_corners, q->interp, frac, deriv, g)) {
            for (uint32_t c = 0; c < C; ++c) out[c] = 0.0f;
            if (dd) memset(dd, 0, sizeof(float) * D * C);
            if (co) for (uint32_t d = 0; d < D; ++d) co[d] = 0xFFFFFFFFu;
            if (io) for (uint32_t k = 0; k < NC; ++k) io[k] = 0xFFFFFFFFu;
            continue;
        }
        if (co) for (uint32_t d = 0; d < D; ++d) co[d] = g[d];
        float acc[8] = {0};
        for (uint32_t k = 0; k < NC; ++k) {
            float w = 1.0f;
            uint32_t gl[MAX_D];
            for (uint32_t d = 0; d < D; ++d) {
                if ((k & (1u << d)) == 0) { w *= 1.0f - frac[d]; gl[d] = g[d]; }
                else                      { w *= frac[d];        gl[d] = g[d] + 1; }
            }
            uint32_t index = cell_index(D, C, q->gridtype, q->align_corners, hashmap_size, resolution, gl);
            if (io) io[k] = index / C;
            for (uint32_t c = 0; c < C; ++c) acc[c] = fmaf(w, grid[index + c], acc[c]);
        }
        for (uint32_t c = 0; c < C; ++c) out[c] = acc[c];
        if (!dd) continue;
        for (uint32_t gd = 0; gd < D; ++gd) {
            float ga[8] = {0};
            for (uint32_t k = 0; k < (1u << (D - 1)); ++k) {
                float w = scale;
                uint32_t gl[MAX_D];
                for (uint32_t nd = 0; nd < D - 1; ++nd) {
                    const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                    if ((k & (1u << nd)) == 0) { w *= 1.0f - frac[d]; gl[d] = g[d]; }
                    else                       { w *= frac[d];        gl[d] = g[d] + 1; }
                }
                gl[gd] = g[gd];
                uint32_t il = cell_index(D, C, q->gridtype, q->align_corners, hashmap_size, resolution, gl);
                gl[gd] = g[gd] + 1;
                uint32_t ir = cell_index(D, C, q->gridtype, q->align_corners, hashmap_size, resolution, gl);
                for (uint32_t c = 0; c < C; ++c)
                    ga[c] += w * (grid[ir + c] - grid[il + c]) * deriv[gd];
            }
            for (uint32_t c = 0; c < C; ++c) dd[gd * C + c] = ga[c];
        }
    }
}

/*
 * outputs: [L,B,C] (the reference's layout) when out_lbc != 0,
 *          else [B, L*C] row-major (what grid.py:58 permutes to).
 * dy_dx:   [B, L, D, C] or NULL.
 * cell_out:[B, L, D] uint32 integer cell coordinates (pos_grid) or NULL; 0xFFFFFFFF when out of range.
 * idx_out: [B, L, 2^D] uint32 cell slot within the level (table element index / C) or NULL.
 */
void hg_forward(const float *inputs, const float *emb, const int32_t *offsets, float *outputs,
                uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                float *dy_dx, uint32_t gridtype, int align_corners, int interp,
                const float *level_scales, int out_lbc, uint32_t *cell_out, uint32_t *idx_out) {
    hg_ctx q;
    memset(&q, 0, sizeof(q));
    q.inputs = inputs; q.emb = emb; q.offsets = offsets; q.outputs = outputs; q.dy_dx = dy_dx;
    q.level_scales = level_scales; q.cell_out = cell_out; q.idx_out = idx_out;
    q.B = B; q.D = D; q.C = C; q.L = L; q.S = S; q.H = H; q.gridtype = gridtype;
    q.align_corners = align_corners; q.interp = interp; q.lbc = out_lbc;
    for (uint32_t level = 0; level < L; ++level) { q.level = level; parallel_ranges(B, fwd_range, &q); }
}

/* scatter-add for one level, samples [b0,b1)  (gridencoder.cu:249-340) */
static void bwd_range(uint32_t b0, uint32_t b1, void *vp) {
    const hg_ctx *q = (const hg_ctx *)vp;
    const uint32_t D = q->D, C = q->C, L = q->L, B = q->B, level = q->level, NC = 1u << D;
    float *gg = q->grad_emb + (size_t)(uint32_t)q->offsets[level] * C;
    double *gg64 = q->grad_emb_f64 ? q->grad_emb_f64 + (size_t)(uint32_t)q->offsets[level] * C : NULL;
    const uint32_t hashmap_size = (uint32_t)(q->offsets[level + 1] - q->offsets[level]);
    const float scale = level_scale(level, q->S, q->H, q->level_scales);
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
    for (uint32_t b = b0; b < b1; ++b) {
        const float *x = q->inputs + (size_t)b * D;
        const float *go = q->lbc ? q->grad + ((size_t)level * B + b) * C
                                 : q->grad + (size_t)b * L * C + (size_t)level * C;
        float frac[MAX_D], deriv[MAX_D];
        uint32_t g[MAX_D];
        if (!locate(D, x, scale, q->align_corners, q->interp, frac, deriv, g)) continue;
        for (uint32_t k = 0; k < NC; ++k) {
            float w = 1.0f;
            uint32_t gl[MAX_D];
            for (uint32_t d = 0; d < D; ++d) {
                if ((k & (1u << d)) == 0) { w *= 1.0f - frac[d]; gl[d] = g[d]; }
                else                      { w *= frac[d];        gl[d] = g[d] + 1; }
            }
            uint32_t index = cell_index(D, C, q->gridtype, q->align_corners, hashmap_size, resolution, gl);
            for (uint32_t c = 0; c < C; ++c) {
                atomic_add_f32(&gg[index + c], w * go[c]);
                if (gg64) atomic_add_f64(&gg64[index + c], (double)w * (double)go[c]);
            }
        }
    }
}

/*
 * grad: [L,B,C] when grad_lbc != 0 else [B, L*C]; grad_emb is accumulated in place (caller
 * zeroes it, as grid.py:78 does).  Float accumulation in sample order is one legal ordering
 * of the reference's float atomics; grad_emb_f64 (optional, same shape) receives a double
 * accumulation used as the "true" value in tolerance tests.
 */
void hg_backward(const float *grad, const float *inputs, const int32_t *offsets, float *grad_emb,
                 double *grad_emb_f64, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                 uint32_t H, uint32_t gridtype, int align_corners, int interp,
                 const float *level_scales, int grad_lbc) {
    hg_ctx q;
    memset(&q, 0, sizeof(q));
    q.grad = grad; q.inputs = inputs; q.offsets = offsets; q.grad_emb = grad_emb; q.grad_emb_f64 = grad_emb_f64;
    q.level_scales = level_scales;
    q.B = B; q.D = D; q.C = C; q.L = L; q.S = S; q.H = H; q.gridtype = gridtype;
    q.align_corners = align_corners; q.interp = interp; q.lbc = grad_lbc;
    for (uint32_t level = 0; level < L; ++level) { q.level = level; parallel_ranges(B, bwd_range, &q); }
}

static void inbwd_range(uint32_t b0, uint32_t b1, void *vp) {
    const hg_ctx *q = (const hg_ctx *)vp;
    const uint32_t D = q->D, C = q->C, L = q->L, B = q->B;
    for (uint32_t b = b0; b < b1; ++b)
        for (uint32_t d = 0; d < D; ++d) {
            float r = 0.0f;
            for (uint32_t l = 0; l < L; ++l)
                for (uint32_t c = 0; c < C; ++c) {
                    float go = q->lbc ? q->grad[((size_t)l * B + b) * C + c]
                                      : q->grad[(size_t)b * L * C + (size_t)l * C + c];
                    r += go * q->dy_dx[(((size_t)b * L + l) * D + d) * C + c];
                }
            q->grad_inputs[(size_t)b * D + d] = r;
        }
}

/* grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]   (gridencoder.cu:343-369) */
void hg_input_backward(const float *grad, const float *dy_dx, float *grad_inputs, uint32_t B,
                       uint32_t D, uint32_t C, uint32_t L, int grad_lbc) {
    hg_ctx q;
    memset(&q, 0, sizeof(q));
    q.grad = grad; q.dy_dx = (float *)dy_dx; q.grad_inputs = grad_inputs;
    q.B = B; q.D = D; q.C = C; q.L = L; q.lbc = grad_lbc;
    parallel_ranges(B, inbwd_range, &q);
}
