/*
 * occnerf_b200 -- C ABI of the B200-native OccNeRF per-ray rendering path.
 *
 * Every entry point takes raw DEVICE pointers (unless a parameter is suffixed `_host`), sizes and a
 * CUDA stream (passed as void* so that this header needs no CUDA include), never allocates, and returns
 * 0 on success or a non-zero OCCNERF_E* code; occnerf_last_error() gives the thread-local message.
 * There is no CPU fallback anywhere behind this interface.
 *
 * Reference interfaces replaced (paths relative to the reference tree tiangexiang/OccNeRF):
 *   core/nets/occnerf/gridencoder/src/gridencoder.h:12-15 + bindings.cpp:5-9  -> occnerf_hashgrid_*
 *   core/nets/occnerf/network.py:416-432,456,351-402 (_get_samples_along_ray, _stratified_sampling,
 *       _sample_motion_fields)                                                -> occnerf_warp_*
 *   core/nets/occnerf/knn.py:33-85,102-174 + network.py:236-255 (pykeops K-min) -> occnerf_knn
 *   core/nets/occnerf/canonical_mlps/occnerf_mlp.py:146-167                   -> occnerf_sample_geometry
 *   core/nets/occnerf/canonical_mlps/occnerf_mlp.py:110-126,175-178 (simple_agg) -> occnerf_aggregate_*
 *   core/nets/occnerf/canonical_mlps/occnerf_mlp.py:183-199 (nn.Linear stacks) -> occnerf_sgemm, occnerf_mlp_*
 *   core/nets/occnerf/embedders/hannw_fourier.py:27-45                        -> occnerf_hann_pe
 *   core/nets/occnerf/network.py:320-348,486-499 (_raw2outputs, comp_loss)    -> occnerf_composite_*
 *   core/nets/occnerf/network.py:502-517 (visibility counter)                 -> occnerf_visibility_hits
 *   core/utils/camera_util.py:133-160,163-212 (get_rays_from_KRT, rays_intersect_3d_bbox) + the masking at
 *       core/data/occnerf/freeview.py:208-219, train.py:440-461                 -> occnerf_generate_rays
 *   run.py:39-66 (unpack_alpha_map, unpack_to_image) + core/utils/image_util.py:19-20 (to_8b_image) -> occnerf_unpack_image
 *   core/data/occnerf/train.py:160-165,167-222,225-273 (patch selection of the training loader)    -> occnerf_sample_patches
 */
#ifndef OCCNERF_B200_H
#define OCCNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *occnerf_stream_t; /* cudaStream_t */

enum {
    OCCNERF_OK = 0,
    OCCNERF_EINVAL = 1,   /* bad argument / unsupported configuration */
    OCCNERF_ECUDA = 2     /* a CUDA runtime call or kernel launch failed */
};

/* layout of the [.., L, C] feature axis of the hash grid */
enum {
    OCCNERF_LAYOUT_BLC = 0, /* [B, ld] rows, feature l*C+c at column l*C+c  (what grid.py:58 produces) */
    OCCNERF_LAYOUT_LBC = 1  /* [L, B, C]  (what the reference kernels read/write, gridencoder.cu:99-197) */
};

/* sgemm epilogue flags */
enum {
    OCCNERF_GEMM_BIAS = 1,       /* += bias[j]                                            */
    OCCNERF_GEMM_RELU = 2,       /* max(.,0)                                              */
    OCCNERF_GEMM_ACCUM = 4,      /* += C (read-modify-write; with split_k>1 uses atomics)  */
    OCCNERF_GEMM_RELUMASK = 8    /* *= (mask[i*ldmask+j] > 0)  (backward through a ReLU)   */
};

const char *occnerf_last_error(void);
int occnerf_abi_version(void);

/* ---- K1: ray sampling + 24-bone inverse-LBS warp (network.py:416-432,456,351-402) -------------------
 * rays [N,8] = (o3,d3,near,far); t_lin [S] = linspace(0,1,S) as the host framework computes it;
 * t_rand [N,S] or NULL (stratified jitter, network.py:423-432); Rs [nb,3,3], Ts [nb,3];
 * vol [>=nb, vd,vh,vw] (only the first nb channels are read, network.py:363); bbox_min/scale [3].
 * Outputs: z [N,S], x_skel [N,S,3], mask [N,S]; bins [N,S,nb,3] int32 (floor of the un-normalised voxel
 * coordinate, x,y,z) or NULL. */
int occnerf_warp_forward(const float *rays, const float *t_lin, const float *t_rand, const float *Rs,
                         const float *Ts, const float *vol, const float *bbox_min, const float *bbox_scale,
                         int N, int S, int nb, int vd, int vh, int vw, float *z, float *x_skel, float *mask,
                         int32_t *bins, occnerf_stream_t stream);
/* d(mask) -> g_vol [nb,vd,vh,vw] (accumulated; caller zeroes).  x_skel has no gradient consumer
 * (network.py:225-299 uses it under no_grad only); Rs/Ts gradients: see occnerf_warp_backward_packed. */
int occnerf_warp_backward(const float *rays, const float *t_lin, const float *t_rand, const float *Rs,
                          const float *Ts, const float *bbox_min, const float *bbox_scale, const float *g_mask,
                          int N, int S, int nb, int vd, int vh, int vw, float *g_vol, occnerf_stream_t stream);

/* Corner-packed variant (what the Python operator layer uses).  The 24 bones are sampled at 24 different positions, so
 * what one lookup's 8 corners share is the floor voxel: occnerf_warp_pack_volume re-lays the per-frame volume out as
 * vol8 [nb][vd+1][vh+1][vw+1][8] (cell (z0+1,y0+1,x0+1) = the 8 corners of floor voxel (x0,y0,z0), k = dz*4+dy*2+dx, zero
 * padding baked in; occnerf_warp_packed_floats() floats, 16-byte aligned) and the kernels read one 32-byte sector per
 * (sample, bone).  Block inputs (bone table, rays, jitter tile) are staged by TMA bulk copies.  Same z / x_skel / mask as
 * occnerf_warp_forward, bit for bit. */
long occnerf_warp_packed_floats(int nb, int vd, int vh, int vw);
int occnerf_warp_pack_volume(const float *vol, int nb, int vd, int vh, int vw, float *vol8, occnerf_stream_t stream);
int occnerf_warp_forward_packed(const float *rays, const float *t_lin, const float *t_rand, const float *Rs,
                                const float *Ts, const float *vol8, const float *bbox_min, const float *bbox_scale,
                                int N, int S, int nb, int vd, int vh, int vw, float *z, float *x_skel, float *mask,
                                occnerf_stream_t stream);
/* d(mask) -> g_vol8 (packed layout, accumulated with vector reductions; caller zeroes), folded back into the reference
 * layout g_vol [channels >= nb, vd,vh,vw] (overwritten; channels >= nb get 0) by occnerf_warp_unpack_grad.
 * g_Rs [nb,9] / g_Ts [nb,3] (both or neither; accumulated, caller zeroes): d mask / d motion_scale_Rs, motion_Ts -- what
 * F.grid_sample's grid gradient gives the reference (network.py:367-370) once the pose decoder trains; needs vol8. */
int occnerf_warp_backward_packed(const float *rays, const float *t_lin, const float *t_rand, const float *Rs,
                                 const float *Ts, const float *vol8, const float *bbox_min, const float *bbox_scale,
                                 const float *g_mask, int N, int S, int nb, int vd, int vh, int vw, float *g_vol8,
                                 float *g_Rs, float *g_Ts, occnerf_stream_t stream);
int occnerf_warp_unpack_grad(const float *g_vol8, int nb, int channels, int vd, int vh, int vw, float *g_vol,
                             occnerf_stream_t stream);

/* ---- per-frame prologue of Network.forward (network.py:556-597) ------------------------------------------------------
 * occnerf_pose_refine: BodyPoseRefiner (pose_decoders/mlp_delta_body_pose.py:35-41) + Rodrigues (network_util.py:98-127) +
 *   dst_Rs[1:] . R_delta (network.py:558-570).  w5/b5: HOST arrays of the 5 nn.Linear weight / bias device pointers
 *   (69->256, 256->256 x3, 256->69; [out,in] row-major); posevec69 [69]; dst_Rs [24,3,3] -> Rs_out [24,3,3] (root unchanged).
 * occnerf_motion_basis: MotionBasisComputer (network_util.py:138-200): dst_Rs [24,3,3], dst_Ts [24,3], cnl_gtfms [24,4,4] ->
 *   motion_scale_Rs [24,3,3], motion_Ts [24,3] (forward-kinematics chain along SMPL_PARENT, affine inverse, cnl . inverse).
 * occnerf_weight_volume_forward: vol = softmax_channels(logits + log(priors)) per voxel (deconv_vol_decoder.py:25-33), all
 *   [channels, voxels]; _backward: g_logits = vol * (g_vol - sum_c g_vol * vol).  Forward-only entry points carry no gradient. */
int occnerf_pose_refine(const void *const *w5_host, const void *const *b5_host, const float *posevec69, const float *dst_Rs,
                        int n_bones, float *Rs_out, occnerf_stream_t stream);
int occnerf_motion_basis(const float *dst_Rs, const float *dst_Ts, const float *cnl_gtfms, int n_bones, float *Rs_out,
                         float *Ts_out, occnerf_stream_t stream);
int occnerf_weight_volume_forward(const float *logits, const float *priors, int channels, long voxels, float *vol,
                                  occnerf_stream_t stream);
int occnerf_weight_volume_backward(const float *vol, const float *g_vol, int channels, long voxels, float *g_logits,
                                   occnerf_stream_t stream);

/* ---- loss epilogue: global grad-norm clip + Adam over all trainable tensors (trainer.py:248-249, optimizer.py:12-43) ----
 * n tensors given as HOST arrays of device pointers (params, grads, exp_avg, exp_avg_sq: fp32, numel[t] elements each) with
 * one learning rate per tensor.  clip_grad_norm_ semantics: coef = min(1, max_norm / (||g||_2 + 1e-6)) over ALL tensors
 * (max_norm <= 0: no clipping), applied on the fly; torch.optim.Adam semantics (no weight decay, no amsgrad).
 * steps[t]: one fp32 step counter per tensor in device memory (torch keeps one per parameter), incremented by the call;
 * sumsq: one double in device memory, receives the squared gradient norm.
 * Nothing is read back to the host: CUDA-graph capturable.  Gradients are not modified. */
int occnerf_clip_adam_step(void *const *params, const void *const *grads, void *const *exp_avg, void *const *exp_avg_sq,
                           void *const *steps, const long *numel, const float *lr, int n, float beta1, float beta2, float eps,
                           float max_norm, double *sumsq, occnerf_stream_t stream);

/* ---- exact k-nearest-neighbour search (knn.py:33-85; network.py:236-255,265,508) --------------------
 * queries [m,3]; supports4 [ns,4] = (x,y,z,unused); the support set is split into n_levels contiguous
 * blocks level_begin_host[0..n_levels] (host array); support_gid [ns] maps a support row to the id written
 * to the output (NULL = row index within its level).  out_idx [m, n_levels, k] int32, ascending distance,
 * ties to the lower support row.  Distance = (dx*dx + dy*dy) + dz*dz in fp32, no contraction.
 * query_sel [m] uint8 or NULL: rows with 0 are skipped (their output is left untouched).  k in {3,10}. */
int occnerf_knn(const float *queries, int m, const float *supports4, const int32_t *support_gid,
                const int32_t *level_begin_host, int n_levels, int k, const uint8_t *query_sel,
                int32_t *out_idx, occnerf_stream_t stream);

/* Cluster-pruned exact k-NN over a two-level hierarchy: the `nf` fine points are grouped around the `nc` points of a
 * coarser level (OccNeRF: level 0 around level 2, level 1 around level 3; network.py:113-129).  One call returns the
 * k-NN in BOTH levels, bit-identical to occnerf_knn on the same sets.
 * fine4 [nf,4]: fine points sorted by cluster, .w = bit pattern of the point's original row in its level;
 * centers4 [nc,4]: coarse points in their original order, .w = cluster radius (max member distance, inflated);
 * cluster_ranges [nc,2] int32 = (begin, count) into fine4;  fine_gid [nf] / center_gid [nc]: original row -> output id
 * (NULL = the row itself).  out_fine / out_center: row q at q*out_stride, k entries each.
 * group_stride: queries are laid out [rays, group_stride]; a warp takes 32 consecutive rays at one sample index. */
int occnerf_knn_hier(const float *queries, int m, int group_stride, const float *fine4, const float *centers4,
                     const int32_t *cluster_ranges, int nf, int nc, const int32_t *fine_gid, const int32_t *center_gid,
                     int k, int32_t *out_fine, int32_t *out_center, int out_stride, occnerf_stream_t stream);

/* All four OccNeRF levels in one launch through a three-level cluster tree (level 3 clusters levels 2 and 1, level 2
 * clusters level 0); ids bit-identical to occnerf_knn.  Tables are built once per subject (occnerf_b200.ops.build_knn_tree):
 * p0s [n0,4] level-0 points sorted by level-2 cluster in p2s order (.w = vertex id bits); p1s [n1,4] / p2s [n2,4] sorted by
 * level-3 cluster (.w = row in the level); p3 [n3,4]; c2tab [n2,4] = (radius, begin0, count0, -) per p2s entry;
 * c3tab [n3,4] = (r32, r31, R30, -); c3rng [n3,4] int32 = (begin2, count2, begin1, count1); gid1/2/3: level row -> vertex
 * id; inv2 [n2]: level-2 row -> position in p2s.  out [m,4,k] int32.
 * Queries are laid out [rays, group_stride]; a warp takes lane_rays consecutive rays x (32 / lane_rays) consecutive
 * samples (lane_rays a power of two <= 32) -- a pure scheduling hint, the ids do not depend on it. */
int occnerf_knn_tree(const float *queries, int m, int group_stride, int lane_rays, const float *p0s, const float *p1s, const float *p2s,
                     const float *p3, const float *c2tab, const float *c3tab, const int32_t *c3rng, const int32_t *gid1,
                     const int32_t *gid2, const int32_t *gid3, const int32_t *inv2, int n0, int n1, int n2, int n3, int k,
                     int32_t *out, occnerf_stream_t stream);

/* All four levels through precomputed per-cell candidate lists over a static uniform grid (the support sets never move);
 * ids bit-identical to occnerf_knn.  p0..p3 [n,4]: the level points in their own row order; gid1/2/3: level row ->
 * vertex id; cell_tab [cells,4,2] int32 = (offset into lists, count) per cell and level, cells x-fastest;
 * lists: uint16 level-local rows, every cell's list being a superset of the k nearest of any query inside the cell
 * (occnerf_b200.ops.build_knn_grid states the bound).  grid_min_invh_host = (min x, min y, min z, 1 / cell size) and
 * grid_dims_host = (nx, ny, nz) are HOST arrays.  Queries outside the grid are searched exhaustively (still exact). */
int occnerf_knn_grid(const float *queries, int m, int group_stride, int lane_rays, const float *p0, const float *p1,
                     const float *p2, const float *p3, int n0, int n1, int n2, int n3, const int32_t *gid1,
                     const int32_t *gid2, const int32_t *gid3, const int32_t *cell_tab, const uint16_t *lists,
                     const float *grid_min_invh_host, const int32_t *grid_dims_host, int k, int32_t *out,
                     occnerf_stream_t stream);

/* ---- per-sample surface geometry -> 4-D hash-grid input (occnerf_mlp.py:146-167) --------------------
 * knn_idx rows have `knn_stride` int32 entries, the first 10 being the level-0 neighbours.
 * enc_in [m,4] = (cos-weighted mean of the 3 nearest base vertices normalised to [0,1]^3, clamp((d+.2)/.5));
 * dist[i*dist_stride] signed mean distance (inside vote in fp64); dist_stride = 5 writes raw[:,4] in place. */
int occnerf_sample_geometry(const float *xyz, const int32_t *knn_idx, int knn_stride, const float *point_base,
                            const float *point_norms, float bound, int m, float *enc_in, float *dist, int dist_stride,
                            occnerf_stream_t stream);

/* ---- per-vertex block (network.py:263-284 + occnerf_mlp.py:171-175) ---------------------------------------
 * pc = point_base + point_dist (point_dist [V], one scalar per vertex, added to all three coordinates); kidx3 [V,3] = the 3
 * nearest base vertices of pc (occnerf_knn, k = 3).  Forward writes v_in [V,4] = ((|cos|-weighted mean of the 3 base
 * vertices + bound) / (2 bound), clamp((signed mean distance + 0.2) / 0.8, 0, 1)) -- the vertex's hash-grid input -- and
 * feats_tail[v*ld + 0..3] = (pc, 0) (columns 32..35 of the per-vertex feature table when ld = 36).
 * Backward: g_point_dist [V] from g_v_in [V,4] (the hash grid's input gradient) and g_feats_tail (direct gradient of pc). */
int occnerf_vertex_block_forward(const float *point_base, const float *point_dist, const float *point_norms,
                                 const int32_t *kidx3, float bound, int V, float *v_in, float *feats_tail, int ld,
                                 occnerf_stream_t stream);
int occnerf_vertex_block_backward(const float *point_base, const float *point_dist, const float *point_norms,
                                  const int32_t *kidx3, float bound, int V, const float *g_v_in, const float *g_feats_tail,
                                  int ld, float *g_point_dist, occnerf_stream_t stream);

/* ---- multi-resolution hash grid (gridencoder.h:12-15; gridencoder.cu:50-369) -------------------------
 * level_scales [L] = exp2f(l*S)*H-1 evaluated ON THE DEVICE exactly as gridencoder.cu:138 does. */
int occnerf_hashgrid_level_scales(float S, uint32_t H, uint32_t L, float *level_scales, occnerf_stream_t stream);
/* inputs [B,D] in [0,1]; embeddings [offsets[L], C]; offsets [L+1] int32 (device); outputs per `layout`
 * (ld = row stride in floats for BLC); dy_dx [B, L*D*C] or NULL; cells [B,L,D] / slots [B,L,2^D] uint32 or
 * NULL (integer cell coordinates and table slots, 0xFFFFFFFF where the sample is outside [0,1]^D).
 * run_length 16 (only without dy_dx/cells/slots): inputs are ordered along rays -> one thread walks 16 consecutive
 * samples of a level and re-fetches the 2^D corner vectors only when the cell changes; bitwise the same outputs. */
int occnerf_hashgrid_forward(const float *inputs, const float *embeddings, const int32_t *offsets,
                             const float *level_scales, float *outputs, int layout, int ld, uint32_t B, uint32_t D,
                             uint32_t C, uint32_t L, float *dy_dx, uint32_t *cells, uint32_t *slots, int run_length,
                             occnerf_stream_t stream);
/* grad per `layout`; grad_embeddings accumulated in place (caller zeroes, as grid.py:78 does).
 * run_length: 0 = one thread per (sample, level); 8 or 16 = one thread per (run of that many consecutive samples,
 * level), which merges the reductions of consecutive samples that share a grid cell (inputs ordered along rays).
 * Same sums either way. */
int occnerf_hashgrid_backward(const float *grad, int layout, int ld, const float *inputs, const int32_t *offsets,
                              const float *level_scales, float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C,
                              uint32_t L, int run_length, occnerf_stream_t stream);
/* grad_inputs[b,d] = sum_{l,c} grad[b,l,c] * dy_dx[b,l,d,c]  (gridencoder.cu:343-369) */
int occnerf_hashgrid_input_backward(const float *grad, int layout, int ld, const float *dy_dx, float *grad_inputs,
                                    uint32_t B, uint32_t D, uint32_t C, uint32_t L, occnerf_stream_t stream);

/* ---- visibility-weighted neighbour aggregation (occnerf_mlp.py:110-126,175-178) ----------------------
 * knn_idx [m,nn] int32 vertex ids (nn <= 64); point_counter [V]; feats [V,36] (35 features + 1 pad);
 * writes X[i*ldx + 0..34] = sum_n softmax(att)_n * feats[idx_n], X[i*ldx + 35] = unbiased var(att).
 * att_w [m,nn] or NULL: the forward pass also stores the attention weights there; the backward pass, given them (and
 * nn % 4 == 0), takes the path for samples ordered along rays -- one thread per (run of 16 consecutive samples, 4 neighbour
 * slots, 4 columns) sums its contributions in registers and issues one reduction per run of equal vertices (the vertex in
 * a given slot stays the same from one sample of a ray to the next 69-86 % of the time).  Same sums. */
int occnerf_aggregate_forward(const int32_t *knn_idx, const float *point_counter, const float *feats, int m, int nn,
                              float *X, int ldx, float *att_w, occnerf_stream_t stream);
/* g_feats [copies][V,36] += att_n * gX[i*ldg + 0..34]  (att is detached in the reference, occnerf_mlp.py:123).
 * `copies` >= 1 privatised replicas spread the L2 reductions (CTA b adds into replica b % copies); the caller sums them. */
int occnerf_aggregate_backward(const int32_t *knn_idx, const float *point_counter, const float *gX, int ldg, int m,
                               int nn, float *g_feats, int V, int copies, const float *att_w, occnerf_stream_t stream);


/* ---- Hann-windowed positional encoding (hannw_fourier.py:27-45), window weights from the host ------- */
int occnerf_hann_pe(const float *xyz, int m, const float *window_host, int multires, float *out, int ldo,
                    occnerf_stream_t stream);

/* ---- fp32 SIMT GEMM (the exact-fp32 MLP path): C[i,j] = epi( sum_r A[i*sAi + r*sAr] * B[r*sBr + j*sBj] )
 * split_k > 1 partitions r over blockIdx.z and accumulates with atomics (C must be pre-zeroed or ACCUM). */
int occnerf_sgemm(const float *A, long sAi, long sAr, const float *B, long sBr, long sBj, float *C, long ldc,
                  const float *bias, const float *mask, long ldmask, int Mi, int Nj, int Kr, int flags, int split_k,
                  occnerf_stream_t stream);
/* out[j] (+)= sum_i A[i*lda + j] * (mask ? mask[i*ldmask+j] > 0 : 1)   (bias gradients) */
int occnerf_colsum(const float *A, long lda, const float *mask, long ldmask, int Mi, int Nj, float *out,
                   occnerf_stream_t stream);

/* ---- fused canonical MLP on tcgen05/TMEM (occnerf_mlp.py:183-199) ------------------------------------
 * Layer table (10 layers): pts 68->256,256,256,256 ; geo 256->65 ; rgb 131->256,256,256,256 ; out 256->3.
 * occnerf_mlp_pack_weights re-lays the fp32 nn.Linear weights out as bf16 (hi[,lo]) UMMA operand images.
 * n_pass = 1: bf16 x bf16 -> fp32 ; n_pass = 3: split-bf16 (hi*hi + hi*lo + lo*hi), fp32-grade accuracy ;
 * n_pass = 2: kind::tf32 (fp32 operands rounded to tf32, 10-bit mantissa, two tensor units per product).
 * cta_pair = 1: the chain runs as cta_group::2 CTA pairs (one M = 256 UMMA over two 128-sample tiles, each CTA streams and holds
 * half of the weight rows); the packed image is split by weight-row halves, so pack and run with the same flag. */
typedef struct {
    const float *w[10]; /* pts0..3, geo, rgb0..3, out : [out,in] row-major (nn.Linear.weight) */
    const float *b[10];
} occnerf_mlp_params;
/* chain 0 = forward images (W, plus the padded biases), chain 1 = data-gradient images (W^T). */
long occnerf_mlp_packed_bytes(int n_pass, int chain);
/* Debug only (tools/mlp_stalls.py): with OCCNERF_MLP_DEBUG=1 in the environment the chain kernels accumulate the cycles
 * their role threads spend blocked on each mbarrier.  Synchronises the device, copies the 16 counters to host8 (may be
 * NULL) and clears them when reset != 0.  [0] MMA<-weights [1] MMA<-A operand [2] epilogue<-accumulator
 * [3] producer<-free slot [4] MMA thread total [5] epilogue thread total [6] CTAs [8] epilogue<-TMEM load
 * [9] epilogue in fence+arrive; others unused. */
int occnerf_mlp_debug_counters(unsigned long long *host8, int reset);
/* Debug only: `iters` back-to-back tcgen05.mma (M=128, N=n, K=16, bf16, shared-memory operands) on `ctas` CTAs (one per
 * SM) at once; out_dev[cta] = cycles from the first issue to the completion of the last (tools/mma_rate.py). */
int occnerf_mlp_debug_mma_rate(int iters, int n, int tf32, unsigned long long *out_dev, int ctas, occnerf_stream_t stream);
/* debug only: per-layer clock64 stamps of one tile of CTA 0 (OCCNERF_MLP_DEBUG bit 4): 16 x 12 u64, see csrc/mlp_tc.cu */
int occnerf_mlp_debug_trace(unsigned long long *host192);
/* debug only: weight-stream round-trip stamps of layers 2 and 3 of the same tile: 2 x 8 chunks x 6 events u64 */
int occnerf_mlp_debug_trace_w(unsigned long long *host96);
/* Debug only: cudaOccupancyMaxActiveClusters of the tc3 forward chain kernel for clusters of `cluster_size` CTAs (< 0: error). */
int occnerf_mlp_debug_max_clusters(int cluster_size);
/* debug only: override OCCNERF_MLP_DEBUG for the following launches (debug < 0: back to the environment's value) */
int occnerf_mlp_debug_set(int debug);
int occnerf_mlp_pack_weights(const occnerf_mlp_params *p_host, int n_pass, int chain, int cta_pair, void *packed, occnerf_stream_t stream);
/* XB [m,132]: columns 64..131 = (agg35, var1, h32) are read; columns 0..63 receive the 64 geometry features when
 * act_dtype != 0.  raw [m, ldr]: columns 0..3 = (rgb_pre3, sigma_pre1) are written.
 * act_save: NULL (act_dtype 0, inference) or a buffer receiving the post-ReLU activations of the 8 hidden layers
 * (slots 0..3 = pts1..4, 4..7 = rgb1..4) for the backward pass:
 * bf16 CHUNK-MAJOR [10][32][slot_stride][8] (act_dtype 2; element (slot, row, col) at ((slot*32 + col/8)*slot_stride +
 * row)*8 + col%8; slot 8 = input of pts0 (80 columns), slot 9 = input of rgb0 (144 columns)) -- the layout in which a
 * warp's stores are contiguous and which the weight-gradient kernel's TMA consumes directly.
 * relu_mask (required with act_dtype 2): [8][32][slot_stride] bytes; byte (slot, k8, row) holds [hidden unit 8*k8+c > 0] of columns
 * c = 0..7 at bit (c >> 1) + 4 * (c & 1) (the order in which the packed bf16 pairs are compared). */
int occnerf_mlp_forward_tc(float *XB, int m, const void *packed, int n_pass, int cta_pair, float *raw, int ldr, void *act_save,
                           int act_dtype, long slot_stride, void *relu_mask, occnerf_stream_t stream);

/* Fused data-gradient chain (the transposed layers in reverse order, ReLU masks from the saved bf16 activations).
 * g_raw [m,5] (d rgb_pre3, d sigma_pre, unused); relu_mask [8][32][slot_stride] as saved by occnerf_mlp_forward_tc.
 * Writes gXB [m,132] columns 64..131 = d(agg35, var1, h32) summed over both trunks, and g_save, bf16 chunk-major
 * [10][32][slot_stride][8] like act_save = gradients w.r.t. the pre-activations: slots 0..3 = rgb3, rgb2, rgb1, rgb0;
 * 4 = geo (columns 0..63 features, 64 sigma); 5..8 = pts3, pts2, pts1, pts0; 9 = d raw[:, :3]. */
int occnerf_mlp_backward_tc(const float *g_raw, int m, const void *packed_bwd, int n_pass, int cta_pair, const void *relu_mask, float *gXB,
                            void *g_save, long slot_stride, occnerf_stream_t stream);

/* Weight and bias gradients dW_l = G_l^T X_l, db_l = colsum(G_l) of all 10 layers from the two bf16 buffers above
 * (chunk-major layout; TMA + MN-major tcgen05, contraction over the sample axis).  slot_stride must be a multiple of 64
 * and the rows m..slot_stride-1 of every slot zero.  dW [10][256][256], dB [10][256] fp32 are ACCUMULATED (caller zeroes), in the
 * padded/permuted layer layout of the fused kernels (see occnerf_b200/mlp_tc.py for the mapping back to nn.Linear). */
int occnerf_mlp_wgrad_tc(const void *g_save, const void *act_bf16, int m, long slot_stride, float *dW, float *dB,
                         occnerf_stream_t stream);

/* ---- non-rigid motion MLP on tcgen05 (non_rigid_motion_mlps/mlp_offset.py:7-62; forward only, as in the reference its
 * output feeds no_grad code only).  out [m,3] = xyz + MLP([cond69, pe36]) with the 6 x 128 ReLU layers and the skip
 * connection of the reference; the Hann-windowed positional encoding (6 bands, hannw_fourier.py:27-45, window weights
 * window6_host = 6 HOST floats) is evaluated inside the kernel.  w7_host / b7_host: HOST arrays of 7 device pointers
 * (nn.Linear weights [128,105], [128,128] x3, [128,164], [128,128], [3,128] and their biases); cond_dev: the 69-value pose
 * condition on the device (folded into the first bias) or NULL (= zeros).  Buffer size: occnerf_mlp_packed_bytes(n_pass, 2). */
/* cta_pair: as for the canonical chain (0 = cta_group::1 CTAs sharing the weight stream, 1 = cta_group::2 pairs); the same value for
 * the packing and the forward call. */
int occnerf_nonrigid_pack_weights(const void *const *w7_host, const void *const *b7_host, const float *cond_dev, int n_pass,
                                  int cta_pair, void *packed, occnerf_stream_t stream);
int occnerf_nonrigid_forward_tc(const float *xyz, const float *window6_host, int m, const void *packed, int n_pass, int cta_pair,
                                float *out, occnerf_stream_t stream);

/* ---- alpha compositing (network.py:320-348) + completeness term (network.py:486-499) ----------------
 * raw [N,S,5] = (rgb_pre3, sigma_pre, dist); mask, z [N,S]; rays [N,8]; bg [3] (0..255).
 * Outputs rgb [N,3], acc [N], depth [N], term [N] int64 (argmax alpha, first maximum),
 * weights [N,S] or NULL, comp [N,S] or NULL. */
int occnerf_composite_forward(const float *raw, const float *mask, const float *z, const float *rays, const float *bg,
                              int N, int S, float *rgb, float *acc, float *depth, int64_t *term, float *weights,
                              float *comp, occnerf_stream_t stream);
/* Transmittance is recomputed, nothing but the forward inputs is read.  g_comp may be NULL.
 * Writes g_raw [N,S,5] (channel 4 = 0) and g_mask [N,S]. */
int occnerf_composite_backward(const float *raw, const float *mask, const float *z, const float *rays, const float *bg,
                               const float *g_rgb, const float *g_acc, const float *g_depth, const float *g_comp, int N,
                               int S, float *g_raw, float *g_mask, occnerf_stream_t stream);

/* ---- visibility counter update (network.py:502-517) --------------------------------------------------
 * For rays with depth > thresh (only if more than one such ray) the canonical sample x_skel[ray, term[ray]]
 * votes for its k nearest points of cloud4 [V,4]; hits [V] receives 1.0 for every voted point (else 0).
 * scratch: >= (N*(k+4) + 4) * 4 bytes. */
int occnerf_visibility_hits(const float *depth, const int64_t *term, const float *x_skel, int N, int S, float thresh,
                            const float *cloud4, int V, int k, float *hits, void *scratch, occnerf_stream_t stream);

/* ---- rays in front of the path: camera_util.py:133-160 + :163-212 + freeview.py:208-219 ----------------------
 * All camera / box arguments are HOST pointers to float64 (they ride in the kernel arguments; there is no H2D copy):
 * kinv_host [9] = inv(K) row major as numpy computed it in K's own dtype; k_is_f32 = OCCNERF_RAYS_K_F32 when K was
 * float32 (the ZJU pickles), in which case `pixel_camera` is evaluated in float32 like numpy does, OCCNERF_RAYS_ALL_F32
 * when K, R and T all were float32 (tpose.py:66-84: origin, directions and |d| stay float32), else OCCNERF_RAYS_F64;
 * R_host [9], T_host [3] extrinsics;
 * bbox_min/max_host [3] (the 1 cm margin of camera_util.py:180 is added inside).
 * Outputs (device): mask [H*W] uint8 = `ray_mask`; count [1] = number of valid rays; rays [capacity, 8] =
 * (o3, d3, near, far) float32 of the valid rays in pixel order (what `rays_o[ray_mask]` etc. produce; d carries the
 * in-place 1e-5 clamp of camera_util.py:183); pixel_index [capacity] int32 = flat pixel of each ray, or NULL.
 * Rays beyond `capacity` are dropped (compare count with capacity); capacity = 0 only fills mask and count.
 * scratch: occnerf_rays_scratch_bytes(H, W) bytes. */
enum { OCCNERF_RAYS_F64 = 0, OCCNERF_RAYS_K_F32 = 1, OCCNERF_RAYS_ALL_F32 = 2 };
long occnerf_rays_scratch_bytes(int H, int W);
int occnerf_generate_rays(const double *kinv_host, int k_is_f32, const double *R_host, const double *T_host,
                          const double *bbox_min_host, const double *bbox_max_host, int H, int W, int capacity,
                          float *rays, uint8_t *mask, int *pixel_index, int *count, void *scratch,
                          occnerf_stream_t stream);

/* ---- image assembly behind the path: run.py:39-66 + image_util.py:19-20 ---------------------------------------
 * rgb [n,3], alpha [n] or NULL: outputs of the n rays whose flat pixels are pixel_index [n] (int32, from
 * occnerf_generate_rays; a rank of a sharded render passes its own range).  bgcolor_host [3]: HOST floats in [0,1]
 * (run.py:118 passes cfg.bgcolor / 255).  fill != 0 first paints the whole frame with the background (alpha 0), as
 * np.full / np.zeros do in the reference; fill = 0 only scatters (further shards into the same frame).
 * Outputs: rgb8 [H*W,3] uint8 = to_8b_image of the assembled frame, alpha8 [H*W] uint8 or NULL (one channel; the
 * reference stacks it three times for display).  bad [1] int32 (caller-zeroed) counts pixel indices outside the frame. */
int occnerf_unpack_image(const float *rgb, const float *alpha, const int *pixel_index, int n, int H, int W,
                         const float *bgcolor_host, int fill, uint8_t *rgb8, uint8_t *alpha8, int *bad,
                         occnerf_stream_t stream);

/* ---- training-patch selection (core/data/occnerf/train.py:167-222 get_patch_ray_indices, :225-273 _get_patch_ray_indices, and the
 * gathers of :160-165 sample_patch_rays), csrc/patches.cu.  ray_mask / subject_mask / bbox_mask: [H*W] bytes (0 / non-zero).
 * use_subject [n_patch] bytes and select_idx [n_patch] i32 (device): the caller's two random draws per patch -- the reference draws
 * np.random.rand(1)[0] < cfg.patch.sample_subject_ratio and np.random.choice(n_candidates, size=[1], replace=False)[0], in that order.
 * Outputs: select_inds [n_patch * patch^2] i32 (ranks in the compacted ray list; the first patch_div[n_patch] entries are valid),
 * patch_div [n_patch + 1] i32, patch_masks [n_patch, patch, patch] bytes, xy_min / xy_max [n_patch, 2] i32 as (x, y), status [1] i32
 * (caller-zeroed; 1 = a select_idx was outside its candidate list).  rays [n_rays, 8] -> rays_out [n_patch * patch^2, 8]: optional
 * gather of the selected rays (both NULL to skip).  scratch: occnerf_patches_scratch_bytes(H, W, n_patch, patch) bytes. */
long occnerf_patches_scratch_bytes(int H, int W, int n_patch, int patch);
int occnerf_sample_patches(const uint8_t *ray_mask, const uint8_t *subject_mask, const uint8_t *bbox_mask, int H, int W, int patch,
                           int n_patch, const uint8_t *use_subject, const int32_t *select_idx, const float *rays, float *rays_out,
                           int32_t *select_inds, int32_t *patch_div, uint8_t *patch_masks, int32_t *xy_min, int32_t *xy_max,
                           int32_t *status, void *scratch, occnerf_stream_t stream);

/* ---- loss epilogue without the LPIPS term (core/train/trainers/occnerf/trainer.py:31-41 _unpack_imgs, :24 img2mse, :135-189 get_loss),
 * csrc/loss.cu: value and gradients in one call.  rgb [n, 3] with n = div[N]; masks [N, P, P] bytes; div [N + 1] i32; bgcolor [3] in
 * [0, 1]; targets [N, P, P, 3]; comp [comp_numel] or NULL.  Outputs: imgs [N, P, P, 3] or NULL (the unpacked patch images, what the
 * perceptual loss consumes), out [4] = (w_mse * mse + w_comp * mean(comp), w_mse * mse, w_comp * mean(comp), d loss / d comp_i),
 * g_rgb [n, 3] = d loss / d rgb.  acc: 16 bytes of scratch. */
int occnerf_patch_loss(const float *rgb, const uint8_t *masks, const int32_t *div, const float *bgcolor, const float *targets,
                       const float *comp, long comp_numel, int N, int P, float w_mse, float w_comp, float *imgs, float *out, float *g_rgb,
                       void *acc, occnerf_stream_t stream);

/* ---- per-frame prologue: the motion-weight volume decoder (deconv_vol_decoder.py:25-33, network_util.py:12-50), csrc/deconv.cu ----
 * ConvTranspose3d(kernel 4, stride 2, padding 1), batch 1, as tf32 tensor-core GEMMs on the reference's weight layout
 * W [Cin][Cout][4][4][4].  Activations are PRE-activations [C][D^3] (NCDHW, N = 1); the LeakyReLU(slope) between layers is applied
 * when a layer loads its input (slope = 1: none).  D (input edge) must be a power of two <= 64.
 * forward : Yout [Cout][(2D)^3] = bias + deconv(act(Yin)).
 * backward: from dYout: dW (stored; reduced into a caller-zeroed buffer when accumulate_dw or w_splits > 1), dbias [Cout], and --
 *           unless dYin is NULL -- dYin [Cin][D^3] = act'(Yin) * (data gradient) (caller-zeroed when d_splits > 1).
 * splits: how many CTAs share one contraction (split-K with fp32 reductions); 1 = none.
 * exact: 0 = one tf32 MMA per product (what the library path computes under torch.backends.cudnn.allow_tf32, the default);
 *        1 = 3 x tf32 on (hi, lo) operand splits, fp32-grade (allow_tf32 = False). */
int occnerf_deconv3d_forward(const float *W, const float *bias, const float *Yin, int Cin, int Cout, int D, float slope, int splits,
                             int exact, float *Yout, occnerf_stream_t stream);
int occnerf_deconv3d_backward(const float *W, const float *Yin, const float *dYout, int Cin, int Cout, int D, float slope, int w_splits,
                              int d_splits, int accumulate_dw, int exact, float *dW, float *dbias, float *dYin, occnerf_stream_t stream);
/* backward runs bias + weight gradient on an internal side stream, forked from and joined to `stream` inside the call (event dependencies,
 * capturable), beside the data gradient; 0 switches that off (everything in order on `stream`).  Process-wide, default 1. */
int occnerf_deconv_set_overlap(int on);
/* the 256 -> 1024 linear layer in front of the stack (pre-activation out; network_util.py:25-28), batch 1, and its gradients */
int occnerf_decoder_linear_forward(const float *w, const float *b, const float *e, int n_out, int n_in, float *y, occnerf_stream_t stream);
int occnerf_decoder_linear_backward(const float *w, const float *e, const float *g, int n_out, int n_in, float *dw, float *db, float *de,
                                    occnerf_stream_t stream);

/* ---- data-parallel training: gradient all-reduce as one kernel over NVSwitch peer memory (csrc/collective.cu) ----------------
 * Replaces the gradient reduction that nn.DataParallel / a NCCL all-reduce performs between backward() and the optimizer step
 * (core/train/trainers/occnerf/trainer.py:246-248).  In-place SUM over `world` ranks of a flat fp32 buffer that every rank holds in
 * symmetric (peer-mapped) memory.  peer_bufs_host / peer_pads_host: HOST arrays of `world` device pointers to every rank's buffer /
 * signal pad (entry `rank` is the local one); multicast: the NVLS multicast mapping of the buffer (multimem.ld_reduce / multimem.st
 * through the switch) or NULL (peer loads and stores); n_floats % 4 == 0; pad: >= blocks * world u32 per rank and epochs: [blocks] u32
 * of local device memory, both zeroed once before the first call; blocks: 1..148, identical on all ranks.  Ranks synchronise inside
 * the kernel (epoch counters in device memory), so the launch is CUDA-graph replayable. */
/* debug only: nanoseconds block 0 spent (0) waiting for the peers to arrive, (1) in the data phase, (2) waiting for them to finish,
 * and (3) the number of launches, summed since the last reset */
int occnerf_allreduce_debug(unsigned long long *host4, int reset);
int occnerf_allreduce_sum_f32(const void *const *peer_bufs_host, const void *const *peer_pads_host, void *multicast, long n_floats,
                              int rank, int world, int blocks, void *epochs, occnerf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OCCNERF_B200_H */
