/* sc_b200.h — C ABI of libsc_b200.so: the B200 (sm_100a) implementation of ShapeClipper's hot path.
 *
 * Every entry point takes raw DEVICE pointers, explicit sizes and an explicit cudaStream_t, launches
 * asynchronously on that stream, keeps no reference to caller memory, holds no mutable global state
 * (re-entrant: autograd calls backward from another host thread) and returns 0 on success or the
 * cudaError_t of the failed call. Outputs are allocated by the caller, as in the reference's binding.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   sc_chamfer_forward   external/chamfer3D/chamfer_cuda.cpp:17-19  chamfer_forward  -> chamfer3D.cu:137-154
 *   sc_chamfer_backward  external/chamfer3D/chamfer_cuda.cpp:22-26  chamfer_backward -> chamfer3D.cu:176-195
 *   sc_clip_encode       CLIP_anno.py:166 clip_model.encode_image (openai/CLIP VisionTransformer)
 *   sc_cosine_topk       CLIP_anno.py:29-57 NN_annotator.calc_matches
 *   sc_render_*          model/renderer.py:57 Renderer.forward, model/implicit.py:163 get_conditional_output
 *   sc_mc_* / sc_tri_*   utils/eval_3D.py:123-153 (mcubes.marching_cubes, trimesh sample)
 *   sc_boundary_distance utils/util.py:237-248 compute_sampling_prob (vigra.filters.boundaryDistanceTransform)
 */
#ifndef SC_B200_H_
#define SC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

/* ABI version of this header; bumped whenever a signature changes. */
#define SC_B200_ABI_VERSION 4
int sc_abi_version(void);

/* ---- chamfer3D ------------------------------------------------------------
 * xyz1 [batch,n,3], xyz2 [batch,m,3] contiguous fp32.
 * forward : dist1 [batch,n], dist2 [batch,m] = SQUARED distance to the nearest point of the other cloud,
 *           idx1/idx2 int32 = its index (lowest index among exact ties). Bit-exact with the reference.
 *           workspace: device scratch of sc_chamfer_workspace_bytes(batch,n,m) bytes.
 * backward: gradxyz1/gradxyz2 must be zeroed by the caller (reference contract); float atomics. */
size_t sc_chamfer_workspace_bytes(int batch, int n, int m);
int sc_chamfer_forward(const float* xyz1, const float* xyz2, int batch, int n, int m,
                       float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream);
int sc_chamfer_backward(const float* xyz1, const float* xyz2, int batch, int n, int m,
                        const float* graddist1, const float* graddist2,
                        const int32_t* idx1, const int32_t* idx2,
                        float* gradxyz1, float* gradxyz2, cudaStream_t stream);

/* ---- ray geometry of the selected pixels --------------------------------------------
 * Replaces utils/camera.py:157-196 (get_camera_grid, get_center_and_ray; perspective camera) + model/renderer.py:59-68.
 * pose [B,3,4] world->camera, intr [B,3,3], ray_idx [B,R] int64 flat row-major pixel ids (NULL = pixels 0..R-1).
 * -> cam_loc [B,3], unit ray_dirs [B,R,3], depth_fac [B,R]. backward: adjoints of those three (NULL = zero) ->
 * pose_bar [B,3,4], intr_bar [B,3,3] (either may be NULL); workspace: batch * 18 floats. */
int sc_pixel_rays_forward(const float* pose, const float* intr, const int64_t* ray_idx, int batch, int n_rays,
                          int width, float* cam_loc, float* ray_dirs, float* depth_fac, cudaStream_t stream);
int sc_pixel_rays_backward(const float* pose, const float* intr, const int64_t* ray_idx, int batch, int n_rays,
                           int width, const float* cam_loc_bar, const float* ray_dirs_bar, const float* depth_fac_bar,
                           float* workspace, float* pose_bar, float* intr_bar, cudaStream_t stream);

/* Eikonal sample points (model/renderer.py:154-170): points [B,2R,3] = cat(uniform_pts [B,R,3], cam_loc + z_eik * ray_dirs),
 * z_eik [B,R] = depth of sample eik_idx [B*R] (int64) of each ray, stratified with jitter [B*R,S] when jitter != NULL
 * (model/renderer.py:13-37). backward: points_bar -> ray_dirs_bar [B,R,3] and acc [B][4] = (cam_loc_bar xyz, scale_dist_bar). */
int sc_eikonal_points_forward(const float* cam_loc, const float* ray_dirs, const float* scale_dist, const float* t_vals,
                              const float* jitter, const int64_t* eik_idx, const float* uniform_pts, int batch, int n_rays,
                              int n_samples, float cam_dist, float half_range, float* points, float* z_eik,
                              cudaStream_t stream);
int sc_eikonal_points_backward(const float* ray_dirs, const float* z_eik, const float* points_bar, int batch, int n_rays,
                               float cam_dist, float* ray_dirs_bar, float* acc, cudaStream_t stream);

/* ---- render-consuming losses of one render: values and unit gradients -------------
 * Replaces model/loss.py:19-97 as combined in model/graph.py:220-265: MSE(rgb), soft-IoU mask loss, trimmed normal loss,
 * eikonal MSE. pass1 -> (caller: order = stable argsort(key)) -> pass2. rgb / normal [B,R,3], mask [B,R], eik [n_eik]
 * (normal / eik may be NULL). workspace: sc_render_losses_workspace_floats(batch) floats. Outputs: losses[4] = render,
 * mask, normal, eikonal and the gradient of each loss w.r.t. its inputs for unit upstream weight. */
size_t sc_render_losses_workspace_floats(int batch);
int sc_render_losses_pass1(const float* rgb, const float* rgb_t, const float* mask, const float* mask_t,
                           const float* normal, const float* normal_t, const float* eik, int n_eik, int batch,
                           int n_rays, float normal_l1, float* workspace, float* key, float* per_px,
                           float* rgb_unit, float* normal_unit, float* normal_t_unit, float* eik_unit,
                           cudaStream_t stream);
int sc_render_losses_pass2(const float* mask, const float* mask_t, const float* normal, const float* normal_t,
                           const int64_t* order, const float* per_px, int n_eik, int batch, int n_rays, float normal_l1,
                           double normal_tol, float mask_mse, float* workspace, float* mask_unit, float* normal_unit,
                           float* normal_t_unit, float* losses, cudaStream_t stream);

/* ---- fused SDF/RGB-MLP volume renderer ----------------------------------
 * Replaces model/renderer.py:57-209 (Renderer.forward), model/implicit.py:138-239 (SDFNetwork.forward,
 * get_conditional_output, RGBNetwork.forward, LaplaceDensity) and the slice loop of utils/eval_3D.py:21-38.
 *
 * Weights travel as one packed blob (sc_render_pack_weights) built from the nn.Linear tensors of
 * sdf_network.lin0..5 / rgb_network.lin0..3 ([out,in] row-major fp32, the checkpoint layout).
 * Per-image latent biases cb [B][4][64] come from sc_render_latent_bias (z_sdf, z_rgb are [B,64]).
 * scratch: sc_render_scratch_bytes(backward) bytes of device memory private to one call in flight. */
typedef struct ScRenderArgs {
    int mode;               /* 0 = rays (Renderer.forward), 1 = points (SDFNetwork.get_conditional_output) */
    int batch;              /* B */
    int n_per_image;        /* rays per image R (mode 0) / points per image N (mode 1) */
    int n_samples;          /* S samples per ray (mode 0): 4 | S, S | 128 */
    int want_grad;          /* mode 1: also produce d sdf / d point */
    int want_feat;          /* mode 1: also produce the 64 SDF features */
    int detach_latent;      /* backward, mode 1: latent bias adjoints do not reach z_sdf (compute_grad=True path) */
    float beta_min;         /* LaplaceDensity beta_min (1e-4) */
    float cam_dist;         /* opt.camera.dist */
    float half_range;       /* 0.7 */
    float bg_color;         /* opt.data.bgcolor */
    float normal_pow;       /* opt.reg.normal_pow */
    const float* blob;      /* packed weights, sc_render_blob_floats() floats */
    const float* cb;        /* [B][4][64] */
    const float* beta_param;/* device scalar: renderer.density.beta */
    /* mode 0 inputs */
    const float* cam_loc;   /* [B,3] */
    const float* ray_dirs;  /* [B,R,3] unit */
    const float* depth_fac; /* [B,R] */
    const float* scale_dist;/* [B] */
    const float* t_vals;    /* [S] linspace(0,1,S) */
    const float* jitter;    /* [B*R,S] stratified u in [0,1) or NULL (bin edges) */
    /* mode 1 input */
    const float* points;    /* [B,N,3] */
    /* forward outputs (mode 0) */
    float* rgb;             /* [B,R,3] */
    float* mask;            /* [B,R] */
    float* mask_hard;       /* [B,R] */
    float* depth;           /* [B,R] */
    float* normal;          /* [B,R,3] */
    /* forward outputs (mode 1) */
    float* sdf;             /* [B,N] */
    float* feat;            /* [B,N,64] or NULL */
    float* grad;            /* [B,N,3] or NULL */
    /* backward inputs: adjoints of the outputs above (NULL = zero) */
    const float* rgb_bar; const float* mask_bar; const float* depth_bar; const float* normal_bar;
    const float* sdf_bar; const float* grad_bar;
    /* backward outputs */
    float* grad_partial;    /* [n_ctas][sc_render_grad_floats()] per-CTA folded weight-gradient partials (zeroed by callee) */
    float* cb_bar;          /* [B][7][64] per-image bias adjoints (zeroed by caller) */
    float* ray_dirs_bar;    /* [B,R,3] */
    float* depth_fac_bar;   /* [B,R] */
    float* cam_loc_bar;     /* [B,3]  (zeroed by caller; atomics) */
    float* scale_dist_bar;  /* [B]    (zeroed by caller; atomics) */
    float* points_bar;      /* [B,N,3] (mode 1) */
    void* scratch;          /* sc_render_scratch_bytes() */
    /* mode 0, tensor-core kernels (sc_render_tc_forward / sc_render_tc_backward), optional: a buffer of
     * sc_render_tc_saved_bytes(batch, n_per_image, n_samples) bytes. Forward: the per-point activations the backward
     * needs (H, Q, FEAT, R, GPE planes + per-point vectors, 3.75 KB per sample point) are written to it. Backward: they are
     * read back instead of recomputing the forward per tile (19 of its 45 GEMM phases). NULL = recompute, no memory. */
    void* saved;
    /* tensor-core kernels only. 0: every product is three MMAs on hi/lo bf16 operand pairs (fp32-class, the 1e-4 parity mode).
     * 1: ONE MMA per product on the hi planes alone — plain bf16 operands with fp32 accumulation, the arithmetic BASELINE.json
     * configs[2] names (about 1e-2 relative on rendered outputs); element-wise work (posenc, activations, compositing) stays fp32. */
    int precision;
} ScRenderArgs;

size_t sc_render_blob_floats(void);
size_t sc_render_grad_floats(void);
int sc_render_num_ctas(void);                      /* persistent grid = SM count of the current device */
size_t sc_render_scratch_bytes(int backward);
/* w/b: arrays of 10 device pointers: sdf lin0..5 then rgb lin0..3 (weights [out,in], biases [out]) */
int sc_render_pack_weights(const float* const* w, const float* const* b, float* blob, cudaStream_t stream);
int sc_render_latent_bias(const float* blob, const float* z_sdf, const float* z_rgb, int batch, float* cb,
                          cudaStream_t stream);
int sc_render_forward(const ScRenderArgs* args, cudaStream_t stream);
/* backward: same args + the *_bar adjoints; recomputes the forward per tile (no saved activations). Writes
 * grad_partial / cb_bar / geometry adjoints; sc_render_grad_finalize turns those into nn.Linear-layout gradients:
 * out_w/out_b = arrays of 10 device pointers (NULL entries are skipped), z_sdf_bar/z_rgb_bar [B,64],
 * beta_bar = d/d(|beta|+beta_min) (1 float). */
int sc_render_backward(const ScRenderArgs* args, cudaStream_t stream);
/* Tensor-core edition of the same kernels (tcgen05.mma on hi/lo bf16 operand pairs, accumulators in TMEM; same
 * ScRenderArgs, same outputs to ~1e-5). args->blob must then be the blob of sc_render_tc_pack_weights and args->scratch
 * sc_render_tc_scratch_bytes() bytes; cb / grad_partial / cb_bar and sc_render_grad_finalize are shared. */
size_t sc_render_tc_blob_bytes(void);
size_t sc_render_tc_scratch_bytes(int backward);
int sc_render_tc_pack_weights(const float* const* w, const float* const* b, const float* ffma_blob, void* tc_blob,
                              cudaStream_t stream);
size_t sc_render_tc_saved_bytes(int batch, int n_per_image, int n_samples);
int sc_render_tc_forward(const ScRenderArgs* args, cudaStream_t stream);
int sc_render_tc_backward(const ScRenderArgs* args, cudaStream_t stream);
/* Second generation of the tensor-core kernels: two independent 64-point tile chains per CTA (csrc/render_tc2.cuh). Same
 * ScRenderArgs, blob (sc_render_tc_pack_weights), scratch (sc_render_tc_scratch_bytes), partials and finalize as above.
 * Rays must fit one 64-point tile: sc_render_tc2_supported(mode, n_samples) says whether (4 <= S <= 64, S | 64). */
int sc_render_tc2_supported(int mode, int n_samples);
int sc_render_tc2_forward(const ScRenderArgs* args, cudaStream_t stream);
int sc_render_tc2_backward(const ScRenderArgs* args, cudaStream_t stream);
int sc_render_grad_finalize(const float* grad_partial, int n_ctas, const float* cb_bar, const float* z_sdf,
                            const float* z_rgb, const float* blob, int batch, float* const* out_w,
                            float* const* out_b, float* z_sdf_bar, float* z_rgb_bar, float* beta_bar,
                            cudaStream_t stream);

/* Same, but out_w / out_b are ADDED to (fused gradient accumulation into existing .grad tensors: one launch instead of one
 * accumulation kernel per parameter); z_sdf_bar / z_rgb_bar / beta_bar are still overwritten. */
int sc_render_grad_finalize_accumulate(const float* grad_partial, int n_ctas, const float* cb_bar, const float* z_sdf,
                                       const float* z_rgb, const float* blob, int batch, float* const* out_w,
                                       float* const* out_b, float* z_sdf_bar, float* z_rgb_bar, float* beta_bar,
                                       cudaStream_t stream);

/* ---- CLIP ViT image tower + cosine k-NN ------------------------------------------
 * Replaces clip_model.encode_image + F.normalize (CLIP_anno.py:166-167; openai/CLIP, un-vendored dependency) and
 * NN_annotator.calc_matches (CLIP_anno.py:29-57). GEMMs run on tcgen05 tensor cores from TMA-staged bf16 tiles with
 * fp32 TMEM accumulation; with `split` set every operand is a hi/lo bf16 pair and each product is 3 MMAs
 * (fp32-class accuracy); otherwise plain bf16 (throughput mode).
 * Weight matrices are [out,in] row-major bf16 planes (lo planes may be NULL when split == 0); vectors are fp32. */
typedef struct ScClipConfig {
    int image_size;   /* 224 */
    int patch;        /* 32 (ViT-B/32) or 14 (ViT-L/14) */
    int width;        /* 768 / 1024 ; heads * 64 */
    int layers;       /* 12 / 24 */
    int heads;        /* 12 / 16 */
    int out_dim;      /* 512 / 768 */
    int split;        /* 1 = hi/lo operands (3 MMAs per product), 0 = plain bf16 */
} ScClipConfig;
typedef struct ScClipLayer {
    const float *ln1_w, *ln1_b;
    const void *qkv_w_hi, *qkv_w_lo; const float* qkv_b;     /* in_proj  [3W, W] */
    const void *out_w_hi, *out_w_lo; const float* out_b;     /* out_proj [W, W]  */
    const float *ln2_w, *ln2_b;
    const void *fc1_w_hi, *fc1_w_lo; const float* fc1_b;     /* c_fc     [4W, W] */
    const void *fc2_w_hi, *fc2_w_lo; const float* fc2_b;     /* c_proj   [W, 4W] */
} ScClipLayer;
typedef struct ScClipWeights {
    const void *conv_w_hi, *conv_w_lo;                       /* conv1.weight viewed as [W, 3*P*P], columns zero-padded to a multiple of 64 */
    const float *class_emb, *pos_emb;                        /* [W], [T, W] */
    const float *lnpre_w, *lnpre_b, *lnpost_w, *lnpost_b;
    const void *proj_w_hi, *proj_w_lo;                       /* proj^T: [out_dim, W] */
    const ScClipLayer* layers;                               /* HOST array of cfg.layers entries (device pointers inside) */
} ScClipWeights;

/* C[M,N] = A[M,K] . W[N,K]^T * scale (+bias[N]) (act 1 = QuickGELU) (+residual fp32 [M,N]) -> out_f32 and/or hi/lo planes.
 * a_lo / w_lo NULL => single bf16 product. K % 64 == 0, N % 64 == 0. */
int sc_gemm_bf16_tc(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, int M, int N, int K,
                    const float* bias, const float* residual, int act, float scale, float* out_f32, void* out_hi,
                    void* out_lo, cudaStream_t stream);
size_t sc_clip_workspace_bytes(const ScClipConfig* cfg, int batch);
/* images [B,3,S,S] fp32 (already CLIP-normalised) -> emb [B,out_dim] L2-normalised fp32 (+ optional unnormalised copy and
 * hi/lo bf16 planes of the normalised embedding for sc_cosine_topk). */
int sc_clip_encode(const ScClipConfig* cfg, const ScClipWeights* weights, const float* images, int batch,
                   float* emb, float* emb_unnormalised, void* emb_hi, void* emb_lo, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream);
/* values/indices [n_query,k]: the k largest cosine similarities of each query against the first n_bank_valid rows of
 * the bank (ties: lowest index). sim_workspace: fp32 [n_query, n_bank]. n_bank (allocated rows) % 64 == 0, dim % 64 == 0. */
int sc_cosine_topk(const void* q_hi, const void* q_lo, const void* bank_hi, const void* bank_lo, int n_query,
                   int n_bank, int n_bank_valid, int dim, int k, float* sim_workspace, float* values,
                   int32_t* indices, cudaStream_t stream);

/* ---- CLIP image tower, persistent single-kernel edition (csrc/clip_tower.cu) ---------------------------------
 * Same contract as sc_clip_encode (clip_model.encode_image + F.normalize, CLIP_anno.py:166-167), executed as ONE cooperative
 * launch: every GEMM a persistent warp-specialised tcgen05 pipeline with the accumulator double-buffered in TMEM, LayerNorm
 * folded into the consuming GEMM, attention on mma.sync. cfg->split = 1: hi/lo bf16 operand pairs (fp32-class parity mode);
 * cfg->split = 0: fp16 operands, one MMA per product (the precision the reference itself runs CLIP at on CUDA).
 * The weights arrive PRE-FOLDED (shapeclipper_b200/clip.py::_pack_tower): W' = gain-scaled (and, for the q rows, 1/8-scaled)
 * matrices as 16-bit planes (fp16 when split == 0, bf16 hi/lo when split == 1), s[n] = sum_k W'[n,k] of the ROUNDED planes,
 * c[n] = sum_k ln_bias[k] W[n,k] + bias[n]. */
typedef struct ScClipTowerLayer {
    const void *qkv_w_hi, *qkv_w_lo; const float *qkv_s, *qkv_c;    /* [3W, W] folded with ln_1; s, c [3W] */
    const void *out_w_hi, *out_w_lo; const float* out_b;            /* [W, W] */
    const void *fc1_w_hi, *fc1_w_lo; const float *fc1_s, *fc1_c;    /* [4W, W] folded with ln_2 */
    const void *fc2_w_hi, *fc2_w_lo; const float* fc2_b;            /* [W, 4W] */
} ScClipTowerLayer;
typedef struct ScClipTowerWeights {
    const void *conv_w_hi, *conv_w_lo;                              /* [W, Kp] (columns zero-padded to a multiple of 64) */
    const float *class_emb, *pos_emb, *lnpre_w, *lnpre_b, *lnpost_w, *lnpost_b;
    const float* proj_t;                                            /* fp32 proj^T [out_dim, W] */
    const ScClipTowerLayer* layers;                                 /* HOST array of cfg.layers entries */
} ScClipTowerWeights;
size_t sc_clip_tower_workspace_bytes(const ScClipConfig* cfg, int batch);
size_t sc_clip_tower_plan_bytes(const ScClipConfig* cfg);
/* Builds the phase table + every TMA descriptor once per (weights, batch, workspace) into `plan` (device memory). Synchronises
 * `stream`; not capturable. */
int sc_clip_tower_plan(const ScClipConfig* cfg, const ScClipTowerWeights* weights, int batch, void* workspace,
                       size_t workspace_bytes, void* plan, size_t plan_bytes, int* n_phases_out, cudaStream_t stream);
/* Diagnostics: byte offsets inside the workspace of {x fp32, x16 hi, lo, qkv hi, lo, attn hi, lo, h hi, lo, patch hi, lo,
 * patch_out fp32, y fp32, raw fp32, stats a, stats b}. */
int sc_clip_tower_workspace_layout(const ScClipConfig* cfg, int batch, size_t* offsets16);
/* mode 0: one cooperative launch (+ one 4-byte memset); mode 1: one launch per phase (profiling form, same device code);
 * mode k >= 2: per-phase launches of the first k - 1 phases only (diagnostics); mode -g (g > 0): cooperative launches of g phases
 * each, the three embedding phases first (a tower that another stream's kernels can interleave with). */
int sc_clip_tower_encode(const ScClipConfig* cfg, const void* plan, int n_phases, void* workspace, const float* images,
                         float* emb, float* emb_unnormalised, void* emb_hi, void* emb_lo, int mode, cudaStream_t stream);

/* ---- iso-surface extraction + surface sampling (csrc/mcubes.cu) ---------------------------------
 * Replaces the CPU leg of utils/eval_3D.py:123-153: mcubes.marching_cubes(level, isovalue) + trimesh.Trimesh(..).sample(n)
 * (both third-party, absent: tables generated from the definition, sampler restated from trimesh's published algorithm).
 * level [B, n, n, n] fp32 device ([b][ix][iy][iz]); a lattice point is inside when level < isovalue.
 *   sc_mc_set_tables : the case tables of shapeclipper_b200/mcubes_tables.py (HOST pointers), once per device
 *   sc_mc_count      : counts [B * (n-1)^3] int32 = triangles per cell (cell id = ((b (n-1) + x)(n-1) + y)(n-1) + z)
 *   sc_mc_emit       : offsets = exclusive scan of counts (int64); triangles [total, 3, 3] in world units, vertex = index / n *
 *                      (hi - lo) + lo as the reference scales them (utils/eval_3D.py:136-140)
 *   sc_tri_area      : area [n_tris]
 *   sc_tri_sample    : points[i] = uniform point of triangle face[i] from the two uniforms uv[i] (folded when r1 + r2 > 1) */
int sc_mc_set_tables(const int8_t* tri_count, const int8_t* tri_edges, const int8_t* edge_corner);
int sc_mc_count(const float* level, int batch, int n, float isovalue, int32_t* counts, cudaStream_t stream);
int sc_mc_emit(const float* level, int batch, int n, float isovalue, const int64_t* offsets, float lo, float hi, float* triangles,
               cudaStream_t stream);
int sc_tri_area(const float* triangles, int64_t n_tris, float* area, cudaStream_t stream);
int sc_tri_sample(const float* triangles, const int64_t* face, const float* uv, int64_t count, float* points, cudaStream_t stream);

/* ---- boundary-distance ray sampler (csrc/sampler.cu) ---------------------------------------------
 * Replaces the CPU leg of utils/util.py:237-248 (compute_sampling_prob, per image in the DataLoader workers, data/pix3d.py:234-239):
 * vigra.filters.boundaryDistanceTransform(mask > 0.5) and the sampling weights 1 / (distance + uniform_fac).
 * mask [B, H, W] fp32 device (foreground = mask > threshold); dist [B, H, W] fp32 (or NULL) = Euclidean distance of every pixel
 * to the nearest pixel of the OTHER class, minus 0.5 (vigra's default InterpixelBoundary); an image without a pixel of the other
 * class: H + W. keys [B, H, W] (or NULL; needs uniforms [B, H, W] in (0, 1]) = -log(u) (dist + uniform_fac): the n smallest keys of
 * an image are a sample of n pixels without replacement with probability proportional to 1 / (dist + uniform_fac), i.e. what
 * np.random.choice(H W, n, p = prob, replace = False) draws (utils/util.py:247). scratch: sc_boundary_distance_scratch_bytes.
 * H, W <= 23170; exact (integer d^2, correctly rounded sqrt) while H^2 + W^2 < 2^24. */
size_t sc_boundary_distance_scratch_bytes(int batch, int H, int W);
int sc_boundary_distance(const float* mask, int batch, int H, int W, float threshold, void* scratch, float* dist,
                         const float* uniforms, float uniform_fac, float* keys, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SC_B200_H_ */
