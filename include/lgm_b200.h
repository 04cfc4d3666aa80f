/*
 * lgm_b200.h — C-ABI of the sm_100a Gaussian-splat render path (liblgm_b200.so).
 *
 * Drop-in boundary for the device side of LGM's renderer: these entry points are what a binding of
 * `diff_gaussian_rasterization._C` would call in place of the external rasterizer's pybind exports
 *   rasterize_gaussians            (called from /root/reference/core/gs.py:76-85 via GaussianRasterizer.forward)
 *   rasterize_gaussians_backward   (autograd backward of the same call, driven by /root/reference/main.py:102)
 *   mark_visible                   (GaussianRasterizer.markVisible; API surface, not used by LGM)
 * but batched: one call renders ALL B x V views of a step (the Python double loop of /root/reference/core/gs.py:42-93)
 * and one call back-propagates them, accumulating per-Gaussian gradients over the views of a scene.
 *
 * Conventions
 *  - plain pointers and sizes; every pointer is DEVICE memory on the current device unless it says "host";
 *    fp32, contiguous.  The caller owns all memory; the library never allocates, frees or keeps pointers.
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); NO entry point synchronises, copies to the
 *    host or reads the environment.  The two data-dependent sizes of a step (instance count, longest tile) are left
 *    in a device-side lgm_step_counts that the caller reads back ONCE per step (the reference: one blocking
 *    cudaMemcpy per VIEW inside CudaRasterizer::Rasterizer::forward).
 *  - returns 0 on success, a negative lgm_status for an invalid argument, a positive cudaError_t for a CUDA
 *    failure; lgm_last_error_string() describes the last non-zero return of the calling thread.
 *    No C++ exception crosses the ABI.
 *  - gaussians [n_scenes, P, 14]: 0:3 position, 3 opacity, 4:7 scale, 7:11 rotation (w,x,y,z; used as given),
 *    11:14 rgb  — the channel split of /root/reference/core/gs.py:45-49.
 *  - view_mats / proj_mats [n_views, 16]: cam_view / cam_view_proj of /root/reference/core/provider_lvis.py:207-208
 *    flattened row-major (the kernels read m[i + 4k] as row i of the transform, as the external rasterizer does).
 *  - bg [3]: background colour, DEVICE memory (the reference passes a CUDA tensor, /root/reference/core/gs.py:20,63),
 *    read by the kernels so that no host copy / synchronisation is needed.
 *  - view_scene [n_views] int32: scene index of every view; views of one scene must be contiguous.
 *    scene_view_offsets [n_scenes + 1] int32: first view of every scene (exclusive prefix of views per scene).
 */
#ifndef LGM_B200_H
#define LGM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum lgm_status {
    LGM_OK = 0,
    LGM_ERR_NULL_POINTER = -1,
    LGM_ERR_BAD_SHAPE = -2,
    LGM_ERR_WORKSPACE_TOO_SMALL = -3,
    LGM_ERR_TOO_MANY_INSTANCES = -4, /* L >= 2^30 in one call: split the views into chunks */
    LGM_ERR_BAD_VALUE = -5
};

/* What GaussianRasterizationSettings carries (/root/reference/core/gs.py:58-71), for a batch of views. */
typedef struct lgm_render_params {
    int32_t n_scenes;      /* B */
    int32_t n_gaussians;   /* P, per scene */
    int32_t n_views;       /* all views of this call */
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    float scale_modifier;
} lgm_render_params;

#define LGM_GRAD_ROW 12 /* floats per (view, Gaussian) gradient row (moment form, see lgm_backward): 5 moments,
                           opacity 1, rgb 3, depth 1, pad 2 */

/* 2: moment-form gradient rows (lgm_backward_geom takes conic_opacity), want_sorted_keys, direct binning.
 * 3: enqueue-only binning (lgm_forward_count, lgm_step_counts; lgm_forward_bin takes longest_tile / bin_mode),
 *    lgm_set_tuning instead of environment variables, activations with the reference's normalisation axis.
 * 4: lgm_forward_geom_rows (K1 zeroes the gradient rows), tuning "sparse_lanes". */
#define LGM_ABI_VERSION 4
int lgm_abi_version(void);
const char* lgm_last_error_string(void);

/* Tuning / test hooks (process-wide; the defaults are the measured optimum).  name: "fwd_batch" / "bwd_batch" (Gaussians
 * staged per block barrier by the compositing kernels, multiple of 32), "patch_lanes" (32 | 16 | 8), "sort_variant"
 * (launch shape of the onesweep sort), "enum_global" (1: binning enumeration without the per-CTA shared-memory stage),
 * "coarse_ratio" (instances per coarse entry from which the direct binning groups by super-tile first; 0 = never),
 * "c2_occ" (100 x forward + backward resident CTAs per SM of the compositing kernels), "sort_bulk" (form of the per-tile
 * sort of the direct path: 2 = keys read from global memory and grouped in shared memory, the default; 1 = segment staged by
 * a bulk copy (TMA), indices grouped; 0 = the same staged by a load / store loop), "sparse_lanes" (backward compositing:
 * hits with at most this many lanes holding a contributing pixel send their terms with vector reductions instead of the
 * warp reduction; 0 = never), "fine_tile_major" (scatter of the heavy steps: 1 = every fine tile sweeps the super-tile's
 * entries, 0 = every entry walks its tiles; default: by the step's instances per entry).
 * value < 0 restores the default. */
int lgm_set_tuning(const char* name, int32_t value);

/* The step's data-dependent sizes, DEVICE memory, 16 bytes: written by lgm_forward_geom (total_instances) and
 * lgm_forward_count (longest_tile, coarse_entries); the caller reads them back with one 16-byte copy. */
typedef struct lgm_step_counts {
    uint64_t total_instances; /* sum of tiles_touched over all (view, Gaussian) pairs of the call */
    uint32_t longest_tile;    /* instances of the fullest tile */
    uint32_t coarse_entries;  /* (view, Gaussian, 8x8-tile super-tile) entries: what the coarse grouping of steps with large
                                 footprints would hold (0 when the image shape does not allow it) */
} lgm_step_counts;

/* Number of 16x16 tiles of one view. */
int lgm_tiles_per_view(int32_t image_height, int32_t image_width);
/* Number of per-(view, 256-Gaussian block) partial sums forward_geom writes: n_views * ceil(P / 256). */
int64_t lgm_num_block_sums(int32_t n_gaussians, int32_t n_views);
/* Scratch bytes forward_bin needs for L instances and E coarse entries, both as read back from lgm_step_counts
 * (alternate key/value buffers, histograms, look-back state; the direct path's pairs alias them; 16 B per entry when
 * the step takes the coarse grouping). */
int lgm_bin_workspace_bytes(const lgm_render_params* prm, int64_t n_instances, int64_t coarse_entries, size_t* bytes);

/* K1 preprocess + instance-offset scan.  Replaces preprocessCUDA + InclusiveSum (+ its blocking D2H: here the
 * total stays on the device in total_instances[0], read back once per step by the host wrapper).
 * Outputs, each [n_views * P]: depth f32, radii i32 (0 = culled), xy float2, conic_opacity float4;
 * tiles_touched u32 is optional (NULL to skip).  block_offsets [lgm_num_block_sums] u32 (exclusive scan),
 * block_sums same size (scratch).  total_instances: 1 x u64.                                              */
int lgm_forward_geom(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                     const float* proj_mats, const int32_t* view_scene, float* depth, int32_t* radii, float* xy,
                     float* conic_opacity, uint32_t* tiles_touched, uint32_t* block_sums, uint32_t* block_offsets,
                     uint64_t* total_instances);

/* The same with upstream's cov3D_precomp: cov3d [n_scenes, P, 6] (xx, xy, xz, yy, yz, zz) replaces the covariance built
 * from the scale / rotation columns of `gaussians` (which are then ignored); NULL = lgm_forward_geom. */
int lgm_forward_geom_cov3d(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                           const float* proj_mats, const int32_t* view_scene, float* depth, int32_t* radii, float* xy,
                           float* conic_opacity, uint32_t* tiles_touched, uint32_t* block_sums, uint32_t* block_offsets,
                           uint64_t* total_instances, const float* cov3d);

/* The same, and K1 also zeroes the step's gradient rows: grad_rows f32 [n_views * P, LGM_GRAD_ROW] (16-byte aligned), the
 * buffer lgm_backward(_composite) accumulates into — K1 is bound by instruction issue, so the 48 B per pair ride along
 * instead of a separate fill of ~1 GB per step.  NULL = lgm_forward_geom_cov3d (the caller zeroes the rows). */
int lgm_forward_geom_rows(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                          const float* proj_mats, const int32_t* view_scene, float* depth, int32_t* radii, float* xy,
                          float* conic_opacity, uint32_t* tiles_touched, uint32_t* block_sums, uint32_t* block_offsets,
                          uint64_t* total_instances, const float* cov3d, float* grad_rows);

/* Binning, first half (direct path D1 + D2): per-tile instance counts and their scan.  After it `ranges`
 * uint2[n_views * tiles] holds every tile's [start, end) in the final instance list (empty tiles (0,0)) and
 * counts->longest_tile the fullest tile; counts->total_instances must already have been written (lgm_forward_geom, same
 * stream).  count_workspace (lgm_count_workspace_bytes; does not depend on the instance
 * count) must be handed unchanged to lgm_forward_bin.  Enqueue before the step's readback. */
int lgm_count_workspace_bytes(const lgm_render_params* prm, size_t* bytes);
int lgm_forward_count(void* stream, const lgm_render_params* prm, const int32_t* radii, const float* xy, uint32_t* ranges,
                      void* count_workspace, size_t count_workspace_bytes, lgm_step_counts* counts);
/* Longest tile the direct path can order in shared memory (20,480). */
int lgm_direct_bin_tile_cap(void);

/* Binning, second half.  Replaces duplicateWithKeys + SortPairs + identifyTileRanges.
 * n_instances, longest_tile, coarse_entries = the values read back from lgm_step_counts (longest_tile < 0: unknown;
 * coarse_entries 0: no coarse grouping).  Steps with large footprints (>= 6 instances per entry; lgm_set_tuning
 * "coarse_ratio") first group the (view, Gaussian) pairs by 8x8-tile super-tile, so that the scatter of the direct path
 * writes long runs instead of isolated 8-byte pairs (lgm_last_bin_coarse tells).
 * keys_sorted u64[L] (view*tiles+tile << 32 | depth bits), vals_sorted u32[L] (view * P + Gaussian index), ranges
 * uint2[n_views * tiles] = [start,end).  The list order is upstream's: by tile, then depth bits, ties in emit order
 * (ascending value).  Three internal paths produce it bit for bit (lgm_last_bin_mode tells which ran):
 *   direct   count -> scan (lgm_forward_count) -> scatter -> per-tile shared-memory sort.  Taken when bin_mode is
 *            LGM_BIN_AUTO or LGM_BIN_DIRECT, count_workspace is the one lgm_forward_count filled, and
 *            0 <= longest_tile <= lgm_direct_bin_tile_cap();
 *   onesweep emit -> stable LSD onesweep radix sort over the (compressed) 64-bit keys -> ranges (everything else);
 *   hybrid   (LGM_BIN_HYBRID) onesweep over the (view|tile) bits, then a per-tile radix sort of the depth bits.
 * block_offsets is only read by the onesweep / hybrid paths.  want_sorted_keys == 0 lets direct / hybrid skip writing
 * keys_sorted (its contents are then unspecified; on the direct path it may then be NULL); vals_sorted and ranges, all
 * the renderer consumes, are always final.
 * Enqueue-only: nothing is read back. */
#define LGM_BIN_AUTO 0
#define LGM_BIN_ONESWEEP 1
#define LGM_BIN_HYBRID 2
#define LGM_BIN_DIRECT 3
int lgm_forward_bin(void* stream, const lgm_render_params* prm, const int32_t* radii, const float* xy,
                    const float* depth, const uint32_t* block_offsets, int64_t n_instances, int64_t longest_tile,
                    int64_t coarse_entries, int32_t bin_mode, uint64_t* keys_sorted, uint32_t* vals_sorted, uint32_t* ranges,
                    void* workspace, size_t workspace_bytes, void* count_workspace, int32_t want_sorted_keys);

/* K5 compositing.  Replaces renderCUDA fwd.  image [n_views,3,H,W], alpha / depth_img [n_views,H,W], n_contrib u32
 * [n_views,H,W] (bits 0..28: number of list entries the pixel consumed, as upstream).
 * clamp_image == 0: image as upstream writes it (not clamped).  clamp_image != 0: the caller's clamp(0,1) of
 * /root/reference/core/gs.py:87 is fused into the store, and bits 29..31 of n_contrib flag the colour channels whose
 * gradient that clamp blocks; lgm_backward* honours the flags, so d(clamped image) is what it must be given.
 * depth_img may be NULL: the depth image is not computed (LGM computes it and drops it, core/gs.py:76).          */
int lgm_forward_composite(void* stream, const lgm_render_params* prm, const float* gaussians,
                          const int32_t* view_scene, const float* xy, const float* conic_opacity, const float* depth,
                          const uint32_t* vals_sorted, const uint32_t* ranges, const float* bg, int32_t clamp_image,
                          float* image, float* alpha, float* depth_img, uint32_t* n_contrib);

/* forward_bin followed by forward_composite (SURVEY.md §8b level 3, entry 3). */
int lgm_forward_bin_render(void* stream, const lgm_render_params* prm, const float* gaussians,
                           const int32_t* view_scene, const int32_t* radii, const float* xy,
                           const float* conic_opacity, const float* depth, const uint32_t* block_offsets,
                           int64_t n_instances, int64_t longest_tile, int64_t coarse_entries, int32_t bin_mode,
                           uint64_t* keys_sorted, uint32_t* vals_sorted, uint32_t* ranges, void* workspace,
                           size_t workspace_bytes, void* count_workspace, const float* bg, int32_t clamp_image, float* image,
                           float* alpha, float* depth_img, uint32_t* n_contrib);

/* K6 + K7.  Replaces renderCUDA bwd + computeCov2DCUDA + preprocessCUDA bwd.
 * dL_ddepth may be NULL (= zero gradient w.r.t. the depth image: LGM's losses never use depth; a cheaper kernel
 * instantiation runs).  grad_rows [n_views * P, LGM_GRAD_ROW] must be ZERO on entry (K6 accumulates with atomics); on
 * return it holds, per (view, Gaussian), the MOMENT form of the screen-space gradients:
 *   [0..4] Sx, Sy, Sxx, Sxy, Syy = sums over the Gaussian's pixels of q d, q d d^T with q = G dL/dalpha and
 *   d = mean2D - pixel;  [5] dL/dopacity (= sum q);  [6..8] dL/dcolour;  [9] dL/ddepth;  [10..11] unused.
 * K7 turns the moments into upstream's dL/dmean2D and dL/dconic with per-Gaussian coefficients (conic, opacity);
 * lgm_screen_gradients does the same into a separate array for callers that want upstream's values (means2D.grad).
 * dL_dgaussians [n_scenes, P, 14]: per-Gaussian gradients summed over the views of each scene; overwritten when
 * accumulate == 0, added to when accumulate != 0 (view chunks of one step).  It may point into peer-mapped memory
 * of another GPU of the node: K7 then delivers a view-sharded rank's gradient block straight to the rank that owns
 * the Gaussians (lgm_b200/dist.py, peer_gradients).                                                          */
int lgm_backward(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                 const float* proj_mats, const int32_t* view_scene, const int32_t* scene_view_offsets,
                 const int32_t* radii, const float* xy, const float* conic_opacity, const float* depth,
                 const uint32_t* vals_sorted, const uint32_t* ranges, const float* bg, const float* alpha,
                 const uint32_t* n_contrib, const float* dL_dimage, const float* dL_dalpha, const float* dL_ddepth, float* grad_rows,
                 float* dL_dgaussians, int32_t accumulate);

/* The two halves of lgm_backward, exposed for the parity tests. */
int lgm_backward_composite(void* stream, const lgm_render_params* prm, const float* gaussians,
                           const int32_t* view_scene, const float* xy, const float* conic_opacity, const float* depth,
                           const uint32_t* vals_sorted, const uint32_t* ranges, const float* bg, const float* alpha,
                           const uint32_t* n_contrib, const float* dL_dimage, const float* dL_dalpha,
                           const float* dL_ddepth, float* grad_rows);
int lgm_backward_geom(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                      const float* proj_mats, const int32_t* scene_view_offsets, const int32_t* radii,
                      const float* conic_opacity, const float* grad_rows, float* dL_dgaussians, int32_t accumulate);
/* K7 for a forward that used cov3D_precomp: the gradient stops at the covariance — dL_dcov3d [n_scenes, P, 6] (upstream's
 * dL_dcov3D layout: off-diagonal entries carry both symmetric halves) is written (added to when accumulate != 0), the
 * scale / rotation columns of dL_dgaussians are zero.  cov3d == NULL and dL_dcov3d == NULL = lgm_backward_geom. */
int lgm_backward_geom_cov3d(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                            const float* proj_mats, const int32_t* scene_view_offsets, const int32_t* radii,
                            const float* conic_opacity, const float* grad_rows, float* dL_dgaussians, int32_t accumulate,
                            const float* cov3d, float* dL_dcov3d);
/* grad_rows (moment form) -> screen_grads [n_views * P, LGM_GRAD_ROW] in upstream's form: [0..1] dL/dmean2D (what
 * means2D.grad receives upstream), [2..4] dL/dconic (xx, xy, yy), [5] dL/dopacity, [6..8] dL/dcolour, [9] dL/ddepth. */
int lgm_screen_gradients(void* stream, const lgm_render_params* prm, const float* conic_opacity, const float* grad_rows,
                         float* screen_grads);

/* markVisible: visible[i] = !(view-space z <= 0.2).  means [P,3], view_mat [16], visible u8[P]. */
/* Which binning path the calling thread's last lgm_forward_bin took (LGM_BIN_ONESWEEP / _HYBRID / _DIRECT; LGM_BIN_NONE
 * when there was nothing to bin): diagnostics / launch accounting. */
#define LGM_BIN_NONE 0
int lgm_last_bin_mode(void);
int lgm_last_bin_coarse(void); /* 1: the last direct binning grouped the pairs by super-tile first */

int lgm_mark_visible(void* stream, int32_t n_points, const float* means, const float* view_mat, uint8_t* visible);

/* The step right before the path: /root/reference/core/models.py:40-44,107-115 — raw splatter image
 * x [n_scenes, n_per_scene, 14] -> Gaussians of the same shape: pos = clamp(x[0:3],-1,1), opacity = sigmoid(x[3]),
 * scale = 0.1 softplus(x[4:7]), rotation = F.normalize(x[7:11]), rgb = 0.5 tanh(x[11:14]) + 0.5; and its backward
 * (dL_dgaussians -> dL_dx).
 * rot_axis: the reference calls F.normalize WITHOUT a dim argument (models.py:43 `self.rot_act = F.normalize`, :112), so
 * torch's default dim = 1 applies to the [B,N,4] slice: every quaternion COMPONENT is divided by its L2 norm over the N
 * Gaussians of the scene (max(norm, 1e-12)) — LGM_ROT_NORM_REFERENCE, what reference checkpoints were trained under.
 * LGM_ROT_NORM_QUATERNION normalises each quaternion to unit length (dim = -1) instead.
 * col_scratch: n_scenes * 8 doubles of device scratch (column sums), needed for LGM_ROT_NORM_REFERENCE. */
#define LGM_ROT_NORM_REFERENCE 0
#define LGM_ROT_NORM_QUATERNION 1
int lgm_activate_forward(void* stream, int64_t n_scenes, int64_t n_per_scene, const float* x, float* gaussians, int32_t rot_axis,
                         double* col_scratch);
int lgm_activate_backward(void* stream, int64_t n_scenes, int64_t n_per_scene, const float* x, const float* dL_dgaussians,
                          float* dL_dx, int32_t rot_axis, double* col_scratch);

/* The supervision right after the path: /root/reference/core/models.py:153,
 *   loss = mse_loss(pred_images, gt_images) + mse_loss(pred_alphas, gt_masks),
 * and its gradient in one pass:  *loss = w_image sum (image - gt_image)^2 + w_alpha sum (alpha - gt_alpha)^2 (double),
 * d_image = 2 s w_image (image - gt_image), d_alpha = 2 s w_alpha (alpha - gt_alpha) — the dL_dimage / dL_dalpha inputs
 * of lgm_backward; s = *grad_scale (a DEVICE float: autograd's incoming dL/dloss), 1 when grad_scale is NULL.
 * w = 1 / element count gives the reference's mean reduction.  loss, d_image and d_alpha may each be NULL (output not
 * wanted).  All pointers 16-byte aligned. */
int lgm_mse_loss_grad(void* stream, const float* image, const float* gt_image, float* d_image, int64_t n_image,
                      float w_image, const float* alpha, const float* gt_alpha, float* d_alpha, int64_t n_alpha, float w_alpha,
                      double* loss, const float* grad_scale);

/* The LPIPS input preparation of /root/reference/core/models.py:155-163:
 *   F.interpolate(images.view(-1, 3, S, S) * 2 - 1, (256, 256), mode='bilinear', align_corners=False)
 * y [n_planes, h_out, w_out] = mul * bilinear_resize(x [n_planes, h_in, w_in]) + add (mul = 2, add = -1 there), and its
 * backward dx = mul * resize^T(dy) (dx is overwritten).  n_planes <= 65535 per call. */
int lgm_resize_bilinear_forward(void* stream, const float* x, float* y, int64_t n_planes, int32_t h_in, int32_t w_in,
                                int32_t h_out, int32_t w_out, float mul, float add);
int lgm_resize_bilinear_backward(void* stream, const float* dy, float* dx, int64_t n_planes, int32_t h_in, int32_t w_in,
                                 int32_t h_out, int32_t w_out, float mul);

/* The same with 8-bit ground truth (what an image file holds; a quarter of the host->device bytes): gt = value / 255.
 * gt_image / gt_alpha 4-byte aligned. */
int lgm_mse_loss_grad_u8(void* stream, const float* image, const uint8_t* gt_image, float* d_image, int64_t n_image,
                         float w_image, const float* alpha, const uint8_t* gt_alpha, float* d_alpha, int64_t n_alpha,
                         float w_alpha, double* loss, const float* grad_scale);

/* Colours from spherical harmonics — the `shs` argument of GaussianRasterizer.forward
 * (diff_gaussian_rasterization/__init__.py: shs / sh_degree / campos; upstream computeColorFromSH in
 * cuda_rasterizer/forward.cu and its backward in backward.cu).  LGM itself passes colors_precomp
 * (core/gs.py:79-80); this exists so that the rasterizer class is a full drop-in.
 *   means [P,3], campos [3], shs [P, max_coeffs, 3] (max_coeffs >= (degree+1)^2, degree in 0..3),
 *   colors [P,3] = max(0.5 + SH(dir), 0), clamped u8[P,3] = the clamp mask the backward needs.
 * Backward: dL_dcolor [P,3] -> dL_dshs [P, max_coeffs, 3] (inactive bands zero) and dL_dmeans [P,3] (through the
 * normalised view direction), both overwritten. */
int lgm_sh_forward(void* stream, int32_t n_points, int32_t degree, int32_t max_coeffs, const float* means,
                   const float* campos, const float* shs, float* colors, uint8_t* clamped);
int lgm_sh_backward(void* stream, int32_t n_points, int32_t degree, int32_t max_coeffs, const float* means,
                    const float* campos, const float* shs, const uint8_t* clamped, const float* dL_dcolor,
                    float* dL_dshs, float* dL_dmeans);

/* The sort on its own (parity / benchmark hook): sorts n pairs, stably, on key bits [0, end_bit).  The unsorted
 * input must be in (keys_tmp, vals_tmp) when lgm_sort_input_is_tmp(end_bit) != 0, else in (keys_out, vals_out);
 * the result is always in (keys_out, vals_out).  compress != 0: every key has bit 31 clear (a positive float in
 * the low word) and the sort runs on the 63-bit value key[30:0] | key[63:32] << 31 — end_bit counts THOSE bits —
 * which is the renderer's configuration (same order as the full key, one pass fewer at 208 views x 400 tiles). */
int lgm_sort_input_is_tmp(int32_t end_bit);
int lgm_sort_workspace_bytes(int64_t n, int32_t end_bit, size_t* bytes);
int lgm_sort_pairs(void* stream, uint64_t* keys_out, uint32_t* vals_out, uint64_t* keys_tmp, uint32_t* vals_tmp,
                   int64_t n, int32_t end_bit, int32_t compress, void* workspace, size_t workspace_bytes);

#ifdef __cplusplus
}
#endif
#endif /* LGM_B200_H */
