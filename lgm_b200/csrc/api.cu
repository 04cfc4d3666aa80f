// api.cu — the C-ABI of include/lgm_b200.h: argument validation, workspace carving, launch sequencing.
// No device memory is allocated here and NO entry point synchronises or copies to the host: every call only enqueues
// work on the caller's stream (the step's single host readback — instance count + longest tile — is the caller's).
// No exception leaves these functions.
#include "../../include/lgm_b200.h"
#include "common.cuh"
#include <stdio.h>
#include <string.h>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* what)
{
    snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}
int fail_cuda(cudaError_t e, const char* where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}
#define LGM_CUDA(call, where)                               \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return fail_cuda(e__, where); \
    } while (0)
#define LGM_NOTNULL(p)                                                         \
    do {                                                                       \
        if ((p) == nullptr) return fail(LGM_ERR_NULL_POINTER, "null pointer: " #p); \
    } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int make_params(const lgm_render_params* in, lgm::RenderParams& p)
{
    if (!in) return fail(LGM_ERR_NULL_POINTER, "null pointer: prm");
    if (in->n_scenes < 0 || in->n_gaussians < 0 || in->n_views < 0 || in->image_height <= 0 || in->image_width <= 0)
        return fail(LGM_ERR_BAD_SHAPE, "negative size or empty image in lgm_render_params");
    if (!(in->tanfovx > 0.f) || !(in->tanfovy > 0.f)) return fail(LGM_ERR_BAD_VALUE, "tanfovx / tanfovy must be > 0");
    // the per-(view, Gaussian) kernels put the view in gridDim.y
    if (in->n_views > 65535) return fail(LGM_ERR_BAD_SHAPE, "n_views must be <= 65535 per call: split the views into chunks");
    if ((int64_t)in->n_views * in->n_gaussians >= (int64_t)1 << 32)
        return fail(LGM_ERR_BAD_SHAPE, "n_views * n_gaussians must be < 2^32 per call: split the views into chunks");
    p.n_scenes = in->n_scenes;
    p.P = in->n_gaussians;
    p.n_views = in->n_views;
    p.H = in->image_height;
    p.W = in->image_width;
    p.gx = (p.W + 15) / 16;
    p.gy = (p.H + 15) / 16;
    p.n_tiles = p.gx * p.gy;
    if ((int64_t)p.n_views * p.n_tiles >= (int64_t)1 << 31)
        return fail(LGM_ERR_BAD_SHAPE, "n_views * tiles must be < 2^31 per call: split the views into chunks");
    p.tanx = in->tanfovx;
    p.tany = in->tanfovy;
    p.fx = (float)p.W / (2.0f * p.tanx);  // A.0 focal lengths
    p.fy = (float)p.H / (2.0f * p.tany);
    p.mod = in->scale_modifier;
    return LGM_OK;
}

int bits_for(uint64_t n_minus_1)
{
    int b = 0;
    while (n_minus_1) { b++; n_minus_1 >>= 1; }
    return b;
}
// sort bits of the compressed key (radix_sort.cu): 31 depth bits (the sign bit of a depth > 0.2 is always 0) + the
// bits of the largest global tile id
int key_end_bit(const lgm::RenderParams& p)
{
    const uint64_t gtiles = (uint64_t)p.n_views * (uint64_t)p.n_tiles;
    return 31 + (gtiles > 1 ? bits_for(gtiles - 1) : 1);
}

struct BinWorkspace {
    size_t keys_tmp, vals_tmp, sort_scratch, sort_scratch_bytes, tile_scratch, entries, total;
};
// E: coarse entries of the step (0: none) — a step that takes the coarse grouping needs 16 B per entry on top
BinWorkspace bin_layout(const lgm::RenderParams& p, uint32_t L, uint64_t E)
{
    BinWorkspace w;
    size_t off = 0;
    w.keys_tmp = off; off = align_up(off + (size_t)L * 8, 256);
    w.vals_tmp = off; off = align_up(off + (size_t)L * 4, 256);
    w.sort_scratch = off;
    w.sort_scratch_bytes = lgm::sort_scratch_bytes(L, 0, key_end_bit(p));  // the full sort needs the most
    off = align_up(off + w.sort_scratch_bytes, 256);
    w.tile_scratch = off;
    off = align_up(off + lgm::tile_sort_scratch_bytes((uint32_t)((size_t)p.n_views * p.n_tiles)), 256);
    w.entries = off;
    if (lgm::direct_bin_use_coarse(p, L, E)) off = align_up(off + (size_t)E * 16, 256);
    w.total = off;
    return w;
}

thread_local int g_last_bin_mode = LGM_BIN_NONE;
thread_local bool g_last_coarse = false;

std::atomic<int> g_tuning[lgm::kTuneCount] = {{-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}, {-1}};
const char* const kTuningNames[lgm::kTuneCount] = {"fwd_batch", "patch_lanes", "bwd_batch", "sort_variant", "enum_global", "coarse_ratio", "c2_occ", "sort_bulk", "sparse_lanes", "fine_tile_major"};
std::atomic<int> g_sm_count[lgm::kMaxDevices];

}  // namespace

namespace lgm {
int tuning(Tuning which) { return g_tuning[which].load(std::memory_order_relaxed); }
int device_sm_count()
{
    const int dev = current_device();
    int n = g_sm_count[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
}  // namespace lgm

extern "C" {

int lgm_abi_version(void) { return LGM_ABI_VERSION; }
int lgm_last_bin_mode(void) { return g_last_bin_mode; }
int lgm_last_bin_coarse(void) { return g_last_coarse ? 1 : 0; }
const char* lgm_last_error_string(void) { return g_err; }

int lgm_set_tuning(const char* name, int32_t value)
{
    LGM_NOTNULL(name);
    for (int i = 0; i < lgm::kTuneCount; i++)
        if (strcmp(name, kTuningNames[i]) == 0) {
            g_tuning[i].store(value, std::memory_order_relaxed);
            return LGM_OK;
        }
    return fail(LGM_ERR_BAD_VALUE, "lgm_set_tuning: unknown name (fwd_batch, bwd_batch, patch_lanes, sort_variant, enum_global, coarse_ratio, c2_occ, sort_bulk, sparse_lanes, fine_tile_major)");
}

int lgm_direct_bin_tile_cap(void) { return lgm::direct_bin_tile_cap(); }

int lgm_tiles_per_view(int32_t H, int32_t W) { return ((W + 15) / 16) * ((H + 15) / 16); }
int64_t lgm_num_block_sums(int32_t P, int32_t n_views) { return (int64_t)n_views * ((P + lgm::kBlock - 1) / lgm::kBlock); }

int lgm_bin_workspace_bytes(const lgm_render_params* prm, int64_t n_instances, int64_t coarse_entries, size_t* bytes)
{
    lgm::RenderParams p;
    if (int rc = make_params(prm, p)) return rc;
    LGM_NOTNULL(bytes);
    if (n_instances < 0) return fail(LGM_ERR_BAD_SHAPE, "n_instances < 0");
    if (n_instances >= ((int64_t)1 << 30)) return fail(LGM_ERR_TOO_MANY_INSTANCES, "n_instances >= 2^30: split the views into chunks");
    if (coarse_entries < 0) return fail(LGM_ERR_BAD_SHAPE, "coarse_entries < 0");
    *bytes = bin_layout(p, (uint32_t)n_instances, (uint64_t)coarse_entries).total;
    return LGM_OK;
}

int lgm_forward_geom(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                     const float* proj_mats, const int32_t* view_scene, float* depth, int32_t* radii, float* xy,
                     float* conic_opacity, uint32_t* tiles_touched, uint32_t* block_sums, uint32_t* block_offsets,
                     uint64_t* total_instances)
{
    return lgm_forward_geom_cov3d(stream, prm, gaussians, view_mats, proj_mats, view_scene, depth, radii, xy, conic_opacity,
                                  tiles_touched, block_sums, block_offsets, total_instances, nullptr);
}

int lgm_forward_geom_cov3d(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                           const float* proj_mats, const int32_t* view_scene, float* depth, int32_t* radii, float* xy,
                           float* conic_opacity, uint32_t* tiles_touched, uint32_t* block_sums, uint32_t* block_offsets,
                           uint64_t* total_instances, const float* cov3d)
{
    return lgm_forward_geom_rows(stream, prm, gaussians, view_mats, proj_mats, view_scene, depth, radii, xy, conic_opacity,
                                 tiles_touched, block_sums, block_offsets, total_instances, cov3d, nullptr);
}

int lgm_forward_geom_rows(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                          const float* proj_mats, const int32_t* view_scene, float* depth, int32_t* radii, float* xy,
                          float* conic_opacity, uint32_t* tiles_touched, uint32_t* block_sums, uint32_t* block_offsets,
                          uint64_t* total_instances, const float* cov3d, float* grad_rows)
{
    lgm::RenderParams p;
    if (int rc = make_params(prm, p)) return rc;
    LGM_NOTNULL(total_instances);
    cudaStream_t s = (cudaStream_t)stream;
    if (p.P == 0 || p.n_views == 0) {
        LGM_CUDA(cudaMemsetAsync(total_instances, 0, sizeof(uint64_t), s), "forward_geom: memset total");
        return LGM_OK;
    }
    LGM_NOTNULL(gaussians); LGM_NOTNULL(view_mats); LGM_NOTNULL(proj_mats); LGM_NOTNULL(view_scene);
    LGM_NOTNULL(depth); LGM_NOTNULL(radii); LGM_NOTNULL(xy); LGM_NOTNULL(conic_opacity);
    LGM_NOTNULL(block_sums); LGM_NOTNULL(block_offsets);
    if (reinterpret_cast<uintptr_t>(grad_rows) & 15u) return fail(LGM_ERR_BAD_VALUE, "grad_rows must be 16-byte aligned");
    LGM_CUDA(lgm::launch_preprocess_fwd(s, p, gaussians, view_mats, proj_mats, view_scene, depth, radii,
                                        reinterpret_cast<float2*>(xy), reinterpret_cast<float4*>(conic_opacity),
                                        tiles_touched, block_sums, cov3d, grad_rows),
             "forward_geom: preprocess");
    LGM_CUDA(lgm::launch_scan_block_sums(s, block_sums, (uint32_t)p.n_views, (uint32_t)((p.P + lgm::kBlock - 1) / lgm::kBlock),
                                         block_offsets, reinterpret_cast<unsigned long long*>(total_instances)),
             "forward_geom: scan");
    return LGM_OK;
}

int lgm_count_workspace_bytes(const lgm_render_params* prm, size_t* bytes)
{
    lgm::RenderParams p;
    if (int rc = make_params(prm, p)) return rc;
    LGM_NOTNULL(bytes);
    *bytes = lgm::direct_bin_scratch_bytes(p);
    return LGM_OK;
}

int lgm_forward_count(void* stream, const lgm_render_params* prm, const int32_t* radii, const float* xy, uint32_t* ranges,
                      void* count_workspace, size_t count_workspace_bytes, lgm_step_counts* counts)
{
    lgm::RenderParams p;
    if (int rc = make_params(prm, p)) return rc;
    LGM_NOTNULL(counts);
    cudaStream_t s = (cudaStream_t)stream;
    if (p.P == 0 || p.n_views == 0) {
        LGM_CUDA(cudaMemsetAsync(&counts->longest_tile, 0, 2 * sizeof(uint32_t), s), "forward_count: memset");
        return LGM_OK;
    }
    LGM_NOTNULL(radii); LGM_NOTNULL(xy); LGM_NOTNULL(ranges); LGM_NOTNULL(count_workspace);
    if (count_workspace_bytes < lgm::direct_bin_scratch_bytes(p))
        return fail(LGM_ERR_WORKSPACE_TOO_SMALL, "forward_count: workspace too small (see lgm_count_workspace_bytes)");
    LGM_CUDA(lgm::launch_direct_bin_count(s, p, radii, reinterpret_cast<const float2*>(xy), reinterpret_cast<uint2*>(ranges),
                                          count_workspace, &counts->longest_tile, &counts->coarse_entries,
                                          reinterpret_cast<const unsigned long long*>(&counts->total_instances)),
             "forward_count");
    return LGM_OK;
}

int lgm_forward_bin(void* stream, const lgm_render_params* prm, const int32_t* radii, const float* xy,
                    const float* depth, const uint32_t* block_offsets, int64_t n_instances, int64_t longest_tile,
                    int64_t coarse_entries, int32_t bin_mode, uint64_t* keys_sorted, uint32_t* vals_sorted, uint32_t* ranges,
                    void* workspace, size_t workspace_bytes, void* count_workspace, int32_t want_sorted_keys)
{
    lgm::RenderParams p;
    if (int rc = make_params(prm, p)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_instances < 0) return fail(LGM_ERR_BAD_SHAPE, "n_instances < 0");
    if (n_instances >= ((int64_t)1 << 30)) return fail(LGM_ERR_TOO_MANY_INSTANCES, "n_instances >= 2^30: split the views into chunks");
    if (bin_mode < LGM_BIN_AUTO || bin_mode > LGM_BIN_DIRECT) return fail(LGM_ERR_BAD_VALUE, "bin_mode must be LGM_BIN_AUTO .. LGM_BIN_DIRECT");
    const size_t n_ranges = (size_t)p.n_views * p.n_tiles;
    if (n_ranges) LGM_NOTNULL(ranges);
    g_last_bin_mode = LGM_BIN_NONE;
    g_last_coarse = false;
    // The DIRECT path (direct_bin.cu: count, scan, scatter, per-tile shared-memory sort — no global radix sort, 20 B
    // instead of 152 B of HBM traffic per instance) needs lgm_forward_count to have run (ranges[] are then already
    // final) and every tile to fit its shared-memory sort; the caller read the longest tile back together with the
    // instance count.  Otherwise (tiles beyond the capacity, no count, or a forced mode): the ONESWEEP path — emit, one
    // stable LSD sort of the compressed 64-bit keys, ranges — or HYBRID: onesweep over the (view|tile) bits, then every
    // tile's segment sorted on its 31 depth bits in shared memory (tile_sort.cu; measured slower on B200, kept as a
    // tested alternative).
    const bool direct = (bin_mode == LGM_BIN_AUTO || bin_mode == LGM_BIN_DIRECT) && count_workspace != nullptr &&
                        longest_tile >= 0 && longest_tile <= (int64_t)lgm::direct_bin_tile_cap();
    if (!direct && n_ranges)
        LGM_CUDA(cudaMemsetAsync(ranges, 0, n_ranges * sizeof(uint2), s), "forward_bin: memset ranges");
    if (n_instances == 0 || p.P == 0 || p.n_views == 0) {
        if (direct && n_ranges) LGM_CUDA(cudaMemsetAsync(ranges, 0, n_ranges * sizeof(uint2), s), "forward_bin: memset ranges");
        return LGM_OK;
    }
    LGM_NOTNULL(radii); LGM_NOTNULL(xy); LGM_NOTNULL(depth);
    LGM_NOTNULL(vals_sorted); LGM_NOTNULL(workspace);
    if (!direct || want_sorted_keys) LGM_NOTNULL(keys_sorted);  // the direct path needs no key buffer of its own
    const uint32_t L = (uint32_t)n_instances;
    if (coarse_entries < 0) return fail(LGM_ERR_BAD_SHAPE, "coarse_entries < 0");
    const BinWorkspace w = bin_layout(p, L, direct ? (uint64_t)coarse_entries : 0);
    if (workspace_bytes < w.total) return fail(LGM_ERR_WORKSPACE_TOO_SMALL, "forward_bin: workspace too small (see lgm_bin_workspace_bytes)");
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    uint64_t* keys_tmp = reinterpret_cast<uint64_t*>(ws + w.keys_tmp);
    uint32_t* vals_tmp = reinterpret_cast<uint32_t*>(ws + w.vals_tmp);
    const int end_bit = key_end_bit(p);
    if (direct) {
        // steps with large footprints group the pairs by super-tile first (16 B per entry at the end of the workspace)
        void* entries = nullptr;
        if (lgm::direct_bin_use_coarse(p, (uint64_t)L, (uint64_t)coarse_entries)) entries = ws + w.entries;
        g_last_coarse = entries != nullptr;
        LGM_CUDA(lgm::launch_direct_bin_sort(s, p, radii, reinterpret_cast<const float2*>(xy), depth,
                                             reinterpret_cast<const uint2*>(ranges), keys_tmp, vals_sorted,
                                             want_sorted_keys ? keys_sorted : nullptr, count_workspace, (uint32_t)longest_tile, entries,
                                             coarse_entries > 0 ? (uint32_t)((uint64_t)L / (uint64_t)coarse_entries) : 0u),
                 "forward_bin: direct sort");
        g_last_bin_mode = LGM_BIN_DIRECT;
        return LGM_OK;
    }
    LGM_NOTNULL(block_offsets);
    const bool full = bin_mode != LGM_BIN_HYBRID;
    g_last_bin_mode = full ? LGM_BIN_ONESWEEP : LGM_BIN_HYBRID;
    const int begin_bit = full ? 0 : 31;
    const bool in_tmp = lgm::sort_input_is_tmp(begin_bit, end_bit);
    LGM_CUDA(lgm::launch_emit(s, p, radii, reinterpret_cast<const float2*>(xy), depth, block_offsets,
                              in_tmp ? keys_tmp : keys_sorted, in_tmp ? vals_tmp : vals_sorted),
             "forward_bin: emit");
    LGM_CUDA(lgm::launch_onesweep_sort(s, keys_sorted, vals_sorted, keys_tmp, vals_tmp, L, begin_bit, end_bit, /*compress=*/1,
                                       ws + w.sort_scratch, w.sort_scratch_bytes),
             "forward_bin: sort");
    LGM_CUDA(lgm::launch_tile_ranges(s, keys_sorted, L, reinterpret_cast<uint2*>(ranges)), "forward_bin: ranges");
    if (!full)
        LGM_CUDA(lgm::launch_tile_depth_sort(s, keys_sorted, vals_sorted, keys_tmp, vals_tmp, reinterpret_cast<const uint2*>(ranges),
                                             (uint32_t)n_ranges, want_sorted_keys ? 1 : 0, ws + w.tile_scratch),
                 "forward_bin: tile depth sort");
    return LGM_OK;
}

int lgm_forward_composite(void* stream, const lgm_render_params* prm, const float* gaussians,
                          const int32_t* view_scene, const float* xy, const float* conic_opacity, const float* depth,
                          const uint32_t* vals_sorted, const uint32_t* ranges, const float* bg, int32_t clamp_image,
                          float* image, float* alpha, float* depth_img, uint32_t* n_contrib)
{
    lgm::RenderParams p;
    if (int rc = make_params(prm, p)) return rc;
    if (p.n_views == 0) return LGM_OK;
    LGM_NOTNULL(view_scene); LGM_NOTNULL(ranges); LGM_NOTNULL(bg); LGM_NOTNULL(image); LGM_NOTNULL(alpha);
    LGM_NOTNULL(n_contrib);
    if (p.P > 0) { LGM_NOTNULL(gaussians); LGM_NOTNULL(xy); LGM_NOTNULL(conic_opacity); LGM_NOTNULL(depth); }
    LGM_CUDA(lgm::launch_composite_fwd((cudaStream_t)stream, p, gaussians, view_scene, reinterpret_cast<const float2*>(xy),
                                       reinterpret_cast<const float4*>(conic_opacity), depth, vals_sorted,
                                       reinterpret_cast<const uint2*>(ranges), bg, clamp_image ? 1 : 0, image, alpha, depth_img,
                                       n_contrib),
             "forward_composite");
    return LGM_OK;
}

int lgm_forward_bin_render(void* stream, const lgm_render_params* prm, const float* gaussians,
                           const int32_t* view_scene, const int32_t* radii, const float* xy,
                           const float* conic_opacity, const float* depth, const uint32_t* block_offsets,
                           int64_t n_instances, int64_t longest_tile, int64_t coarse_entries, int32_t bin_mode,
                           uint64_t* keys_sorted, uint32_t* vals_sorted, uint32_t* ranges, void* workspace,
                           size_t workspace_bytes, void* count_workspace, const float* bg, int32_t clamp_image, float* image,
                           float* alpha, float* depth_img, uint32_t* n_contrib)
{
    if (int rc = lgm_forward_bin(stream, prm, radii, xy, depth, block_offsets, n_instances, longest_tile, coarse_entries, bin_mode,
                                 keys_sorted, vals_sorted, ranges, workspace, workspace_bytes, count_workspace, /*want_sorted_keys=*/1))
        return rc;
    return lgm_forward_composite(stream, prm, gaussians, view_scene, xy, conic_opacity, depth, vals_sorted, ranges, bg,
                                 clamp_image, image, alpha, depth_img, n_contrib);
}

int lgm_backward_composite(void* stream, const lgm_render_params* prm, const float* gaussians,
                           const int32_t* view_scene, const float* xy, const float* conic_opacity, const float* depth,
                           const uint32_t* vals_sorted, const uint32_t* ranges, const float* bg, const float* alpha,
                           const uint32_t* n_contrib, const float* dL_dimage, const float* dL_dalpha,
                           const float* dL_ddepth, float* grad_rows)
{
    lgm::RenderParams p;
    if (int rc = make_params(prm, p)) return rc;
    if (p.n_views == 0 || p.P == 0) return LGM_OK;
    LGM_NOTNULL(gaussians); LGM_NOTNULL(view_scene); LGM_NOTNULL(xy); LGM_NOTNULL(conic_opacity); LGM_NOTNULL(depth);
    LGM_NOTNULL(ranges); LGM_NOTNULL(bg); LGM_NOTNULL(alpha); LGM_NOTNULL(n_contrib); LGM_NOTNULL(dL_dimage); LGM_NOTNULL(dL_dalpha);
    LGM_NOTNULL(grad_rows);  // dL_ddepth may be NULL: no gradient w.r.t. the depth image
    // the rows are accumulated with 16-byte vector reductions and read with 16-byte loads
    if (reinterpret_cast<uintptr_t>(grad_rows) & 15u) return fail(LGM_ERR_BAD_VALUE, "grad_rows must be 16-byte aligned");
    LGM_CUDA(lgm::launch_composite_bwd((cudaStream_t)stream, p, gaussians, view_scene, reinterpret_cast<const float2*>(xy),
                                       reinterpret_cast<const float4*>(conic_opacity), depth, vals_sorted,
                                       reinterpret_cast<const uint2*>(ranges), bg, alpha, n_contrib, dL_dimage, dL_dalpha,
                                       dL_ddepth, grad_rows),
             "backward_composite");
    return LGM_OK;
}

int lgm_backward_geom(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                      const float* proj_mats, const int32_t* scene_view_offsets, const int32_t* radii,
                      const float* conic_opacity, const float* grad_rows, float* dL_dgaussians, int32_t accumulate)
{
    return lgm_backward_geom_cov3d(stream, prm, gaussians, view_mats, proj_mats, scene_view_offsets, radii, conic_opacity,
                                   grad_rows, dL_dgaussians, accumulate, nullptr, nullptr);
}

int lgm_backward_geom_cov3d(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                            const float* proj_mats, const int32_t* scene_view_offsets, const int32_t* radii,
                            const float* conic_opacity, const float* grad_rows, float* dL_dgaussians, int32_t accumulate,
                            const float* cov3d, float* dL_dcov3d)
{
    if ((cov3d == nullptr) != (dL_dcov3d == nullptr)) return fail(LGM_ERR_NULL_POINTER, "cov3d and dL_dcov3d go together");
    lgm::RenderParams p;
    if (int rc = make_params(prm, p)) return rc;
    if (p.n_scenes == 0 || p.P == 0) return LGM_OK;
    LGM_NOTNULL(gaussians); LGM_NOTNULL(scene_view_offsets); LGM_NOTNULL(dL_dgaussians);
    if (p.n_views > 0) {
        LGM_NOTNULL(view_mats); LGM_NOTNULL(proj_mats); LGM_NOTNULL(radii); LGM_NOTNULL(conic_opacity); LGM_NOTNULL(grad_rows);
    }
    LGM_CUDA(lgm::launch_preprocess_bwd((cudaStream_t)stream, p, gaussians, view_mats, proj_mats, scene_view_offsets, radii,
                                        reinterpret_cast<const float4*>(conic_opacity), grad_rows, dL_dgaussians, accumulate,
                                        cov3d, dL_dcov3d),
             "backward_geom");
    return LGM_OK;
}

int lgm_screen_gradients(void* stream, const lgm_render_params* prm, const float* conic_opacity, const float* grad_rows,
                         float* screen_grads)
{
    lgm::RenderParams p;
    if (int rc = make_params(prm, p)) return rc;
    if (p.n_views == 0 || p.P == 0) return LGM_OK;
    LGM_NOTNULL(conic_opacity); LGM_NOTNULL(grad_rows); LGM_NOTNULL(screen_grads);
    LGM_CUDA(lgm::launch_screen_gradients((cudaStream_t)stream, p, reinterpret_cast<const float4*>(conic_opacity), grad_rows,
                                          screen_grads),
             "screen_gradients");
    return LGM_OK;
}

int lgm_backward(void* stream, const lgm_render_params* prm, const float* gaussians, const float* view_mats,
                 const float* proj_mats, const int32_t* view_scene, const int32_t* scene_view_offsets,
                 const int32_t* radii, const float* xy, const float* conic_opacity, const float* depth,
                 const uint32_t* vals_sorted, const uint32_t* ranges, const float* bg, const float* alpha,
                 const uint32_t* n_contrib, const float* dL_dimage, const float* dL_dalpha, const float* dL_ddepth,
                 float* grad_rows, float* dL_dgaussians, int32_t accumulate)
{
    if (int rc = lgm_backward_composite(stream, prm, gaussians, view_scene, xy, conic_opacity, depth, vals_sorted, ranges,
                                        bg, alpha, n_contrib, dL_dimage, dL_dalpha, dL_ddepth, grad_rows))
        return rc;
    return lgm_backward_geom(stream, prm, gaussians, view_mats, proj_mats, scene_view_offsets, radii, conic_opacity, grad_rows,
                             dL_dgaussians, accumulate);
}

int lgm_mark_visible(void* stream, int32_t n_points, const float* means, const float* view_mat, uint8_t* visible)
{
    if (n_points < 0) return fail(LGM_ERR_BAD_SHAPE, "n_points < 0");
    if (n_points == 0) return LGM_OK;
    LGM_NOTNULL(means); LGM_NOTNULL(view_mat); LGM_NOTNULL(visible);
    LGM_CUDA(lgm::launch_mark_visible((cudaStream_t)stream, n_points, means, view_mat, visible), "mark_visible");
    return LGM_OK;
}

int lgm_mse_loss_grad(void* stream, const float* image, const float* gt_image, float* d_image, int64_t n_image,
                      float w_image, const float* alpha, const float* gt_alpha, float* d_alpha, int64_t n_alpha, float w_alpha,
                      double* loss, const float* grad_scale)
{
    if (n_image < 0 || n_alpha < 0) return fail(LGM_ERR_BAD_SHAPE, "negative element count");
    if (n_image > 0) { LGM_NOTNULL(image); LGM_NOTNULL(gt_image); }
    if (n_alpha > 0) { LGM_NOTNULL(alpha); LGM_NOTNULL(gt_alpha); }
    const uintptr_t bits = (uintptr_t)image | (uintptr_t)gt_image | (uintptr_t)d_image | (uintptr_t)alpha | (uintptr_t)gt_alpha |
                           (uintptr_t)d_alpha;
    if (bits & 15u) return fail(LGM_ERR_BAD_SHAPE, "mse_loss_grad: pointers must be 16-byte aligned");
    LGM_CUDA(lgm::launch_mse_loss_grad((cudaStream_t)stream, image, gt_image, d_image, (size_t)n_image, w_image, alpha, gt_alpha,
                                       d_alpha, (size_t)n_alpha, w_alpha, loss, grad_scale),
             "mse_loss_grad");
    return LGM_OK;
}

static int activate_shape_ok(int64_t n_scenes, int64_t n_per_scene, int32_t rot_axis, const double* col_scratch)
{
    if (n_scenes < 0 || n_per_scene < 0 || n_scenes > 65535) return fail(LGM_ERR_BAD_SHAPE, "activate: n_scenes must be 0..65535, n_per_scene >= 0");
    if (rot_axis != LGM_ROT_NORM_REFERENCE && rot_axis != LGM_ROT_NORM_QUATERNION) return fail(LGM_ERR_BAD_VALUE, "activate: rot_axis");
    if (rot_axis == LGM_ROT_NORM_REFERENCE && n_scenes * n_per_scene > 0 && col_scratch == nullptr)
        return fail(LGM_ERR_NULL_POINTER, "activate: col_scratch [n_scenes * 8 doubles] is needed for LGM_ROT_NORM_REFERENCE");
    return LGM_OK;
}

int lgm_activate_forward(void* stream, int64_t n_scenes, int64_t n_per_scene, const float* x, float* gaussians, int32_t rot_axis,
                         double* col_scratch)
{
    if (int rc = activate_shape_ok(n_scenes, n_per_scene, rot_axis, col_scratch)) return rc;
    if (n_scenes * n_per_scene == 0) return LGM_OK;
    LGM_NOTNULL(x); LGM_NOTNULL(gaussians);
    LGM_CUDA(lgm::launch_activate_fwd((cudaStream_t)stream, (size_t)n_scenes, (size_t)n_per_scene, x, gaussians,
                                      rot_axis == LGM_ROT_NORM_REFERENCE ? col_scratch : nullptr), "activate_forward");
    return LGM_OK;
}

int lgm_activate_backward(void* stream, int64_t n_scenes, int64_t n_per_scene, const float* x, const float* dL_dgaussians,
                          float* dL_dx, int32_t rot_axis, double* col_scratch)
{
    if (int rc = activate_shape_ok(n_scenes, n_per_scene, rot_axis, col_scratch)) return rc;
    if (n_scenes * n_per_scene == 0) return LGM_OK;
    LGM_NOTNULL(x); LGM_NOTNULL(dL_dgaussians); LGM_NOTNULL(dL_dx);
    LGM_CUDA(lgm::launch_activate_bwd((cudaStream_t)stream, (size_t)n_scenes, (size_t)n_per_scene, x, dL_dgaussians, dL_dx,
                                      rot_axis == LGM_ROT_NORM_REFERENCE ? col_scratch : nullptr), "activate_backward");
    return LGM_OK;
}

static int resize_shape_ok(int64_t n_planes, int32_t h_in, int32_t w_in, int32_t h_out, int32_t w_out)
{
    if (n_planes < 0 || n_planes > 65535) return fail(LGM_ERR_BAD_SHAPE, "n_planes must be 0..65535");
    if (h_in < 1 || w_in < 1 || h_out < 0 || w_out < 0) return fail(LGM_ERR_BAD_SHAPE, "bad image size");
    return LGM_OK;
}

int lgm_resize_bilinear_forward(void* stream, const float* x, float* y, int64_t n_planes, int32_t h_in, int32_t w_in,
                                int32_t h_out, int32_t w_out, float mul, float add)
{
    if (int rc = resize_shape_ok(n_planes, h_in, w_in, h_out, w_out)) return rc;
    if (n_planes == 0 || h_out == 0 || w_out == 0) return LGM_OK;
    LGM_NOTNULL(x); LGM_NOTNULL(y);
    LGM_CUDA(lgm::launch_resize_bilinear_fwd((cudaStream_t)stream, x, y, (int)n_planes, h_in, w_in, h_out, w_out, mul, add),
             "resize_bilinear_forward");
    return LGM_OK;
}

int lgm_resize_bilinear_backward(void* stream, const float* dy, float* dx, int64_t n_planes, int32_t h_in, int32_t w_in,
                                 int32_t h_out, int32_t w_out, float mul)
{
    if (int rc = resize_shape_ok(n_planes, h_in, w_in, h_out, w_out)) return rc;
    if (n_planes == 0) return LGM_OK;
    LGM_NOTNULL(dx);
    if (h_out > 0 && w_out > 0) LGM_NOTNULL(dy);
    LGM_CUDA(lgm::launch_resize_bilinear_bwd((cudaStream_t)stream, dy, dx, (int)n_planes, h_in, w_in, h_out, w_out, mul),
             "resize_bilinear_backward");
    return LGM_OK;
}

int lgm_mse_loss_grad_u8(void* stream, const float* image, const uint8_t* gt_image, float* d_image, int64_t n_image,
                         float w_image, const float* alpha, const uint8_t* gt_alpha, float* d_alpha, int64_t n_alpha,
                         float w_alpha, double* loss, const float* grad_scale)
{
    if (n_image < 0 || n_alpha < 0) return fail(LGM_ERR_BAD_SHAPE, "negative element count");
    if (n_image > 0) { LGM_NOTNULL(image); LGM_NOTNULL(gt_image); }
    if (n_alpha > 0) { LGM_NOTNULL(alpha); LGM_NOTNULL(gt_alpha); }
    if (((uintptr_t)image | (uintptr_t)d_image | (uintptr_t)alpha | (uintptr_t)d_alpha) & 15u)
        return fail(LGM_ERR_BAD_SHAPE, "mse_loss_grad_u8: float pointers must be 16-byte aligned");
    if (((uintptr_t)gt_image | (uintptr_t)gt_alpha) & 3u)
        return fail(LGM_ERR_BAD_SHAPE, "mse_loss_grad_u8: 8-bit ground-truth pointers must be 4-byte aligned");
    LGM_CUDA(lgm::launch_mse_loss_grad_u8((cudaStream_t)stream, image, gt_image, d_image, (size_t)n_image, w_image, alpha,
                                          gt_alpha, d_alpha, (size_t)n_alpha, w_alpha, loss, grad_scale),
             "mse_loss_grad_u8");
    return LGM_OK;
}

static int sh_shape_ok(int32_t n_points, int32_t degree, int32_t max_coeffs)
{
    if (n_points < 0) return fail(LGM_ERR_BAD_SHAPE, "n_points < 0");
    if (degree < 0 || degree > 3) return fail(LGM_ERR_BAD_SHAPE, "sh degree must be 0..3");
    if (max_coeffs < (degree + 1) * (degree + 1)) return fail(LGM_ERR_BAD_SHAPE, "max_coeffs < (degree+1)^2");
    return LGM_OK;
}

int lgm_sh_forward(void* stream, int32_t n_points, int32_t degree, int32_t max_coeffs, const float* means,
                   const float* campos, const float* shs, float* colors, uint8_t* clamped)
{
    if (int rc = sh_shape_ok(n_points, degree, max_coeffs)) return rc;
    if (n_points == 0) return LGM_OK;
    LGM_NOTNULL(means); LGM_NOTNULL(campos); LGM_NOTNULL(shs); LGM_NOTNULL(colors); LGM_NOTNULL(clamped);
    LGM_CUDA(lgm::launch_sh_forward((cudaStream_t)stream, n_points, degree, max_coeffs, means, campos, shs, colors, clamped),
             "sh_forward");
    return LGM_OK;
}

int lgm_sh_backward(void* stream, int32_t n_points, int32_t degree, int32_t max_coeffs, const float* means,
                    const float* campos, const float* shs, const uint8_t* clamped, const float* dL_dcolor,
                    float* dL_dshs, float* dL_dmeans)
{
    if (int rc = sh_shape_ok(n_points, degree, max_coeffs)) return rc;
    if (n_points == 0) return LGM_OK;
    LGM_NOTNULL(means); LGM_NOTNULL(campos); LGM_NOTNULL(shs); LGM_NOTNULL(clamped); LGM_NOTNULL(dL_dcolor);
    LGM_NOTNULL(dL_dshs); LGM_NOTNULL(dL_dmeans);
    LGM_CUDA(lgm::launch_sh_backward((cudaStream_t)stream, n_points, degree, max_coeffs, means, campos, shs, clamped,
                                     dL_dcolor, dL_dshs, dL_dmeans), "sh_backward");
    return LGM_OK;
}

int lgm_sort_input_is_tmp(int32_t end_bit) { return lgm::sort_input_is_tmp(0, end_bit) ? 1 : 0; }

int lgm_sort_workspace_bytes(int64_t n, int32_t end_bit, size_t* bytes)
{
    LGM_NOTNULL(bytes);
    if (n < 0 || n >= ((int64_t)1 << 30)) return fail(LGM_ERR_TOO_MANY_INSTANCES, "sort: n must be in [0, 2^30)");
    if (end_bit < 1 || end_bit > 64) return fail(LGM_ERR_BAD_VALUE, "sort: end_bit must be in [1, 64]");
    *bytes = lgm::sort_scratch_bytes((uint32_t)n, 0, end_bit);
    return LGM_OK;
}

int lgm_sort_pairs(void* stream, uint64_t* keys_out, uint32_t* vals_out, uint64_t* keys_tmp, uint32_t* vals_tmp,
                   int64_t n, int32_t end_bit, int32_t compress, void* workspace, size_t workspace_bytes)
{
    if (n < 0 || n >= ((int64_t)1 << 30)) return fail(LGM_ERR_TOO_MANY_INSTANCES, "sort: n must be in [0, 2^30)");
    if (end_bit < 1 || end_bit > 64) return fail(LGM_ERR_BAD_VALUE, "sort: end_bit must be in [1, 64]");
    if (compress && end_bit > 63) return fail(LGM_ERR_BAD_VALUE, "sort: compressed keys have at most 63 bits");
    if (n == 0) return LGM_OK;
    LGM_NOTNULL(keys_out); LGM_NOTNULL(vals_out); LGM_NOTNULL(keys_tmp); LGM_NOTNULL(vals_tmp); LGM_NOTNULL(workspace);
    if (workspace_bytes < lgm::sort_scratch_bytes((uint32_t)n, 0, end_bit))
        return fail(LGM_ERR_WORKSPACE_TOO_SMALL, "sort: workspace too small (see lgm_sort_workspace_bytes)");
    LGM_CUDA(lgm::launch_onesweep_sort((cudaStream_t)stream, keys_out, vals_out, keys_tmp, vals_tmp, (uint32_t)n, 0, end_bit,
                                       compress ? 1 : 0, workspace, workspace_bytes),
             "sort_pairs");
    return LGM_OK;
}

}  // extern "C"
