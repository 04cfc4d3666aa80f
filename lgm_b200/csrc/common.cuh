// common.cuh — shared declarations of the sm_100a splat-render library (internal; the public C-ABI is include/lgm_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace lgm {

// ---- per-device host-side state.  A process may drive several GPUs (one thread or stream per device): function
// attributes and SM counts are properties of the CURRENT device, so every cache below is keyed by cudaGetDevice(). ----
constexpr int kMaxDevices = 64;
inline int current_device()
{
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < kMaxDevices) ? d : 0;
}
// SM count of the current device (cached per device; api.cu)
int device_sm_count();
// cudaFuncAttributeMaxDynamicSharedMemorySize opt-in, once per (kernel, device).  `done` is one mask per kernel
// (a function-local static at the call site); setting the attribute twice from racing threads is harmless.
template <typename Kernel>
inline cudaError_t opt_in_dynamic_smem(Kernel kernel, size_t bytes, std::atomic<uint64_t>& done)
{
    const uint64_t bit = 1ull << current_device();
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// ---- tuning / test hooks (lgm_set_tuning in include/lgm_b200.h; process-wide, read at launch time; api.cu) ----
enum Tuning { kTuneFwdBatch = 0, kTunePatchLanes, kTuneBwdBatch, kTuneSortVariant, kTuneEnumGlobal, kTuneCoarseRatio, kTuneC2Occ, kTuneSortBulk, kTuneSparseLanes, kTuneFineTileMajor, kTuneCount };
int tuning(Tuning which);  // < 0: not set, the kernel's built-in default applies

struct RenderParams {
    int n_scenes;   // B
    int P;          // Gaussians per scene
    int n_views;    // views rendered by this call (all scenes)
    int H, W;
    int gx, gy, n_tiles;  // tile grid per view
    float tanx, tany, fx, fy, mod;
};

// Per-(view, Gaussian) gradient row written by the backward compositing kernel and consumed by the preprocess
// backward: [0:2] dL/dmean2D (NDC-scaled), [2:5] dL/dconic (xx, xy, yy), [5] dL/dopacity, [6:9] dL/dcolour,
// [9] dL/ddepth, [10:12] padding (keeps rows 16-byte aligned for vector atomics / loads).
constexpr int kGradRow = 12;

constexpr int kBlock = 256;  // threads per block of the per-Gaussian kernels and of a 16x16 tile

// ---- launchers (each enqueues on `stream`, returns the launch status, never synchronises) ----
cudaError_t launch_preprocess_fwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                  const float* view_mats, const float* proj_mats, const int32_t* view_scene,
                                  float* depth, int32_t* radii, float2* xy, float4* conic_opacity,
                                  uint32_t* tiles_touched, uint32_t* block_sums, const float* cov3d = nullptr,
                                  float* zero_rows = nullptr);
cudaError_t launch_mark_visible(cudaStream_t stream, int P, const float* means, const float* view_mat, uint8_t* visible);
// direct_bin.cu: count -> scan -> scatter -> per-tile shared-memory sort (no global radix sort); a step whose longest
// tile exceeds direct_bin_tile_cap() must use the onesweep path
int direct_bin_tile_cap();
size_t direct_bin_scratch_bytes(const RenderParams& prm);
cudaError_t launch_direct_bin_count(cudaStream_t stream, const RenderParams& prm, const int32_t* radii, const float2* xy,
                                    uint2* ranges, void* scratch, uint32_t* longest_out, uint32_t* entries_out,
                                    const unsigned long long* total_instances);
bool direct_bin_use_coarse(const RenderParams& prm, uint64_t n_instances, uint64_t coarse_entries);
cudaError_t launch_direct_bin_sort(cudaStream_t stream, const RenderParams& prm, const int32_t* radii, const float2* xy,
                                   const float* depth, const uint2* ranges, void* pairs, uint32_t* vals_sorted,
                                   uint64_t* keys_sorted, void* scratch, uint32_t longest_tile, void* entries,
                                   uint32_t instances_per_entry = 0);
// loss.cu: MSE(image) + MSE(alpha) and its gradient in one pass (core/models.py:153)
cudaError_t launch_mse_loss_grad(cudaStream_t stream, const float* image, const float* gt_image, float* d_image, size_t n_img,
                                 float w_img, const float* alpha, const float* gt_alpha, float* d_alpha, size_t n_alpha,
                                 float w_alpha, double* loss, const float* grad_scale);
// activations.cu: raw splatter image [.,14] -> Gaussians (core/models.py:40-44,107-115), forward and backward
cudaError_t launch_activate_fwd(cudaStream_t stream, size_t n_scenes, size_t n_per_scene, const float* x, float* g, double* cols);
cudaError_t launch_activate_bwd(cudaStream_t stream, size_t n_scenes, size_t n_per_scene, const float* x, const float* dg,
                                float* dx, double* cols);
// resize.cu: y = mul * bilinear_resize(x) + add (F.interpolate, align_corners = False) and its backward
cudaError_t launch_resize_bilinear_fwd(cudaStream_t stream, const float* x, float* y, int n_planes, int h_in, int w_in, int h_out,
                                       int w_out, float mul, float add);
cudaError_t launch_resize_bilinear_bwd(cudaStream_t stream, const float* dy, float* dx, int n_planes, int h_in, int w_in, int h_out,
                                       int w_out, float mul);
cudaError_t launch_mse_loss_grad_u8(cudaStream_t stream, const float* image, const uint8_t* gt_image, float* d_image, size_t n_img,
                                    float w_img, const float* alpha, const uint8_t* gt_alpha, float* d_alpha, size_t n_alpha,
                                    float w_alpha, double* loss, const float* grad_scale);
// sh.cu: the `shs` input of the Level-1 API (view-dependent colour, degrees 0..3) and its backward
cudaError_t launch_sh_forward(cudaStream_t stream, int P, int deg, int max_coeffs, const float* means, const float* campos,
                              const float* shs, float* colors, uint8_t* clamped);
cudaError_t launch_sh_backward(cudaStream_t stream, int P, int deg, int max_coeffs, const float* means, const float* campos,
                               const float* shs, const uint8_t* clamped, const float* dL_dcolor, float* dL_dshs,
                               float* dL_dmeans);
cudaError_t launch_preprocess_bwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                  const float* view_mats, const float* proj_mats, const int32_t* scene_view_offsets,
                                  const int32_t* radii, const float4* conic_opacity, const float* grad_rows,
                                  float* dL_dgaussians, int accumulate, const float* cov3d = nullptr,
                                  float* dL_dcov3d = nullptr);
cudaError_t launch_screen_gradients(cudaStream_t stream, const RenderParams& prm, const float4* conic_opacity,
                                    const float* grad_rows, float* out);

cudaError_t launch_scan_block_sums(cudaStream_t stream, const uint32_t* block_sums, uint32_t n_views, uint32_t per_view,
                                   uint32_t* block_offsets, unsigned long long* total);
cudaError_t launch_emit(cudaStream_t stream, const RenderParams& prm, const int32_t* radii, const float2* xy,
                        const float* depth, const uint32_t* block_offsets, uint64_t* keys, uint32_t* vals);
cudaError_t launch_tile_ranges(cudaStream_t stream, const uint64_t* keys, uint32_t L, uint2* ranges);

// onesweep radix sort of (u64 key, u32 value) pairs, stable, on key bits [begin_bit, end_bit) (of the compressed key when
// compress != 0).  The sorted result lands in keys_out / vals_out; keys_tmp / vals_tmp are the alternate buffers;
// sort_input_is_tmp(begin_bit, end_bit) tells the caller which of the two to fill with the unsorted data.
int sort_num_passes(int begin_bit, int end_bit);
bool sort_input_is_tmp(int begin_bit, int end_bit);
size_t sort_scratch_bytes(uint32_t n, int begin_bit, int end_bit);
cudaError_t launch_onesweep_sort(cudaStream_t stream, uint64_t* keys_out, uint32_t* vals_out, uint64_t* keys_tmp,
                                 uint32_t* vals_tmp, uint32_t n, int begin_bit, int end_bit, int compress, void* scratch,
                                 size_t scratch_bytes);
// per-tile (segment) stable sort of the instances on the 31 depth bits, in shared memory; keys must already be grouped
// by global tile (ranges filled).  write_keys == 0 leaves keys_sorted grouped by tile only (vals are always sorted).
size_t tile_sort_scratch_bytes(uint32_t n_ranges);
cudaError_t launch_tile_depth_sort(cudaStream_t stream, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                                   const uint2* ranges, uint32_t n_ranges, int write_keys, void* scratch);

cudaError_t launch_composite_fwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                 const int32_t* view_scene, const float2* xy, const float4* conic_opacity,
                                 const float* depth, const uint32_t* vals, const uint2* ranges, const float* bg,
                                 int clamp_image, float* image, float* alpha, float* depth_img, uint32_t* n_contrib);
cudaError_t launch_composite_bwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                 const int32_t* view_scene, const float2* xy, const float4* conic_opacity,
                                 const float* depth, const uint32_t* vals, const uint2* ranges, const float* bg,
                                 const float* alpha, const uint32_t* n_contrib, const float* dL_dimage, const float* dL_dalpha,
                                 const float* dL_ddepth, float* grad_rows);

// composite2.cu: the same two kernels with two pixels per lane and packed fp32 (the default; see launch_composite_fwd)
cudaError_t launch_composite2_fwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                  const int32_t* view_scene, const float2* xy, const float4* conic_opacity,
                                  const float* depth, const uint32_t* vals, const uint2* ranges, const float* bg,
                                  int clamp_image, float* image, float* alpha, float* depth_img, uint32_t* n_contrib);
cudaError_t launch_composite2_bwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                  const int32_t* view_scene, const float2* xy, const float4* conic_opacity,
                                  const float* depth, const uint32_t* vals, const uint2* ranges, const float* bg,
                                  const float* alpha, const uint32_t* n_contrib, const float* dL_dimage, const float* dL_dalpha,
                                  const float* dL_ddepth, float* grad_rows);

// ---- small device helpers ----
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// exclusive scan of one value per thread over a 256-thread block; `total` (optional) receives the block sum.
// s_warp must hold 8 uint32.  Contains two __syncthreads.
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint32_t c = s_warp[w];
        if (w < warp) wbase += c;
        tot += c;
    }
    __syncthreads();
    if (total) *total = tot;
    return wbase + incl - v;
}

}  // namespace lgm
