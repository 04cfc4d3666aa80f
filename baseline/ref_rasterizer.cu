// ref_rasterizer.cu — TEST / MEASUREMENT INFRASTRUCTURE, not product code.
//
// A deliberately REFERENCE-SHAPED CUDA restatement of the external rasterizer LGM calls
// (ashawkey/diff-gaussian-rasterization, imported at /root/reference/core/gs.py:7-10, used at core/gs.py:58-85).  That
// package is CUDA-only, absent from /root/reference and not installable offline, so the GPU baseline SURVEY.md §7 step
// 0(iii) / §8d and BASELINE.md §3 ask for is written here from SURVEY.md Appendix A (A.0 - A.7), in the package's own
// execution shape — NOT in the shape of lgm_b200/csrc:
//   * ONE VIEW per call, every launch on the legacy default stream;
//   * preprocess (1 thread / Gaussian)  ->  cub::DeviceScan::InclusiveSum  ->  blocking 4-byte cudaMemcpy D2H of the
//     instance count  ->  duplicateWithKeys (serial rect loop per thread)  ->  cub::DeviceRadixSort::SortPairs on
//     32 + msb(tiles) key bits  ->  cudaMemset(ranges) + identifyTileRanges  ->  render: one 16x16 block per tile,
//     256-Gaussian rounds staged in shared memory, expf, __syncthreads_count early exit;
//   * backward: the same tiles walked back to front, TEN global atomicAdd per contributing (pixel, Gaussian) pair
//     (3 colour, 1 depth, 2 mean2D, 3 conic, 1 opacity), then cov2D backward and preprocess backward, one thread
//     per Gaussian; ten zero-filled per-view gradient arrays are the caller's (as RasterizeGaussiansBackwardCUDA);
//   * three caller-resized byte arenas (geometry / binning / image state) carved at 128-byte boundaries in the field
//     order of Appendix A.7, so that keys and ranges can be read out of them exactly as out of the real package.
// The per-Gaussian arithmetic is the pinned sequence of lgm_b200/csrc/splat_math.cuh (the sequence nvcc's default
// contraction gives the upstream source form); the per-pair compositing arithmetic uses libdevice expf and IEEE
// division as a stock build of the package does (no -use_fast_math).
//
// Uses: (1) bench.py's `gpu_baseline` — the denominator of the north_star's ">= 3x the reference rasterizer" target;
// (2) tests/test_gpu_baseline.py — GPU-vs-GPU bitwise comparison of radii / keys / ranges / n_contrib and tolerance
// comparison of images and gradients with the product path.  Nothing under lgm_b200/ links or loads this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <cub/cub.cuh>

#include "../lgm_b200/csrc/splat_math.cuh"

namespace refr {

constexpr int BLOCK_X = 16, BLOCK_Y = 16, BLOCK_SIZE = BLOCK_X * BLOCK_Y, NUM_CHANNELS = 3;

typedef char* (*resize_fn)(void* ctx, size_t bytes);

// ---- arenas (A.7): each field at the next 128-byte boundary ----
template <typename T>
void carve(char*& chunk, T*& ptr, size_t count)
{
    const size_t off = (reinterpret_cast<uintptr_t>(chunk) + 127) & ~(uintptr_t)127;
    ptr = reinterpret_cast<T*>(off);
    chunk = reinterpret_cast<char*>(ptr + count);
}

struct GeometryState {
    size_t scan_size;
    float* depths;
    char* scanning_space;
    bool* clamped;
    int* internal_radii;
    float2* means2D;
    float* cov3D;
    float4* conic_opacity;
    float* rgb;
    uint32_t* point_offsets;
    uint32_t* tiles_touched;
    static GeometryState fromChunk(char*& chunk, size_t P)
    {
        GeometryState g;
        carve(chunk, g.depths, P);
        carve(chunk, g.clamped, P * 3);
        carve(chunk, g.internal_radii, P);
        carve(chunk, g.means2D, P);
        carve(chunk, g.cov3D, P * 6);
        carve(chunk, g.conic_opacity, P);
        carve(chunk, g.rgb, P * 3);
        carve(chunk, g.tiles_touched, P);
        cub::DeviceScan::InclusiveSum(nullptr, g.scan_size, g.tiles_touched, g.tiles_touched, (int)P);
        carve(chunk, g.scanning_space, g.scan_size);
        carve(chunk, g.point_offsets, P);
        return g;
    }
};

struct ImageState {
    uint32_t* n_contrib;
    uint2* ranges;
    static ImageState fromChunk(char*& chunk, size_t N)
    {
        ImageState img;
        carve(chunk, img.n_contrib, N);
        carve(chunk, img.ranges, N);
        return img;
    }
};

struct BinningState {
    size_t sorting_size;
    uint64_t* point_list_keys_unsorted;
    uint64_t* point_list_keys;
    uint32_t* point_list_unsorted;
    uint32_t* point_list;
    char* list_sorting_space;
    static BinningState fromChunk(char*& chunk, size_t L)
    {
        BinningState b;
        carve(chunk, b.point_list, L);
        carve(chunk, b.point_list_unsorted, L);
        carve(chunk, b.point_list_keys, L);
        carve(chunk, b.point_list_keys_unsorted, L);
        cub::DeviceRadixSort::SortPairs(nullptr, b.sorting_size, b.point_list_keys_unsorted, b.point_list_keys,
                                        b.point_list_unsorted, b.point_list, (int)L);
        carve(chunk, b.list_sorting_space, b.sorting_size);
        return b;
    }
};

template <typename T>
size_t required(size_t n)
{
    char* size = nullptr;
    T::fromChunk(size, n);
    return reinterpret_cast<size_t>(size) + 128;
}

// A.2: highest set bit by bisection
uint32_t getHigherMsb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step;
        else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

// ---- A.1: one thread per Gaussian ----
__global__ void preprocessCUDA(int P, const float* __restrict__ means3D, const float* __restrict__ scales, float scale_modifier,
                               const float* __restrict__ rotations, const float* __restrict__ opacities,
                               const float* __restrict__ colors_precomp, const float* __restrict__ viewmatrix,
                               const float* __restrict__ projmatrix, int W, int H, float tan_fovx, float tan_fovy, float focal_x,
                               float focal_y, int* radii, float2* points_xy_image, float* depths, float* cov3Ds, float* rgb,
                               float4* conic_opacity, dim3 grid, uint32_t* tiles_touched)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    radii[idx] = 0;
    tiles_touched[idx] = 0;
    float mv[16], mp[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mv[k] = viewmatrix[k]; mp[k] = projmatrix[k]; }
    const lgm::Geom o = lgm::preprocess_point(means3D + 3 * idx, scales + 3 * idx, rotations + 4 * idx, scale_modifier, mv, mp, W, H,
                                              tan_fovx, tan_fovy, focal_x, focal_y, (int)grid.x, (int)grid.y);
    if (o.radius <= 0) return;
    float cov6[6], M[9];
    lgm::cov3d_from_scale_rot(scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2], scale_modifier, rotations[4 * idx],
                              rotations[4 * idx + 1], rotations[4 * idx + 2], rotations[4 * idx + 3], cov6, M);
#pragma unroll
    for (int k = 0; k < 6; k++) cov3Ds[6 * idx + k] = cov6[k];
    rgb[3 * idx + 0] = colors_precomp[3 * idx + 0];
    rgb[3 * idx + 1] = colors_precomp[3 * idx + 1];
    rgb[3 * idx + 2] = colors_precomp[3 * idx + 2];
    depths[idx] = o.depth;
    radii[idx] = o.radius;
    points_xy_image[idx] = make_float2(o.px, o.py);
    conic_opacity[idx] = make_float4(o.cx, o.cy, o.cz, opacities[idx]);
    tiles_touched[idx] = o.tiles;
}

// ---- A.2: one thread per Gaussian, serial loop over its tile rect ----
__global__ void duplicateWithKeys(int P, const float2* points_xy, const float* depths, const uint32_t* offsets,
                                  uint64_t* gaussian_keys_unsorted, uint32_t* gaussian_values_unsorted, const int* radii, dim3 grid)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    if (radii[idx] > 0) {
        uint32_t off = (idx == 0) ? 0 : offsets[idx - 1];
        int x0, y0, x1, y1;
        lgm::tile_rect(points_xy[idx].x, points_xy[idx].y, radii[idx], (int)grid.x, (int)grid.y, x0, y0, x1, y1);
        for (int y = y0; y < y1; y++) {
            for (int x = x0; x < x1; x++) {
                uint64_t key = (uint64_t)(y * grid.x + x);
                key <<= 32;
                key |= *reinterpret_cast<const uint32_t*>(&depths[idx]);
                gaussian_keys_unsorted[off] = key;
                gaussian_values_unsorted[off] = (uint32_t)idx;
                off++;
            }
        }
    }
}

// ---- A.3 ----
__global__ void identifyTileRanges(int L, const uint64_t* point_list_keys, uint2* ranges)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t currtile = (uint32_t)(point_list_keys[idx] >> 32);
    if (idx == 0) {
        ranges[currtile].x = 0;
    } else {
        const uint32_t prevtile = (uint32_t)(point_list_keys[idx - 1] >> 32);
        if (currtile != prevtile) {
            ranges[prevtile].y = idx;
            ranges[currtile].x = idx;
        }
    }
    if (idx == L - 1) ranges[currtile].y = L;
}

// ---- A.4: block = tile, thread = pixel ----
__global__ void __launch_bounds__(BLOCK_SIZE)
renderCUDA(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
           const float2* __restrict__ points_xy_image, const float* __restrict__ features, const float* __restrict__ depths,
           const float4* __restrict__ conic_opacity, float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib,
           const float* __restrict__ bg_color, float* __restrict__ out_color, float* __restrict__ out_depth)
{
    const uint32_t horizontal_blocks = (W + BLOCK_X - 1) / BLOCK_X;
    const uint2 pix_min = {blockIdx.x * BLOCK_X, blockIdx.y * BLOCK_Y};
    const uint2 pix = {pix_min.x + threadIdx.x, pix_min.y + threadIdx.y};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    bool done = !inside;
    const uint2 range = ranges[blockIdx.y * horizontal_blocks + blockIdx.x];
    const int rounds = ((range.y - range.x + BLOCK_SIZE - 1) / BLOCK_SIZE);
    int toDo = range.y - range.x;

    __shared__ int collected_id[BLOCK_SIZE];
    __shared__ float2 collected_xy[BLOCK_SIZE];
    __shared__ float4 collected_conic_opacity[BLOCK_SIZE];

    float T = 1.0f;
    uint32_t contributor = 0, last_contributor = 0;
    float C[NUM_CHANNELS] = {0};
    float weight = 0, D = 0;
    const int tid = threadIdx.y * BLOCK_X + threadIdx.x;

    for (int i = 0; i < rounds; i++, toDo -= BLOCK_SIZE) {
        const int num_done = __syncthreads_count(done);
        if (num_done == BLOCK_SIZE) break;
        const int progress = i * BLOCK_SIZE + tid;
        if (range.x + progress < range.y) {
            const int coll_id = point_list[range.x + progress];
            collected_id[tid] = coll_id;
            collected_xy[tid] = points_xy_image[coll_id];
            collected_conic_opacity[tid] = conic_opacity[coll_id];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BLOCK_SIZE, toDo); j++) {
            contributor++;
            const float2 xy = collected_xy[j];
            const float dx = __fsub_rn(xy.x, pixf.x), dy = __fsub_rn(xy.y, pixf.y);
            const float4 con_o = collected_conic_opacity[j];
            const float power = lgm::pair_power(con_o.x, con_o.y, con_o.z, dx, dy);
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, __fmul_rn(con_o.w, expf(power)));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
            if (test_T < 0.0001f) {
                done = true;
                continue;
            }
            for (int ch = 0; ch < NUM_CHANNELS; ch++)
                C[ch] = __fmaf_rn(__fmul_rn(features[collected_id[j] * NUM_CHANNELS + ch], alpha), T, C[ch]);
            weight = __fmaf_rn(alpha, T, weight);
            D = __fmaf_rn(__fmul_rn(depths[collected_id[j]], alpha), T, D);
            T = test_T;
            last_contributor = contributor;
        }
    }
    if (inside) {
        n_contrib[pix_id] = last_contributor;
        for (int ch = 0; ch < NUM_CHANNELS; ch++) out_color[ch * H * W + pix_id] = __fmaf_rn(T, bg_color[ch], C[ch]);
        out_alpha[pix_id] = weight;
        out_depth[pix_id] = D;
    }
}

// ---- A.5: the tile walked back to front, ten atomics per contributing pair ----
__global__ void __launch_bounds__(BLOCK_SIZE)
renderBackwardCUDA(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                   const float* __restrict__ bg_color, const float2* __restrict__ points_xy_image,
                   const float4* __restrict__ conic_opacity, const float* __restrict__ colors, const float* __restrict__ depths,
                   const float* __restrict__ alphas, const uint32_t* __restrict__ n_contrib,
                   const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dpixel_depths,
                   const float* __restrict__ dL_dalphas, float3* __restrict__ dL_dmean2D, float4* __restrict__ dL_dconic2D,
                   float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolors, float* __restrict__ dL_ddepths)
{
    const uint32_t horizontal_blocks = (W + BLOCK_X - 1) / BLOCK_X;
    const uint2 pix_min = {blockIdx.x * BLOCK_X, blockIdx.y * BLOCK_Y};
    const uint2 pix = {pix_min.x + threadIdx.x, pix_min.y + threadIdx.y};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    const uint2 range = ranges[blockIdx.y * horizontal_blocks + blockIdx.x];
    const int rounds = ((range.y - range.x + BLOCK_SIZE - 1) / BLOCK_SIZE);
    bool done = !inside;
    int toDo = range.y - range.x;

    __shared__ int collected_id[BLOCK_SIZE];
    __shared__ float2 collected_xy[BLOCK_SIZE];
    __shared__ float4 collected_conic_opacity[BLOCK_SIZE];
    __shared__ float collected_colors[NUM_CHANNELS * BLOCK_SIZE];
    __shared__ float collected_depths[BLOCK_SIZE];

    const float T_final = inside ? (1.0f - alphas[pix_id]) : 0.0f;
    float T = T_final;
    uint32_t contributor = toDo;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

    float accum_rec[NUM_CHANNELS] = {0};
    float dL_dpixel[NUM_CHANNELS] = {0};
    float accum_depth_rec = 0, dL_dpixel_depth = 0, accum_alpha_rec = 0, dL_dalpha_pix = 0;
    if (inside) {
        for (int i = 0; i < NUM_CHANNELS; i++) dL_dpixel[i] = dL_dpixels[i * H * W + pix_id];
        dL_dpixel_depth = dL_dpixel_depths[pix_id];
        dL_dalpha_pix = dL_dalphas[pix_id];
    }
    float last_alpha = 0, last_depth = 0;
    float last_color[NUM_CHANNELS] = {0};
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const int tid = threadIdx.y * BLOCK_X + threadIdx.x;

    for (int i = 0; i < rounds; i++, toDo -= BLOCK_SIZE) {
        __syncthreads();
        const int progress = i * BLOCK_SIZE + tid;
        if (range.x + progress < range.y) {
            const int coll_id = point_list[range.y - progress - 1];
            collected_id[tid] = coll_id;
            collected_xy[tid] = points_xy_image[coll_id];
            collected_conic_opacity[tid] = conic_opacity[coll_id];
            for (int ch = 0; ch < NUM_CHANNELS; ch++) collected_colors[ch * BLOCK_SIZE + tid] = colors[coll_id * NUM_CHANNELS + ch];
            collected_depths[tid] = depths[coll_id];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BLOCK_SIZE, toDo); j++) {
            contributor--;
            if (contributor >= (uint32_t)last_contributor) continue;
            const float2 xy = collected_xy[j];
            const float2 d = {__fsub_rn(xy.x, pixf.x), __fsub_rn(xy.y, pixf.y)};
            const float4 con_o = collected_conic_opacity[j];
            const float power = lgm::pair_power(con_o.x, con_o.y, con_o.z, d.x, d.y);
            if (power > 0.0f) continue;
            const float G = expf(power);
            const float alpha = fminf(0.99f, __fmul_rn(con_o.w, G));
            if (alpha < 1.0f / 255.0f) continue;

            T = T / (1.f - alpha);
            const float dchannel_dcolor = alpha * T;
            float dL_dalpha = 0.0f;
            const int global_id = collected_id[j];
            for (int ch = 0; ch < NUM_CHANNELS; ch++) {
                const float c = collected_colors[ch * BLOCK_SIZE + j];
                accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                last_color[ch] = c;
                const float dL_dchannel = dL_dpixel[ch];
                dL_dalpha += (c - accum_rec[ch]) * dL_dchannel;
                atomicAdd(&dL_dcolors[global_id * NUM_CHANNELS + ch], dchannel_dcolor * dL_dchannel);
            }
            const float c_d = collected_depths[j];
            accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
            last_depth = c_d;
            dL_dalpha += (c_d - accum_depth_rec) * dL_dpixel_depth;
            atomicAdd(&dL_ddepths[global_id], dchannel_dcolor * dL_dpixel_depth);

            accum_alpha_rec = last_alpha * 1.0f + (1.f - last_alpha) * accum_alpha_rec;
            dL_dalpha += (1.0f - accum_alpha_rec) * dL_dalpha_pix;

            dL_dalpha *= T;
            last_alpha = alpha;

            float bg_dot_dpixel = 0;
            for (int ch = 0; ch < NUM_CHANNELS; ch++) bg_dot_dpixel += bg_color[ch] * dL_dpixel[ch];
            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

            const float dL_dG = con_o.w * dL_dalpha;
            const float gdx = G * d.x, gdy = G * d.y;
            const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
            const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;

            atomicAdd(&dL_dmean2D[global_id].x, dL_dG * dG_ddelx * ddelx_dx);
            atomicAdd(&dL_dmean2D[global_id].y, dL_dG * dG_ddely * ddely_dy);
            atomicAdd(&dL_dconic2D[global_id].x, -0.5f * gdx * d.x * dL_dG);
            atomicAdd(&dL_dconic2D[global_id].y, -0.5f * gdx * d.y * dL_dG);
            atomicAdd(&dL_dconic2D[global_id].w, -0.5f * gdy * d.y * dL_dG);
            atomicAdd(&dL_dopacity[global_id], G * dL_dalpha);
        }
    }
}

// ---- A.6 (i)-(iii), (v): one thread per Gaussian (upstream: computeCov2DCUDA then preprocessCUDA, two launches) ----
__global__ void computeCov2DBackwardCUDA(int P, const float* __restrict__ means3D, const int* __restrict__ radii,
                                         const float* __restrict__ scales, const float* __restrict__ rotations, float scale_modifier,
                                         const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix, float focal_x,
                                         float focal_y, float tan_fovx, float tan_fovy, const float3* __restrict__ dL_dmean2D,
                                         const float4* __restrict__ dL_dconic, const float* __restrict__ dL_ddepth,
                                         float* __restrict__ dL_dmeans, float* __restrict__ dL_dscale, float* __restrict__ dL_drot)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P || !(radii[idx] > 0)) return;
    float mv[16], mp[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mv[k] = viewmatrix[k]; mp[k] = projmatrix[k]; }
    float dpos[3] = {0.f, 0.f, 0.f}, dscale[3] = {0.f, 0.f, 0.f}, drot[4] = {0.f, 0.f, 0.f, 0.f};
    const float4 gc = dL_dconic[idx];
    lgm::preprocess_point_bwd(means3D + 3 * idx, scales + 3 * idx, rotations + 4 * idx, scale_modifier, mv, mp, tan_fovx, tan_fovy,
                              focal_x, focal_y, dL_dmean2D[idx].x, dL_dmean2D[idx].y, gc.x, gc.y, gc.w, dL_ddepth[idx], dpos, dscale,
                              drot);
#pragma unroll
    for (int k = 0; k < 3; k++) { dL_dmeans[3 * idx + k] = dpos[k]; dL_dscale[3 * idx + k] = dscale[k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) dL_drot[4 * idx + k] = drot[k];
}

thread_local char g_err[256] = "";
int fail(cudaError_t e, const char* where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
#define REF_CUDA(call, where)                          \
    do {                                               \
        cudaError_t e__ = (call);                      \
        if (e__ != cudaSuccess) return -fail(e__, where) - 1000000; \
    } while (0)

}  // namespace refr

extern "C" {

const char* ref_last_error(void) { return refr::g_err; }

// Forward of ONE view.  Returns num_rendered (>= 0) or a negative error.  The three arenas are obtained through the
// caller's resize callbacks (the package's geometryBuffer / binningBuffer / imageBuffer lambdas).
int ref_forward(refr::resize_fn geometryBuffer, void* geom_ctx, refr::resize_fn binningBuffer, void* bin_ctx,
                refr::resize_fn imageBuffer, void* img_ctx, int P, const float* background, int width, int height,
                const float* means3D, const float* colors_precomp, const float* opacities, const float* scales,
                float scale_modifier, const float* rotations, const float* viewmatrix, const float* projmatrix, float tan_fovx,
                float tan_fovy, float* out_color, float* out_depth, float* out_alpha, int* radii)
{
    using namespace refr;
    const float focal_y = height / (2.0f * tan_fovy);
    const float focal_x = width / (2.0f * tan_fovx);
    char* chunkptr = geometryBuffer(geom_ctx, required<GeometryState>(P));
    GeometryState geomState = GeometryState::fromChunk(chunkptr, P);
    const dim3 tile_grid((width + BLOCK_X - 1) / BLOCK_X, (height + BLOCK_Y - 1) / BLOCK_Y, 1);
    const dim3 block(BLOCK_X, BLOCK_Y, 1);
    char* img_chunkptr = imageBuffer(img_ctx, required<ImageState>((size_t)width * height));
    ImageState imgState = ImageState::fromChunk(img_chunkptr, (size_t)width * height);
    if (P == 0) return 0;

    preprocessCUDA<<<(P + 255) / 256, 256>>>(P, means3D, scales, scale_modifier, rotations, opacities, colors_precomp, viewmatrix,
                                             projmatrix, width, height, tan_fovx, tan_fovy, focal_x, focal_y, radii,
                                             geomState.means2D, geomState.depths, geomState.cov3D, geomState.rgb,
                                             geomState.conic_opacity, tile_grid, geomState.tiles_touched);
    REF_CUDA(cudaGetLastError(), "preprocess");
    REF_CUDA(cub::DeviceScan::InclusiveSum(geomState.scanning_space, geomState.scan_size, geomState.tiles_touched,
                                           geomState.point_offsets, P), "InclusiveSum");
    int num_rendered = 0;
    REF_CUDA(cudaMemcpy(&num_rendered, geomState.point_offsets + P - 1, sizeof(int), cudaMemcpyDeviceToHost), "num_rendered D2H");

    char* binning_chunkptr = binningBuffer(bin_ctx, required<BinningState>(num_rendered));
    BinningState binningState = BinningState::fromChunk(binning_chunkptr, num_rendered);
    duplicateWithKeys<<<(P + 255) / 256, 256>>>(P, geomState.means2D, geomState.depths, geomState.point_offsets,
                                                binningState.point_list_keys_unsorted, binningState.point_list_unsorted, radii,
                                                tile_grid);
    REF_CUDA(cudaGetLastError(), "duplicateWithKeys");
    const int bit = (int)getHigherMsb(tile_grid.x * tile_grid.y);
    REF_CUDA(cub::DeviceRadixSort::SortPairs(binningState.list_sorting_space, binningState.sorting_size,
                                             binningState.point_list_keys_unsorted, binningState.point_list_keys,
                                             binningState.point_list_unsorted, binningState.point_list, num_rendered, 0, 32 + bit),
             "SortPairs");
    REF_CUDA(cudaMemset(imgState.ranges, 0, tile_grid.x * tile_grid.y * sizeof(uint2)), "memset ranges");
    if (num_rendered > 0) {
        identifyTileRanges<<<(num_rendered + 255) / 256, 256>>>(num_rendered, binningState.point_list_keys, imgState.ranges);
        REF_CUDA(cudaGetLastError(), "identifyTileRanges");
    }
    renderCUDA<<<tile_grid, block>>>(imgState.ranges, binningState.point_list, width, height, geomState.means2D, geomState.rgb,
                                     geomState.depths, geomState.conic_opacity, out_alpha, imgState.n_contrib, background, out_color,
                                     out_depth);
    REF_CUDA(cudaGetLastError(), "render");
    return num_rendered;
}

// Backward of ONE view from the arenas the forward left.  All dL_d* outputs must be ZERO on entry (the caller
// allocates them zero-filled per view, as RasterizeGaussiansBackwardCUDA does).
int ref_backward(int P, int R, const float* background, int width, int height, const float* means3D, const float* colors_precomp,
                 const float* alphas, const float* scales, float scale_modifier, const float* rotations, const float* viewmatrix,
                 const float* projmatrix, float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
                 char* img_buffer, const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D,
                 float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D, float* dL_dscale,
                 float* dL_drot)
{
    using namespace refr;
    if (P == 0) return 0;
    GeometryState geomState = GeometryState::fromChunk(geom_buffer, P);
    BinningState binningState = BinningState::fromChunk(binning_buffer, R);
    ImageState imgState = ImageState::fromChunk(img_buffer, (size_t)width * height);
    const float focal_y = height / (2.0f * tan_fovy);
    const float focal_x = width / (2.0f * tan_fovx);
    const dim3 tile_grid((width + BLOCK_X - 1) / BLOCK_X, (height + BLOCK_Y - 1) / BLOCK_Y, 1);
    const dim3 block(BLOCK_X, BLOCK_Y, 1);
    renderBackwardCUDA<<<tile_grid, block>>>(imgState.ranges, binningState.point_list, width, height, background, geomState.means2D,
                                             geomState.conic_opacity, colors_precomp, geomState.depths, alphas, imgState.n_contrib,
                                             dL_dpix, dL_dpix_depth, dL_dalphas, reinterpret_cast<float3*>(dL_dmean2D),
                                             reinterpret_cast<float4*>(dL_dconic), dL_dopacity, dL_dcolor, dL_ddepth);
    REF_CUDA(cudaGetLastError(), "render backward");
    computeCov2DBackwardCUDA<<<(P + 255) / 256, 256>>>(P, means3D, radii, scales, rotations, scale_modifier, viewmatrix, projmatrix,
                                                       focal_x, focal_y, tan_fovx, tan_fovy,
                                                       reinterpret_cast<const float3*>(dL_dmean2D),
                                                       reinterpret_cast<const float4*>(dL_dconic), dL_ddepth, dL_dmean3D, dL_dscale,
                                                       dL_drot);
    REF_CUDA(cudaGetLastError(), "preprocess backward");
    return 0;
}

// Byte offsets of the fields the parity tests read out of the arenas (A.7), for P Gaussians / L instances / N pixels.
void ref_arena_offsets(int P, int L, int n_pixels, size_t* out /* [6]: keys, vals, ranges, n_contrib, means2D, conic */)
{
    using namespace refr;
    char* base = nullptr;
    char* c = base;
    BinningState b = BinningState::fromChunk(c, L);
    out[0] = reinterpret_cast<size_t>(b.point_list_keys);
    out[1] = reinterpret_cast<size_t>(b.point_list);
    c = base;
    ImageState im = ImageState::fromChunk(c, n_pixels);
    out[2] = reinterpret_cast<size_t>(im.ranges);
    out[3] = reinterpret_cast<size_t>(im.n_contrib);
    c = base;
    GeometryState g = GeometryState::fromChunk(c, P);
    out[4] = reinterpret_cast<size_t>(g.means2D);
    out[5] = reinterpret_cast<size_t>(g.conic_opacity);
}

}  // extern "C"
