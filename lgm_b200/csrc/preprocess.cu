// preprocess.cu — K1 (3D->2D projection, covariance, EWA conic, radius, tile rect), markVisible, and K7 (preprocess
// backward chain rule to means / scales / rotations).  Replaces preprocessCUDA fwd/bwd + computeCov2DCUDA of the
// external rasterizer, batched over all views of a step in one launch.
//
// HBM roofline: K1 writes 36 B per (view, Gaussian) (depth 4, radius 4, xy 8, conic+opacity 16, tile-count
// block sums amortised) and reads the 56-B AoS Gaussian row once per view through L2 (rows of a scene are shared by
// all its views); K7 reads 4 + 48 B per (view, Gaussian) and writes 56 B per Gaussian.
#include "common.cuh"
#include "splat_math.cuh"

namespace lgm {

// Stage n (<=256) consecutive 14-float Gaussian rows into shared memory with 16-byte vector loads when aligned.
__device__ __forceinline__ void stage_rows(float* s_g, const float* __restrict__ src, int n)
{
    const int nfl = n * 14;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const int nv = nfl >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(s_g);
        for (int i = threadIdx.x; i < nv; i += kBlock) d4[i] = __ldg(s4 + i);
        for (int i = (nv << 2) + threadIdx.x; i < nfl; i += kBlock) s_g[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < nfl; i += kBlock) s_g[i] = __ldg(src + i);
    }
}

__global__ void __launch_bounds__(kBlock)
preprocess_fwd_kernel(const RenderParams prm, const float* __restrict__ gaussians, const float* __restrict__ view_mats,
                      const float* __restrict__ proj_mats, const int32_t* __restrict__ view_scene,
                      float* __restrict__ depth, int32_t* __restrict__ radii, float2* __restrict__ xy,
                      float4* __restrict__ conic_opacity, uint32_t* __restrict__ tiles_touched,
                      uint32_t* __restrict__ block_sums, const float* __restrict__ cov3d, float4* __restrict__ zero_rows)
{
    __shared__ __align__(16) float s_g[kBlock * 14];
    __shared__ float s_mv[16], s_mp[16];
    __shared__ uint32_t s_warp[8];
    const int view = blockIdx.y;
    const int scene = view_scene[view];
    const int base = blockIdx.x * kBlock;
    const int n = min(kBlock, prm.P - base);
    stage_rows(s_g, gaussians + ((size_t)scene * prm.P + base) * 14, n);
    if (threadIdx.x < 16) s_mv[threadIdx.x] = view_mats[view * 16 + threadIdx.x];
    else if (threadIdx.x < 32) s_mp[threadIdx.x - 16] = proj_mats[view * 16 + threadIdx.x - 16];
    __syncthreads();

    uint32_t tiles = 0;
    if ((int)threadIdx.x < n) {
        const float* g = s_g + threadIdx.x * 14;
        // cov3d (upstream's cov3D_precomp, [n_scenes, P, 6]) replaces the covariance built from scale / rotation
        const Geom o = preprocess_point(g, g + 4, g + 7, prm.mod, s_mv, s_mp, prm.W, prm.H, prm.tanx, prm.tany, prm.fx,
                                        prm.fy, prm.gx, prm.gy,
                                        cov3d ? cov3d + ((size_t)scene * prm.P + base + threadIdx.x) * 6 : nullptr);
        const size_t gi = (size_t)view * prm.P + base + threadIdx.x;
        depth[gi] = o.depth;
        radii[gi] = o.radius;
        xy[gi] = make_float2(o.px, o.py);
        conic_opacity[gi] = make_float4(o.cx, o.cy, o.cz, o.radius > 0 ? g[3] : 0.f);
        if (tiles_touched) tiles_touched[gi] = o.tiles;
        tiles = o.tiles;
        if (zero_rows) {
            // the backward's gradient row of this (view, Gaussian) pair, zeroed here: the kernel is bound by instruction
            // issue (correctly rounded divisions), so the 48 B of stores ride along instead of a separate 1 GB fill
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            zero_rows[3 * gi] = z;
            zero_rows[3 * gi + 1] = z;
            zero_rows[3 * gi + 2] = z;
        }
    }
    // block sum of tiles_touched -> one partial per (view, Gaussian block); scanned by scan_block_sums_kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t s = tiles;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_warp[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_warp[w];
        block_sums[(size_t)view * gridDim.x + blockIdx.x] = t;
    }
}

cudaError_t launch_preprocess_fwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                  const float* view_mats, const float* proj_mats, const int32_t* view_scene,
                                  float* depth, int32_t* radii, float2* xy, float4* conic_opacity,
                                  uint32_t* tiles_touched, uint32_t* block_sums, const float* cov3d, float* zero_rows)
{
    static_assert(kGradRow == 12, "three 16-byte vectors per gradient row");
    if (prm.P == 0 || prm.n_views == 0) return cudaSuccess;
    dim3 grid((prm.P + kBlock - 1) / kBlock, prm.n_views);
    preprocess_fwd_kernel<<<grid, kBlock, 0, stream>>>(prm, gaussians, view_mats, proj_mats, view_scene, depth, radii, xy,
                                                       conic_opacity, tiles_touched, block_sums, cov3d,
                                                       reinterpret_cast<float4*>(zero_rows));
    return cudaGetLastError();
}

// ---- markVisible (upstream checkFrustum; not called by LGM, kept for API completeness) ----
__global__ void __launch_bounds__(kBlock)
mark_visible_kernel(int P, const float* __restrict__ means, const float* __restrict__ mv, uint8_t* __restrict__ visible)
{
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= P) return;
    float m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = __ldg(mv + k);
    const float z = affine_row(m, 2, means[3 * i], means[3 * i + 1], means[3 * i + 2]);
    visible[i] = !(z <= kNearCull);
}

cudaError_t launch_mark_visible(cudaStream_t stream, int P, const float* means, const float* view_mat, uint8_t* visible)
{
    if (P == 0) return cudaSuccess;
    mark_visible_kernel<<<(P + kBlock - 1) / kBlock, kBlock, 0, stream>>>(P, means, view_mat, visible);
    return cudaGetLastError();
}

// ---- K7: preprocess backward.  One thread per Gaussian loops over the views of its scene and accumulates in
// registers, so the sum over views needs no atomics and is deterministic. ----
constexpr int kViewChunk = 32;
__global__ void __launch_bounds__(kBlock, 3)
preprocess_bwd_kernel(const RenderParams prm, const float* __restrict__ gaussians, const float* __restrict__ view_mats,
                      const float* __restrict__ proj_mats, const int32_t* __restrict__ scene_view_offsets,
                      const int32_t* __restrict__ radii, const float4* __restrict__ conic_opacity,
                      const float* __restrict__ grad_rows, float* __restrict__ dL_dgaussians, int accumulate,
                      const float* __restrict__ cov3d, float* __restrict__ dL_dcov3d)
{
    __shared__ __align__(16) float s_g[kBlock * 14];
    __shared__ float s_m[kViewChunk * 32];  // view and projection matrices of up to kViewChunk views
    const int scene = blockIdx.y;
    const int base = blockIdx.x * kBlock;
    const int n = min(kBlock, prm.P - base);
    const float* src = gaussians + ((size_t)scene * prm.P + base) * 14;
    stage_rows(s_g, src, n);
    __syncthreads();
    // Carried through the view loop: position, opacity, the view-independent 3D covariance; accumulated: dL/dpos,
    // dL/dopacity, dL/drgb and g6 = dL/dcov3D (mapped to scale / rotation once, after the loop — linear in g6).
    float pos[3] = {0.f, 0.f, 0.f}, opacity = 0.f, cov6[6];
    float d[14], g6[6];
#pragma unroll
    for (int k = 0; k < 14; k++) d[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) g6[k] = cov6[k] = 0.f;
    const bool active = (int)threadIdx.x < n;
    if (active) {
        const float* g = s_g + threadIdx.x * 14;
        float M[9];
        pos[0] = g[0]; pos[1] = g[1]; pos[2] = g[2];
        opacity = g[3];
        if (cov3d) {
            const float* c = cov3d + ((size_t)scene * prm.P + base + threadIdx.x) * 6;
#pragma unroll
            for (int k = 0; k < 6; k++) cov6[k] = c[k];
        } else {
            cov3d_from_scale_rot(g[4], g[5], g[6], prm.mod, g[7], g[8], g[9], g[10], cov6, M);
        }
    }
    const int v0 = scene_view_offsets[scene], v1 = scene_view_offsets[scene + 1];
    // The matrices of kViewChunk views are staged at once, so that the view loop itself runs without block barriers: the
    // warps drift apart and one warp's row loads overlap the others' arithmetic.  The radius of the next view is fetched
    // one iteration ahead (the row loads depend on it).
    for (int c0 = v0; c0 < v1; c0 += kViewChunk) {
        const int nc = min(kViewChunk, v1 - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * 32; i += kBlock) {
            const int v = c0 + (i >> 5), k = i & 31;
            s_m[i] = k < 16 ? view_mats[v * 16 + k] : proj_mats[v * 16 + k - 16];
        }
        __syncthreads();
        if (!active) continue;
        const size_t g0 = (size_t)c0 * prm.P + base + threadIdx.x;
        int rad = radii[g0];
        for (int j = 0; j < nc; j++) {
            const size_t gi = g0 + (size_t)j * prm.P;
            const int rad_next = j + 1 < nc ? radii[gi + prm.P] : 0;
            if (rad > 0) {
                const float4* row = reinterpret_cast<const float4*>(grad_rows + gi * kGradRow);
                const float4 r0 = __ldg(row), r1 = __ldg(row + 1), r2 = __ldg(row + 2);
                // r0 = (Sx, Sy, Sxx, Sxy)  r1 = (Syy, S0 = dL/dopacity, col.r, col.g)  r2 = (col.b, depth, -, -): moment form
                const float* m = s_m + j * 32;
                preprocess_point_bwd_view(pos, cov6, m, m + 16, prm.tanx, prm.tany, prm.fx, prm.fy, r0.x, r0.y, r0.z, r0.w, r1.x,
                                          r2.y, d, g6, /*moments=*/true, (float)prm.W, (float)prm.H, opacity);
                d[3] += r1.y;
                d[11] += r1.z;
                d[12] += r1.w;
                d[13] += r2.x;
            }
            rad = rad_next;
        }
    }
    if (active) {
        if (cov3d) {
            // cov3D_precomp: the gradient stops at the covariance (upstream's dL_dcov3D); scale / rotation get none
            float* o = dL_dcov3d + ((size_t)scene * prm.P + base + threadIdx.x) * 6;
#pragma unroll
            for (int k = 0; k < 6; k++) o[k] = accumulate ? o[k] + g6[k] : g6[k];
        } else {
            preprocess_point_bwd_finish(s_g + threadIdx.x * 14 + 4, s_g + threadIdx.x * 14 + 7, prm.mod, g6, d + 4, d + 7);
        }
    }
    // transpose through shared memory for coalesced stores of the 14-float rows
    __syncthreads();
    if (active) {
#pragma unroll
        for (int k = 0; k < 14; k++) s_g[threadIdx.x * 14 + k] = d[k];
    }
    __syncthreads();
    float* dst = dL_dgaussians + ((size_t)scene * prm.P + base) * 14;
    const int nfl = n * 14;
    if (accumulate) {
        for (int i = threadIdx.x; i < nfl; i += kBlock) dst[i] += s_g[i];
    } else {
        for (int i = threadIdx.x; i < nfl; i += kBlock) dst[i] = s_g[i];
    }
}

cudaError_t launch_preprocess_bwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                  const float* view_mats, const float* proj_mats, const int32_t* scene_view_offsets,
                                  const int32_t* radii, const float4* conic_opacity, const float* grad_rows,
                                  float* dL_dgaussians, int accumulate, const float* cov3d, float* dL_dcov3d)
{
    if (prm.P == 0 || prm.n_scenes == 0) return cudaSuccess;
    dim3 grid((prm.P + kBlock - 1) / kBlock, prm.n_scenes);
    preprocess_bwd_kernel<<<grid, kBlock, 0, stream>>>(prm, gaussians, view_mats, proj_mats, scene_view_offsets, radii,
                                                       conic_opacity, grad_rows, dL_dgaussians, accumulate, cov3d, dL_dcov3d);
    return cudaGetLastError();
}

// Moment rows -> upstream's screen-space gradients (diagnostics, means2D.grad of the Level-1 API, parity tests):
// out row = (dL/dmean2D.xy, dL/dconic.xx .xy .yy, dL/dopacity, dL/dcolour.rgb, dL/ddepth, 0, 0)
__global__ void __launch_bounds__(kBlock)
screen_gradients_kernel(const RenderParams prm, size_t n_rows, const float4* __restrict__ conic_opacity,
                        const float* __restrict__ grad_rows, float* __restrict__ out)
{
    const size_t gi = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (gi >= n_rows) return;
    const float4* row = reinterpret_cast<const float4*>(grad_rows + gi * kGradRow);
    const float4 r0 = row[0], r1 = row[1], r2 = row[2];
    const float4 co = conic_opacity[gi];
    float g2x, g2y, gcx, gcy, gcz;
    moments_to_gradients((float)prm.W, (float)prm.H, co.x, co.y, co.z, co.w, r0.x, r0.y, r0.z, r0.w, r1.x, &g2x, &g2y, &gcx, &gcy,
                         &gcz);
    float4* dst = reinterpret_cast<float4*>(out + gi * kGradRow);
    dst[0] = make_float4(g2x, g2y, gcx, gcy);
    dst[1] = make_float4(gcz, r1.y, r1.z, r1.w);
    dst[2] = make_float4(r2.x, r2.y, 0.f, 0.f);
}

cudaError_t launch_screen_gradients(cudaStream_t stream, const RenderParams& prm, const float4* conic_opacity,
                                    const float* grad_rows, float* out)
{
    const size_t n_rows = (size_t)prm.n_views * prm.P;
    if (n_rows == 0) return cudaSuccess;
    screen_gradients_kernel<<<(unsigned)((n_rows + kBlock - 1) / kBlock), kBlock, 0, stream>>>(prm, n_rows, conic_opacity, grad_rows, out);
    return cudaGetLastError();
}

}  // namespace lgm
