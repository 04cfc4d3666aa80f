// binning.cu — instance offsets (scan), K2 emit of (view|tile|depth) keys, K4 per-tile range identification.
// Replaces cub::DeviceScan::InclusiveSum + the blocking 4-byte D2H, duplicateWithKeys and identifyTileRanges of the
// external rasterizer, for all views of a step at once.
//
// Key layout (64 bit):  [63:32] global tile id = view * n_tiles + tile_y * grid_x + tile_x,  [31:0] float bits of the
// view-space depth.  Per view, key - (view * n_tiles << 32) is exactly the upstream key (tile << 32 | depth bits).
// Value (32 bit): view * P + Gaussian index (the row of the per-(view,Gaussian) geometry arrays).
#include "common.cuh"
#include "splat_math.cuh"

namespace lgm {

// Exclusive scan of the per-(view, Gaussian-block) tile counts, n_views x per_view entries (e.g. 208 x 384).  One CTA
// per view: it first sums the entries of all earlier views (coalesced, L2-resident: 320 KB at most) to get its base,
// then scans its own entries in chunks of 256.  The grand total (= number of instances L) is left on the device by the
// last view's CTA and read back ONCE per step by the host wrapper.  (A single 1024-thread CTA over all entries took
// 73 us on the 208-view step; this takes a third.)
__global__ void __launch_bounds__(kBlock)
scan_block_sums_kernel(const uint32_t* __restrict__ in, uint32_t per_view, uint32_t* __restrict__ out,
                       unsigned long long* __restrict__ total)
{
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_sum[8];
    const uint32_t view = blockIdx.x, t = threadIdx.x;
    const size_t before = (size_t)view * per_view;
    unsigned long long p0 = 0, p1 = 0, p2 = 0, p3 = 0;  // four independent streams of loads; 64 bit: the total is exact
    size_t i = t;
    for (; i + 3 * kBlock < before; i += 4 * kBlock) {
        p0 += in[i];
        p1 += in[i + kBlock];
        p2 += in[i + 2 * kBlock];
        p3 += in[i + 3 * kBlock];
    }
    for (; i < before; i += kBlock) p0 += in[i];
    unsigned long long part = p0 + p1 + p2 + p3;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((t & 31) == 0) s_sum[t >> 5] = part;
    __syncthreads();
    unsigned long long carry = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) carry += s_sum[w];
    for (uint32_t i0 = 0; i0 < per_view; i0 += kBlock) {
        const uint32_t j = i0 + t;
        const uint32_t v = j < per_view ? in[before + j] : 0u;
        uint32_t tot;
        const uint32_t excl = block_excl_scan_256(v, s_warp, &tot);
        if (j < per_view) out[before + j] = (uint32_t)(carry + excl);  // offsets fit 32 bit: the host bounds L below 2^30
        carry += tot;
    }
    if (view == gridDim.x - 1 && t == 0) total[0] = carry;
}

cudaError_t launch_scan_block_sums(cudaStream_t stream, const uint32_t* block_sums, uint32_t n_views, uint32_t per_view,
                                   uint32_t* block_offsets, unsigned long long* total)
{
    if (n_views == 0 || per_view == 0) return cudaMemsetAsync(total, 0, sizeof(unsigned long long), stream);
    scan_block_sums_kernel<<<n_views, kBlock, 0, stream>>>(block_sums, per_view, block_offsets, total);
    return cudaGetLastError();
}

// K2.  One thread per (view, Gaussian); its first output slot = block offset + in-block exclusive scan of the tile
// counts (recomputed from xy / radius with the same pinned tile_rect as preprocess), then rows of its rect in
// row-major order — the emit order the stable sort's tie-break relies on (ascending Gaussian index per tile).
constexpr uint32_t kCoopArea = 12;  // tiles per Gaussian above which the warp emits cooperatively

__global__ void __launch_bounds__(kBlock)
emit_kernel(const RenderParams prm, const int32_t* __restrict__ radii, const float2* __restrict__ xy,
            const float* __restrict__ depth, const uint32_t* __restrict__ block_offsets, uint64_t* __restrict__ keys,
            uint32_t* __restrict__ vals)
{
    __shared__ uint32_t s_warp[8];
    const int view = blockIdx.y;
    const int idx = blockIdx.x * kBlock + threadIdx.x;
    const size_t gi = (size_t)view * prm.P + idx;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    uint32_t area = 0;
    if (idx < prm.P) {
        const int r = radii[gi];
        if (r > 0) {
            const float2 p = xy[gi];
            tile_rect(p.x, p.y, r, prm.gx, prm.gy, x0, y0, x1, y1);
            area = (uint32_t)((x1 - x0) * (y1 - y0));
        }
    }
    const uint32_t excl = block_excl_scan_256(area, s_warp, nullptr);
    const uint32_t off32 = block_offsets[(size_t)view * gridDim.x + blockIdx.x] + excl;  // < 2^30 (api.cu bounds L)
    const uint32_t dbits = area ? __float_as_uint(depth[gi]) : 0u;
    const uint32_t tile_base = (uint32_t)view * (uint32_t)prm.n_tiles;
    const uint32_t val = (uint32_t)gi;
    const bool big = area > kCoopArea;

    // small footprints: the owning thread writes its few instances (all lanes busy, runs of adjacent threads adjoin)
    if (area != 0 && !big) {
        size_t off = off32;
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                keys[off] = ((uint64_t)(tile_base + (uint32_t)(y * prm.gx + x)) << 32) | dbits;
                vals[off] = val;
                off++;
            }
    }
    // large footprints (early training, 1024^2 views: hundreds of tiles per Gaussian): the whole warp writes one
    // Gaussian's instances, 32 consecutive slots per step — coalesced, no divergent serial loops
    const int lane = threadIdx.x & 31;
    unsigned m = __ballot_sync(0xffffffffu, big);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t a = __shfl_sync(0xffffffffu, area, src);
        const uint32_t o = __shfl_sync(0xffffffffu, off32, src);
        const uint32_t sx0 = __shfl_sync(0xffffffffu, (uint32_t)x0, src), sy0 = __shfl_sync(0xffffffffu, (uint32_t)y0, src);
        const uint32_t w = __shfl_sync(0xffffffffu, (uint32_t)(x1 - x0), src);
        const uint32_t sd = __shfl_sync(0xffffffffu, dbits, src), sv = __shfl_sync(0xffffffffu, val, src);
        for (uint32_t i = lane; i < a; i += 32) {  // i-th tile of the rect in row-major order: the emit order
            const uint32_t ry = i / w, rx = i - ry * w;
            keys[(size_t)o + i] = ((uint64_t)(tile_base + (sy0 + ry) * (uint32_t)prm.gx + sx0 + rx) << 32) | sd;
            vals[(size_t)o + i] = sv;
        }
    }
}

cudaError_t launch_emit(cudaStream_t stream, const RenderParams& prm, const int32_t* radii, const float2* xy,
                        const float* depth, const uint32_t* block_offsets, uint64_t* keys, uint32_t* vals)
{
    if (prm.P == 0 || prm.n_views == 0) return cudaSuccess;
    dim3 grid((prm.P + kBlock - 1) / kBlock, prm.n_views);
    emit_kernel<<<grid, kBlock, 0, stream>>>(prm, radii, xy, depth, block_offsets, keys, vals);
    return cudaGetLastError();
}

// K4.  ranges[global tile] = [start, end) into the sorted instance list; ranges must be zero-filled first
// (empty tiles stay (0,0)), as upstream's cudaMemset + identifyTileRanges.
__global__ void __launch_bounds__(kBlock)
tile_ranges_kernel(const uint64_t* __restrict__ keys, uint32_t L, uint2* __restrict__ ranges)
{
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= L) return;
    const uint32_t cur = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[cur].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
        if (cur != prev) {
            ranges[prev].y = i;
            ranges[cur].x = i;
        }
    }
    if (i == L - 1) ranges[cur].y = L;
}

cudaError_t launch_tile_ranges(cudaStream_t stream, const uint64_t* keys, uint32_t L, uint2* ranges)
{
    if (L == 0) return cudaSuccess;
    tile_ranges_kernel<<<(L + kBlock - 1) / kBlock, kBlock, 0, stream>>>(keys, L, ranges);
    return cudaGetLastError();
}

}  // namespace lgm
