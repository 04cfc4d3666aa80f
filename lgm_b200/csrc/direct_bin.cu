// direct_bin.cu — "direct" binning: the (view|tile|depth) order of the sorted instance list WITHOUT a global radix
// sort.  Same output, bit for bit, as emit + onesweep + tile ranges (binning.cu / radix_sort.cu), i.e. as upstream's
// duplicateWithKeys + cub::DeviceRadixSort::SortPairs + identifyTileRanges:
//
//   D1  count     per (view, Gaussian): +1 per touched tile, aggregated per CTA in shared memory, then one global
//                 atomic per (CTA, touched tile) into tile_counts[global tile]
//   D2  ranges    one CTA per view scans the view's tile counts on top of the totals of the views before it = every
//                 tile's [start, end) in the final list (empty tiles stay (0,0)); non-empty tiles appended to a work
//                 list, longest tile recorded
//   D3  scatter   same enumeration as D1; each CTA reserves a run of every touched tile's segment with one global atomic,
//                 its instances take the slots of the run (shared-memory atomics) and store (value, depth bits) there —
//                 8 B, in arbitrary order inside the segment
//   D4  tile sort one CTA per non-empty tile orders its segment in SHARED MEMORY by the 64-bit key
//                 (depth bits << 32 | value) and writes the values (and, on request, the upstream 64-bit keys)
//
// Why the order is exact: inside one tile the stable LSD sort leaves instances ordered by depth bits, ties in emit
// order; emit order inside one tile is ascending value (value = view * P + Gaussian index, each Gaussian touches a
// tile once).  So the final order of a tile is ascending (depth bits, value) — a total order on unique keys, which any
// correct sort reproduces no matter in which order the atomics of D3 filled the segment.
//
// D4 is a bucket sort: keys are mapped monotonically to ~n/2 buckets by (depth - min) * nb / (range + 1), counted and
// grouped with shared-memory atomics (two sweeps), and each element finds its final slot by counting the smaller keys
// in its own bucket (2 on average).  No ballots, no per-warp histograms: ~70 instructions per element against ~6 x 45
// for the LSD passes, and 20 B of HBM traffic per instance instead of 152.  A tile whose keys pile up in few buckets
// (depth ties) is ordered by a bitonic network instead of the rank loop, when the loop would make more compares than the
// network.  Two forms (lgm_set_tuning "sort_bulk"): tile_group_sort_kernel (default) reads the segment from global memory
// and groups the KEYS in shared memory, 8 B per instance; tile_bucket_sort_kernel stages the segment (load / store loop or
// one TMA bulk copy) and groups 16-bit indices, 10 B per instance.  Four size classes (S: <= 2,048 instances, 256 threads;
// M: <= 5,632, 512 threads; X: <= 11,776, 1024 threads, two CTAs per SM in the default form; L: <= 20,480, one 1024-thread
// CTA per SM).  A longer tile cannot be held in shared memory: the caller (api.cu) is told the longest tile and uses the
// onesweep path for such a step.
#include "common.cuh"
#include "splat_math.cuh"

namespace lgm {
namespace {

// Size classes of the per-tile sort (shared memory per CTA in the default / the staged form).
// S: tiles of up to 2,048 instances, 256 threads, 21 / 25 KB, 8 / 6 CTAs per SM.  A trained scene's tiles hold ~750
// instances on average: with 512 threads most of a CTA idles through the barriers, bucket scans and the two-iteration
// sweeps of such a tile (the fixed cost per tile is paid by every warp).
constexpr int kSortThreadsS = 256, kSortCapS = 2048, kLgBucketsS = 10;
// M: up to 5,632 instances, 512 threads, 54 / 64 KB, 4 / 3 CTAs per SM — where most instances of a trained scene live.
constexpr int kSortThreadsM = 512, kSortCapM = 5632, kLgBucketsM = 11;
// X: up to 11,776 instances, 1024 threads, 111 KB in the default form: TWO CTAs per SM, so that one CTA's global phases
// overlap the other's shared-memory phases (1024^2 views of 1 M Gaussians and untrained Gaussians have most of their
// instances in such tiles; 134 KB and one CTA per SM in the staged form).
constexpr int kSortThreadsX = 1024, kSortCapX = 11776, kLgBucketsX = 12;
// L: up to 20,480 instances, 1024 threads, 180 / 216 KB, one CTA per SM; launched first.
constexpr int kSortThreadsL = 1024, kSortCapL = 20480, kLgBucketsL = 12;
// (+ 16 B: the bulk copy of a segment starts at a 16-byte boundary, up to one pair before the segment, and ends at one)
constexpr size_t sort_smem(int cap, int lg_buckets) { return (size_t)cap * 10 + 16 + (size_t)((1 << lg_buckets) + 1) * 4 + 64 * 4; }

// ---- 1-D bulk copy (TMA, cp.async.bulk) + mbarrier: the segment of a tile is contiguous in `pairs` ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
constexpr uint32_t kCoopAreaD = 12;  // as binning.cu: larger footprints are enumerated by the whole warp

// Calls f(tile index inside the view, value, depth bits) once per (Gaussian, touched tile) for the Gaussian `idx` of
// `view` held by this thread; footprints above kCoopAreaD tiles are enumerated by the whole warp (lanes across the
// rect), so every lane of the warp must make this call together.
template <bool WITH_DEPTH, typename F>
__device__ __forceinline__ void for_each_touched_tile(const RenderParams& prm, const int32_t* __restrict__ radii,
                                                      const float2* __restrict__ xy, const float* __restrict__ depth,
                                                      int view, int idx, F&& f)
{
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
    uint32_t dbits = 0;
    if (WITH_DEPTH && area) dbits = __float_as_uint(depth[gi]);
    const uint32_t val = (uint32_t)gi;
    const bool big = area > kCoopAreaD;
    if (area != 0 && !big) {
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) f((uint32_t)(y * prm.gx + x), val, dbits);
    }
    const int lane = threadIdx.x & 31;
    unsigned m = __ballot_sync(0xffffffffu, big);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t a = __shfl_sync(0xffffffffu, area, src);
        const uint32_t sx0 = __shfl_sync(0xffffffffu, (uint32_t)x0, src), sy0 = __shfl_sync(0xffffffffu, (uint32_t)y0, src);
        const uint32_t w = __shfl_sync(0xffffffffu, (uint32_t)(x1 - x0), src);
        const uint32_t sd = __shfl_sync(0xffffffffu, dbits, src), sv = __shfl_sync(0xffffffffu, val, src);
        for (uint32_t i = lane; i < a; i += 32) {
            const uint32_t ry = i / w, rx = i - ry * w;
            f((sy0 + ry) * (uint32_t)prm.gx + sx0 + rx, sv, sd);
        }
    }
}

// The same enumeration for a rect already held in a register: x0 | y0 << 8 | x1 << 16 | y1 << 24 (tile grids up to 255 x 255;
// 0 = no footprint).  The count / scatter kernel loads the rects of its kEnumItems Gaussians up front — kEnumItems
// independent loads in flight instead of a dependent radius -> centre -> rect chain per item (43 % of the scatter's stall
// samples sat on those two loads) — and walks them twice without reloading.
template <typename F>
__device__ __forceinline__ void for_each_tile_of_rect(uint32_t packed, int gx, uint32_t val, uint32_t dbits, F&& f)
{
    const int x0 = packed & 0xffu, y0 = (packed >> 8) & 0xffu, x1 = (packed >> 16) & 0xffu, y1 = packed >> 24;
    const uint32_t area = (uint32_t)((x1 - x0) * (y1 - y0));
    const bool big = area > kCoopAreaD;
    if (area != 0 && !big) {
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) f((uint32_t)(y * gx + x), val, dbits);
    }
    const int lane = threadIdx.x & 31;
    unsigned m = __ballot_sync(0xffffffffu, big);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t pk = __shfl_sync(0xffffffffu, packed, src);
        const uint32_t sd = __shfl_sync(0xffffffffu, dbits, src), sv = __shfl_sync(0xffffffffu, val, src);
        const uint32_t sx0 = pk & 0xffu, sy0 = (pk >> 8) & 0xffu, w = ((pk >> 16) & 0xffu) - sx0;
        const uint32_t a = w * ((pk >> 24) - sy0);
        for (uint32_t i = lane; i < a; i += 32) {
            const uint32_t ry = i / w, rx = i - ry * w;
            f((sy0 + ry) * (uint32_t)gx + sx0 + rx, sv, sd);
        }
    }
}

// D1 / D3.  A CTA owns kEnumItems x 256 consecutive Gaussians of one view and aggregates in shared memory first: with
// ~170 non-empty tiles per view, every one of the 62 M instances of a step going to global memory on its own piles
// ~1,750 same-address atomics on each counter (measured: 2.8 ms count + 3.8 ms scatter, 4 % issue utilisation).
// Per-CTA histograms cut the global atomics to one per (CTA, touched tile).
//   SCATTER = false: counts[gtile] += this CTA's instances of the tile.
//   SCATTER = true : counts[] (holding the totals) is counted back down by the CTA's number, which reserves a
//                    contiguous run of the tile's segment; the CTA's instances take the slots of that run.
constexpr int kEnumItems = 8;
constexpr int kEnumMaxTiles = 6144;  // 2 x 4 B per tile of dynamic shared memory must stay under the default 48 KB

// ---- coarse grouping for steps with large footprints (many tiles per Gaussian) ----
// With ~16 tiles per Gaussian and Gaussians in arbitrary spatial order, the scatter's 8-byte stores of one CTA go to
// thousands of tile segments at once: measured on 32 views of 1 M Gaussians at 1024^2 (499 M instances) 9.3 ms, DRAM
// write 9.7 GB for 4 GB of pairs plus 5.5 GB of read-for-ownership — partial-sector traffic.  So such steps first group
// the (view, Gaussian) pairs by SUPER-TILE (8x8 tiles): one 16-byte entry per touched super-tile (value, depth bits, the
// tile rect clipped to the super-tile, view) in a counting sort whose runs are long, and the fine scatter then works
// through one super-tile's entries at a time: a CTA's stores go to the 64 segments of that super-tile only, in runs of
// hundreds of pairs.  The entry counts come for free with the tile counts (same enumeration).
constexpr int kSuperShift = 3, kSuperTiles = 1 << kSuperShift;
constexpr int kCoarseChunk = 2048;       // entries per work item of the fine scatter
// fine scatter: from this many instances per entry on, the placement is tile-major (cost ~ 64 tests per entry, coalesced
// stores, no atomics) instead of entry-major (cost ~ instances).  Measured: untrained Gaussians (~25 per entry) bin 12.8 ->
// 10.6 ms tile-major; 1 M Gaussians at 1024^2 (9 per entry) 7.2 -> 9.5 ms.
constexpr uint32_t kTileMajorRatio = 16;
constexpr uint32_t kCoopFine = 24;       // fine scatter: rects (clipped to a super-tile, <= 64 tiles) above this are walked by the whole warp
constexpr uint32_t kCoarseMaxSupers = 1024;  // per view (a 4096^2 image); beyond that the plain scatter is used

__host__ __device__ inline int supers_x(const RenderParams& prm) { return (prm.gx + kSuperTiles - 1) >> kSuperShift; }
__host__ __device__ inline int supers_y(const RenderParams& prm) { return (prm.gy + kSuperTiles - 1) >> kSuperShift; }

// the tile rect of (view, idx), area 0 when culled
__device__ __forceinline__ uint32_t load_rect(const RenderParams& prm, const int32_t* __restrict__ radii, const float2* __restrict__ xy,
                                              size_t gi, bool in_range, int& x0, int& y0, int& x1, int& y1)
{
    x0 = y0 = x1 = y1 = 0;
    if (!in_range) return 0;
    const int r = radii[gi];
    if (r <= 0) return 0;
    const float2 p = xy[gi];
    tile_rect(p.x, p.y, r, prm.gx, prm.gy, x0, y0, x1, y1);
    return (uint32_t)((x1 - x0) * (y1 - y0));
}

// The packed rects (x0 | y0 << 8 | x1 << 16 | y1 << 24, 0 = no footprint; tile grids up to 255 x 255) of the kEnumItems
// Gaussians first, first + 256, ... of `view`: all radii, then all centres — independent loads in flight.
__device__ __forceinline__ void load_rects_packed(const RenderParams& prm, const int32_t* __restrict__ radii,
                                                  const float2* __restrict__ xy, int view, int first, uint32_t (&rect)[kEnumItems])
{
    int rad[kEnumItems];
    float2 ctr[kEnumItems];
#pragma unroll
    for (int k = 0; k < kEnumItems; k++) {
        const int idx = first + k * kBlock;
        rad[k] = idx < prm.P ? radii[(size_t)view * prm.P + idx] : 0;
    }
#pragma unroll
    for (int k = 0; k < kEnumItems; k++)
        ctr[k] = rad[k] > 0 ? xy[(size_t)view * prm.P + first + k * kBlock] : make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kEnumItems; k++) {
        rect[k] = 0u;
        if (rad[k] > 0) {
            int x0, y0, x1, y1;
            tile_rect(ctr[k].x, ctr[k].y, rad[k], prm.gx, prm.gy, x0, y0, x1, y1);
            if ((x1 - x0) * (y1 - y0) != 0) rect[k] = (uint32_t)x0 | (uint32_t)y0 << 8 | (uint32_t)x1 << 16 | (uint32_t)y1 << 24;
        }
    }
}

template <bool SCATTER>
__global__ void __launch_bounds__(kBlock)
tile_enumerate_kernel(const RenderParams prm, const int32_t* __restrict__ radii, const float2* __restrict__ xy,
                      const float* __restrict__ depth, uint32_t* __restrict__ counts, uint32_t* __restrict__ view_totals,
                      const uint2* __restrict__ ranges, uint2* __restrict__ pairs, uint32_t* __restrict__ super_counts,
                      uint32_t* __restrict__ view_entries, const unsigned long long* __restrict__ total_instances,
                      unsigned long long coarse_gate)
{
    extern __shared__ uint32_t s_enum[];
    uint32_t* s_hist = s_enum;                 // [n_tiles] instances of this CTA per tile, then the fill cursor
    uint32_t* s_base = s_enum + prm.n_tiles;   // [n_tiles] first slot of this CTA's run (SCATTER); count: coarse entries per super-tile
    const int view = blockIdx.y;
    const int first = blockIdx.x * (kBlock * kEnumItems) + threadIdx.x;
    const uint32_t tile_base = (uint32_t)view * (uint32_t)prm.n_tiles;
    const int nsx = supers_x(prm), n_super = nsx * supers_y(prm);
    // the coarse entries are only counted for steps that can use them: large footprints, i.e. at least coarse_gate
    // instances per (view, Gaussian) pair — the instance total is on the device since the preprocess stage's scan
    const bool heavy = !SCATTER && super_counts &&
                       *total_instances >= coarse_gate * (unsigned long long)prm.n_views * (unsigned long long)prm.P;
    // rects of this thread's Gaussians, loaded up front and kept packed in registers for every pass below
    const bool small_grid = prm.gx <= 255 && prm.gy <= 255;
    uint32_t rect[kEnumItems];
    if (small_grid) load_rects_packed(prm, radii, xy, view, first, rect);
    auto rect_of = [&](int k, int& x0, int& y0, int& x1, int& y1) -> uint32_t {
        if (small_grid) {
            x0 = rect[k] & 0xffu; y0 = (rect[k] >> 8) & 0xffu; x1 = (rect[k] >> 16) & 0xffu; y1 = rect[k] >> 24;
            return rect[k];
        }
        const int idx = first + k * kBlock;
        return load_rect(prm, radii, xy, (size_t)view * prm.P + idx, idx < prm.P, x0, y0, x1, y1);
    };
    if (heavy) {
        // entries per super-tile of this CTA's Gaussians (a Gaussian has one entry per super-tile its rect touches)
        for (int i = threadIdx.x; i < n_super; i += kBlock) s_base[i] = 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kEnumItems; k++) {
            int x0, y0, x1, y1;
            if (rect_of(k, x0, y0, x1, y1) == 0) continue;
            for (int sy = y0 >> kSuperShift; sy <= (y1 - 1) >> kSuperShift; sy++)
                for (int sx = x0 >> kSuperShift; sx <= (x1 - 1) >> kSuperShift; sx++) atomicAdd(&s_base[sy * nsx + sx], 1u);
        }
        __syncthreads();
        uint32_t mine = 0;
        for (int i = threadIdx.x; i < n_super; i += kBlock) {
            const uint32_t c = s_base[i];
            if (c) atomicAdd(&super_counts[(size_t)view * n_super + i], c);
            mine += c;
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&view_entries[view], mine);
    }
    if (heavy && (prm.gx + 1) * (prm.gy + 1) <= 2 * prm.n_tiles) {
        // Large footprints (~16 tiles per Gaussian): count with a DIFFERENCE GRID instead of one atomic per touched tile —
        // a rect adds +1 / -1 / -1 / +1 at its four corners of a (gx+1) x (gy+1) grid, and the 2-D prefix sum of the grid
        // is the per-tile count.  Four shared-memory atomics per Gaussian instead of one per instance.
        const int sx = prm.gx + 1, cells = sx * (prm.gy + 1);
        int* grid = reinterpret_cast<int*>(s_enum);
        __syncthreads();
        for (int i = threadIdx.x; i < cells; i += kBlock) grid[i] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kEnumItems; k++) {
            int x0, y0, x1, y1;
            if (rect_of(k, x0, y0, x1, y1) == 0) continue;
            atomicAdd(&grid[y0 * sx + x0], 1);
            atomicAdd(&grid[y0 * sx + x1], -1);
            atomicAdd(&grid[y1 * sx + x0], -1);
            atomicAdd(&grid[y1 * sx + x1], 1);
        }
        __syncthreads();
        for (int r = threadIdx.x; r <= prm.gy; r += kBlock) {  // prefix along x, one row per thread
            int run = 0;
            for (int x = 0; x < sx; x++) { run += grid[r * sx + x]; grid[r * sx + x] = run; }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < sx; c += kBlock) {       // prefix along y, one column per thread
            int run = 0;
            for (int y = 0; y <= prm.gy; y++) { run += grid[y * sx + c]; grid[y * sx + c] = run; }
        }
        __syncthreads();
        uint32_t mine = 0;
        for (int i = threadIdx.x; i < prm.n_tiles; i += kBlock) {
            const int ty = i / prm.gx, tx = i - ty * prm.gx;
            const uint32_t c = (uint32_t)grid[ty * sx + tx];
            if (c) atomicAdd(&counts[tile_base + i], c);
            mine += c;
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&view_totals[view], mine);
        return;
    }
    // (handing the count launch's per-CTA histogram to the scatter launch through global memory instead of
    // recounting was measured slower: 0.63 vs 0.53 ms — the recount sweep also warms L1 with the rows the scatter reads)
    __syncthreads();
    for (int i = threadIdx.x; i < prm.n_tiles; i += kBlock) s_hist[i] = 0u;
    if (small_grid) {
        uint32_t dbits[kEnumItems];
#pragma unroll
        for (int k = 0; k < kEnumItems; k++)
            dbits[k] = (SCATTER && rect[k]) ? __float_as_uint(depth[(size_t)view * prm.P + first + k * kBlock]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kEnumItems; k++)
            for_each_tile_of_rect(rect[k], prm.gx, 0u, 0u, [&](uint32_t tl, uint32_t, uint32_t) { atomicAdd(&s_hist[tl], 1u); });
        __syncthreads();
        if (!SCATTER) {
            uint32_t mine = 0;
            for (int i = threadIdx.x; i < prm.n_tiles; i += kBlock) {
                const uint32_t c = s_hist[i];
                if (c) atomicAdd(&counts[tile_base + i], c);
                mine += c;
            }
            mine = __reduce_add_sync(0xffffffffu, mine);
            if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&view_totals[view], mine);
        } else {
            for (int i = threadIdx.x; i < prm.n_tiles; i += kBlock) {
                const uint32_t c = s_hist[i];
                if (c) s_base[i] = ranges[tile_base + i].x + atomicSub(&counts[tile_base + i], c) - c;
                s_hist[i] = 0u;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kEnumItems; k++)
                for_each_tile_of_rect(rect[k], prm.gx, (uint32_t)((size_t)view * prm.P + first + k * kBlock), dbits[k],
                                      [&](uint32_t tl, uint32_t val, uint32_t db) {
                                          pairs[s_base[tl] + atomicAdd(&s_hist[tl], 1u)] = make_uint2(val, db);
                                      });
        }
        return;
    }
    // tile grids beyond 255 tiles in one direction: rects are re-derived in each pass
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < kEnumItems; k++)
        for_each_touched_tile<false>(prm, radii, xy, depth, view, first + k * kBlock,
                                     [&](uint32_t tl, uint32_t, uint32_t) { atomicAdd(&s_hist[tl], 1u); });
    __syncthreads();
    if (!SCATTER) {
        uint32_t mine = 0;
        for (int i = threadIdx.x; i < prm.n_tiles; i += kBlock) {
            const uint32_t c = s_hist[i];
            if (c) atomicAdd(&counts[tile_base + i], c);
            mine += c;
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&view_totals[view], mine);
    } else {
        for (int i = threadIdx.x; i < prm.n_tiles; i += kBlock) {
            const uint32_t c = s_hist[i];
            if (c) s_base[i] = ranges[tile_base + i].x + atomicSub(&counts[tile_base + i], c) - c;
            s_hist[i] = 0u;
        }
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < kEnumItems; k++)
            for_each_touched_tile<true>(prm, radii, xy, depth, view, first + k * kBlock,
                                        [&](uint32_t tl, uint32_t val, uint32_t dbits) {
                                            pairs[s_base[tl] + atomicAdd(&s_hist[tl], 1u)] = make_uint2(val, dbits);
                                        });
    }
}

// the same without the shared-memory stage, for views of more than kEnumMaxTiles tiles
template <bool SCATTER>
__global__ void __launch_bounds__(kBlock)
tile_enumerate_global_kernel(const RenderParams prm, const int32_t* __restrict__ radii, const float2* __restrict__ xy,
                             const float* __restrict__ depth, uint32_t* __restrict__ counts, uint32_t* __restrict__ view_totals,
                             const uint2* __restrict__ ranges, uint2* __restrict__ pairs)
{
    const int view = blockIdx.y;
    const uint32_t tile_base = (uint32_t)view * (uint32_t)prm.n_tiles;
    uint32_t mine = 0;
    for_each_touched_tile<SCATTER>(prm, radii, xy, depth, view, blockIdx.x * kBlock + threadIdx.x,
                                   [&](uint32_t tl, uint32_t val, uint32_t dbits) {
                                       const uint32_t gt = tile_base + tl;
                                       if (SCATTER) {
                                           const uint32_t slot = atomicSub(&counts[gt], 1u) - 1u;
                                           pairs[ranges[gt].x + slot] = make_uint2(val, dbits);
                                       } else {
                                           atomicAdd(&counts[gt], 1u);
                                           mine++;
                                       }
                                   });
    if (!SCATTER) {
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&view_totals[view], mine);
    }
}

// whether the per-CTA shared-memory stage is used (lgm_set_tuning "enum_global": test hook for the large-view variant)
bool enumerate_in_smem(const RenderParams& prm) { return prm.n_tiles <= kEnumMaxTiles && tuning(kTuneEnumGlobal) <= 0; }
uint32_t enum_ctas_per_view(const RenderParams& prm) { return (uint32_t)((prm.P + kBlock * kEnumItems - 1) / (kBlock * kEnumItems)); }

template <bool SCATTER>
cudaError_t launch_tile_enumerate(cudaStream_t stream, const RenderParams& prm, const int32_t* radii, const float2* xy,
                                  const float* depth, uint32_t* counts, uint32_t* view_totals, const uint2* ranges,
                                  uint2* pairs, uint32_t* super_counts = nullptr, uint32_t* view_entries = nullptr,
                                  const unsigned long long* total_instances = nullptr, unsigned long long coarse_gate = 0)
{
    if (enumerate_in_smem(prm)) {
        dim3 grid(enum_ctas_per_view(prm), prm.n_views);
        tile_enumerate_kernel<SCATTER><<<grid, kBlock, (size_t)prm.n_tiles * 8, stream>>>(prm, radii, xy, depth, counts, view_totals,
                                                                                          ranges, pairs, super_counts, view_entries,
                                                                                          total_instances, coarse_gate);
    } else {
        dim3 grid((prm.P + kBlock - 1) / kBlock, prm.n_views);
        tile_enumerate_global_kernel<SCATTER><<<grid, kBlock, 0, stream>>>(prm, radii, xy, depth, counts, view_totals, ranges, pairs);
    }
    return cudaGetLastError();
}

// head[] words of the direct path's scratch
enum Head { kHeadM = 0, kHeadMCursor = 1, kHeadLongest = 2, kHeadL = 3, kHeadLCursor = 4, kHeadX = 5, kHeadXCursor = 6,
            kHeadItems = 7, kHeadItemCursor = 8, kHeadItemOverflow = 9, kHeadS = 10, kHeadSCursor = 11 };

// D2.  One CTA per view: the view's first slot is the sum of the totals of the views before it, the tiles of the view
// are scanned in chunks of 256.  Writes ranges[] (empty tiles stay (0,0)), appends the non-empty tiles to the work lists
// (warp-aggregated) and records the longest tile.  Tiles of the M class (<= kSortCapM) fill `list` from the front, those
// of the L class (> kSortCapX) from the back; the X class in between fills `list_x` from the front, the S class
// (<= kSortCapS) from the back.
// Coarse grouping (super_counts != null): the same for the entries — super_offsets[] = first entry of every (view,
// super-tile) group in the grouped entry list, one work item (group, chunk) per kCoarseChunk entries of a group, and
// the total number of entries (the last view's CTA writes it to *entries_out).
__global__ void __launch_bounds__(kBlock)
tile_ranges_scan_kernel(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ view_totals, int n_tiles,
                        uint32_t n_ranges, uint2* __restrict__ ranges, uint32_t* __restrict__ list, uint32_t* __restrict__ list_x,
                        uint32_t* __restrict__ head, uint32_t* __restrict__ longest_out, const uint32_t* __restrict__ super_counts,
                        const uint32_t* __restrict__ view_entries, int n_super, uint32_t* __restrict__ super_offsets,
                        uint2* __restrict__ items, uint32_t item_capacity, uint32_t* __restrict__ entries_out)
{
    __shared__ uint32_t s_warp[8];
    const int view = blockIdx.x, t = threadIdx.x, lane = t & 31;
    uint32_t part = 0;
    for (int v = t; v < view; v += kBlock) part += view_totals[v];
    uint32_t carry;
    block_excl_scan_256(part, s_warp, &carry);
    uint32_t longest = 0;
    for (int i0 = 0; i0 < n_tiles; i0 += kBlock) {
        const int i = i0 + t;
        const uint32_t gt = (uint32_t)view * (uint32_t)n_tiles + (uint32_t)i;
        const uint32_t n = i < n_tiles ? counts[gt] : 0u;
        uint32_t tot;
        const uint32_t start = carry + block_excl_scan_256(n, s_warp, &tot);
        carry += tot;
        if (i < n_tiles) ranges[gt] = n ? make_uint2(start, start + n) : make_uint2(0u, 0u);
        longest = max(longest, n);
        const bool is_s = n != 0u && n <= (uint32_t)kSortCapS;
        const bool is_m = n > (uint32_t)kSortCapS && n <= (uint32_t)kSortCapM, is_l = n > (uint32_t)kSortCapX;
        const bool is_x = n > (uint32_t)kSortCapM && !is_l;
        const unsigned ms = __ballot_sync(0xffffffffu, is_s);
        if (ms) {
            const int leader = __ffs(ms) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&head[kHeadS], (uint32_t)__popc(ms));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (is_s) list_x[n_ranges - 1u - (base + __popc(ms & ((1u << lane) - 1u)))] = gt;
        }
        const unsigned m = __ballot_sync(0xffffffffu, is_m);
        if (m) {
            const int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&head[kHeadM], (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (is_m) list[base + __popc(m & ((1u << lane) - 1u))] = gt;
        }
        const unsigned ml = __ballot_sync(0xffffffffu, is_l);
        if (ml) {
            const int leader = __ffs(ml) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&head[kHeadL], (uint32_t)__popc(ml));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (is_l) list[n_ranges - 1u - (base + __popc(ml & ((1u << lane) - 1u)))] = gt;
        }
        const unsigned mx = __ballot_sync(0xffffffffu, is_x);
        if (mx) {
            const int leader = __ffs(mx) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&head[kHeadX], (uint32_t)__popc(mx));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (is_x) list_x[base + __popc(mx & ((1u << lane) - 1u))] = gt;
        }
    }
    longest = __reduce_max_sync(0xffffffffu, longest);
    if (lane == 0 && longest) {
        atomicMax(&head[kHeadLongest], longest);
        if (longest_out) atomicMax(longest_out, longest);  // the caller's step counters (read back with the instance count)
    }
    if (super_counts == nullptr) return;
    // ---- coarse groups of this view ----
    __syncthreads();
    part = 0;
    for (int v = t; v < view; v += kBlock) part += view_entries[v];
    uint32_t ecarry;
    block_excl_scan_256(part, s_warp, &ecarry);
    for (int i0 = 0; i0 < n_super; i0 += kBlock) {
        const int i = i0 + t;
        const uint32_t g = (uint32_t)view * (uint32_t)n_super + (uint32_t)i;
        const uint32_t n = i < n_super ? super_counts[g] : 0u;
        uint32_t tot;
        const uint32_t start = ecarry + block_excl_scan_256(n, s_warp, &tot);
        ecarry += tot;
        if (i < n_super) super_offsets[g] = start;
        const uint32_t chunks = (n + kCoarseChunk - 1) / kCoarseChunk;
        if (chunks) {
            const uint32_t base = atomicAdd(&head[kHeadItems], chunks);
            if (base + chunks <= item_capacity) {
                for (uint32_t c = 0; c < chunks; c++) items[base + c] = make_uint2(g, c);
            } else {
                head[kHeadItemOverflow] = 1u;
            }
        }
    }
    if (view == (int)gridDim.x - 1 && t == 0 && entries_out) *entries_out = ecarry;
}

// Coarse scatter: the (view, Gaussian) pairs grouped by super-tile.  Same CTA shape as the count kernel; per touched
// super-tile one 16-byte entry (value, depth bits, tile rect clipped to the super-tile packed as x0 | y0 << 8 | x1 << 16 |
// y1 << 24, view) at the next slot of the group's run.
__global__ void __launch_bounds__(kBlock)
coarse_scatter_kernel(const RenderParams prm, const int32_t* __restrict__ radii, const float2* __restrict__ xy,
                      const float* __restrict__ depth, const uint32_t* __restrict__ super_offsets,
                      uint32_t* __restrict__ super_cursor, uint4* __restrict__ entries)
{
    extern __shared__ uint32_t s_enum[];
    const int nsx = supers_x(prm), n_super = nsx * supers_y(prm);
    uint32_t* s_cnt = s_enum;             // [n_super]
    uint32_t* s_base = s_enum + n_super;  // [n_super]
    const int view = blockIdx.y;
    const int first = blockIdx.x * (kBlock * kEnumItems) + threadIdx.x;
    for (int i = threadIdx.x; i < n_super; i += kBlock) s_cnt[i] = 0u;
    // (the grouping is only used for tile grids up to 255 x 255: the rects are loaded up front and kept in registers)
    uint32_t rect[kEnumItems], dbits[kEnumItems];
    load_rects_packed(prm, radii, xy, view, first, rect);
#pragma unroll
    for (int k = 0; k < kEnumItems; k++) dbits[k] = rect[k] ? __float_as_uint(depth[(size_t)view * prm.P + first + k * kBlock]) : 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kEnumItems; k++) {
        if (rect[k] == 0u) continue;
        const int x0 = rect[k] & 0xffu, y0 = (rect[k] >> 8) & 0xffu, x1 = (rect[k] >> 16) & 0xffu, y1 = rect[k] >> 24;
        for (int sy = y0 >> kSuperShift; sy <= (y1 - 1) >> kSuperShift; sy++)
            for (int sx = x0 >> kSuperShift; sx <= (x1 - 1) >> kSuperShift; sx++) atomicAdd(&s_cnt[sy * nsx + sx], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_super; i += kBlock) {
        const uint32_t c = s_cnt[i];
        const size_t g = (size_t)view * n_super + i;
        if (c) s_base[i] = super_offsets[g] + atomicAdd(&super_cursor[g], c);
        s_cnt[i] = 0u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kEnumItems; k++) {
        if (rect[k] == 0u) continue;
        const size_t gi = (size_t)view * prm.P + first + k * kBlock;
        const int x0 = rect[k] & 0xffu, y0 = (rect[k] >> 8) & 0xffu, x1 = (rect[k] >> 16) & 0xffu, y1 = rect[k] >> 24;
        for (int sy = y0 >> kSuperShift; sy <= (y1 - 1) >> kSuperShift; sy++)
            for (int sx = x0 >> kSuperShift; sx <= (x1 - 1) >> kSuperShift; sx++) {
                const int s = sy * nsx + sx;
                const int cx0 = max(x0, sx << kSuperShift), cx1 = min(x1, (sx + 1) << kSuperShift);
                const int cy0 = max(y0, sy << kSuperShift), cy1 = min(y1, (sy + 1) << kSuperShift);
                const uint32_t r = (uint32_t)cx0 | (uint32_t)cy0 << 8 | (uint32_t)cx1 << 16 | (uint32_t)cy1 << 24;
                entries[s_base[s] + atomicAdd(&s_cnt[s], 1u)] = make_uint4((uint32_t)gi, dbits[k], r, (uint32_t)view);
            }
    }
}

// Fine scatter from the grouped entries.  Persistent CTAs pull (group, chunk) work items; all entries of an item lie in
// ONE super-tile, so the CTA's histogram has 64 counters and its stores go to the 64 tile segments of that super-tile in
// long runs.  Same reservation protocol as the plain scatter (counts[] counted down by the CTA's number).
__global__ void __launch_bounds__(kBlock, 4)
tile_scatter_entries_kernel(const RenderParams prm, const uint4* __restrict__ entries, const uint32_t* __restrict__ super_offsets,
                            const uint32_t* __restrict__ super_counts, const uint2* __restrict__ items,
                            const uint32_t* __restrict__ head, uint32_t* __restrict__ item_cursor, uint32_t* __restrict__ counts,
                            const uint2* __restrict__ ranges, uint2* __restrict__ pairs, int tile_major)
{
    constexpr int kFine = kSuperTiles * kSuperTiles, kD = kSuperTiles + 1;
    __shared__ uint32_t s_cnt[kFine], s_base[kFine];
    __shared__ uint32_t s_ev[kCoarseChunk], s_ed[kCoarseChunk], s_er[kCoarseChunk];  // tile_major: the item's entries
    __shared__ int s_diff[kD * kD];
    __shared__ uint32_t s_item;
    const int nsx = supers_x(prm), n_super = nsx * supers_y(prm);
    const uint32_t n_items = head[kHeadItems];
    const int lane = threadIdx.x & 31;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(item_cursor, 1u);
        if (threadIdx.x < kFine) s_cnt[threadIdx.x] = 0u;
        if (threadIdx.x < kD * kD) s_diff[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const uint2 it = items[item];
        const uint32_t g = it.x;
        const int view = (int)(g / (uint32_t)n_super), s = (int)(g - (uint32_t)view * (uint32_t)n_super);
        const int sx0 = (s % nsx) << kSuperShift, sy0 = (s / nsx) << kSuperShift;
        const uint32_t e0 = super_offsets[g] + it.y * kCoarseChunk;
        const uint32_t n = min((uint32_t)kCoarseChunk, super_counts[g] - it.y * kCoarseChunk);
        const uint32_t tile_base = (uint32_t)view * (uint32_t)prm.n_tiles;
        // This thread's entries of the item, loaded up front (independent 16-byte loads) and kept in registers for both
        // sweeps — value, depth bits, rect: in the first version every entry of every sweep waited for its own load (24 % of
        // the kernel's stall samples, source view of the ncu capture).
        constexpr int kPer = kCoarseChunk / kBlock;
        uint32_t ev[kPer], ed[kPer], er[kPer];
#pragma unroll
        for (int u = 0; u < kPer; u++) {
            const uint32_t i = threadIdx.x + (uint32_t)u * kBlock;
            uint4 e = make_uint4(0u, 0u, 0u, 0u);
            if (i < n) e = __ldg(entries + e0 + i);
            ev[u] = e.x; ed[u] = e.y; er[u] = e.z;
        }
        // Count per fine tile with a 9 x 9 DIFFERENCE GRID: a rect adds +1 / -1 / -1 / +1 at its corners and the 2-D prefix
        // sum is the per-tile count — four shared-memory atomics per entry instead of one per instance.
#pragma unroll
        for (int u = 0; u < kPer; u++) {
            if (threadIdx.x + (uint32_t)u * kBlock >= n) continue;
            const int x0 = (int)(er[u] & 255u) - sx0, y0 = (int)((er[u] >> 8) & 255u) - sy0;
            const int x1 = (int)((er[u] >> 16) & 255u) - sx0, y1 = (int)(er[u] >> 24) - sy0;
            atomicAdd(&s_diff[y0 * kD + x0], 1);
            atomicAdd(&s_diff[y0 * kD + x1], -1);
            atomicAdd(&s_diff[y1 * kD + x0], -1);
            atomicAdd(&s_diff[y1 * kD + x1], 1);
        }
        __syncthreads();
        if (threadIdx.x < kD) {
            int run = 0;
            for (int x = 0; x < kD; x++) { run += s_diff[threadIdx.x * kD + x]; s_diff[threadIdx.x * kD + x] = run; }
        }
        __syncthreads();
        if (threadIdx.x < kD) {
            int run = 0;
            for (int y = 0; y < kD; y++) { run += s_diff[y * kD + threadIdx.x]; s_diff[y * kD + threadIdx.x] = run; }
        }
        __syncthreads();
        // reserve this item's run of every touched tile's segment (counts[] holds the totals and is counted down)
        if (threadIdx.x < kFine) {
            const int lx = threadIdx.x & (kSuperTiles - 1), ly = threadIdx.x >> kSuperShift;
            const uint32_t c = (uint32_t)s_diff[ly * kD + lx];
            if (c) {  // (c != 0 implies the tile exists: rects are clipped to the tile grid)
                const uint32_t gt = tile_base + (uint32_t)((sy0 + ly) * prm.gx + sx0 + lx);
                s_base[threadIdx.x] = ranges[gt].x + atomicSub(&counts[gt], c) - c;
            }
        }
        __syncthreads();
        if (tile_major) {
            // TILE-MAJOR placement: a warp owns eight of the 64 fine tiles; for each it sweeps the item's entries (staged in
            // shared memory), and the lanes whose rect contains the tile write their pair at consecutive slots of the tile's
            // run (ballot + popc): no shared-memory atomics, and every store of a warp lands in one run — full 32-byte
            // sectors instead of one sector per 8-byte pair (the entry-major walk wrote 461 M sectors for 499 M pairs and
            // kept L2 at 60 % of its throughput).
#pragma unroll
            for (int u = 0; u < kPer; u++) {
                const uint32_t i = threadIdx.x + (uint32_t)u * kBlock;
                if (i < n) {
                    s_ev[i] = ev[u];
                    s_ed[i] = ed[u];
                    // rect relative to the super-tile, one byte per coordinate: x0 | y0 << 8 | x1 << 16 | y1 << 24
                    s_er[i] = er[u] - ((uint32_t)sx0 | (uint32_t)sy0 << 8 | (uint32_t)sx0 << 16 | (uint32_t)sy0 << 24);
                }
            }
            __syncthreads();
            const int warp = threadIdx.x >> 5;
            const uint32_t lt_mask = (1u << lane) - 1u;
            for (int f = warp; f < kFine; f += kBlock / 32) {
                const int fx = f & (kSuperTiles - 1), fy = f >> kSuperShift;
                if (s_diff[fy * kD + fx] == 0) continue;  // no instance of this item in the tile
                uint32_t out = s_base[f];
                for (uint32_t i0 = 0; i0 < n; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    bool inside = false;
                    if (i < n) {
                        const uint32_t r = s_er[i];
                        inside = (uint32_t)fx >= (r & 255u) && (uint32_t)fx < ((r >> 16) & 255u) &&
                                 (uint32_t)fy >= ((r >> 8) & 255u) && (uint32_t)fy < (r >> 24);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, inside);
                    if (inside) pairs[out + __popc(m & lt_mask)] = make_uint2(s_ev[i], s_ed[i]);
                    out += __popc(m);
                }
            }
            continue;  // (the loop head synchronises before shared memory is reused)
        }
        // place: every entry's rect (<= 64 tiles, ~12 on average) is enumerated by its thread, or by the whole warp when it
        // is large (kCoopFine: the warp-wide walk costs ~10 x a thread's own walk per instance, so only rects that would
        // stall the other lanes for long take it)
#pragma unroll
        for (int u = 0; u < kPer; u++) {
            const bool in_range = threadIdx.x + (uint32_t)u * kBlock < n;
            const int x0 = (int)(er[u] & 255u) - sx0, y0 = (int)((er[u] >> 8) & 255u) - sy0;
            const int w = (int)((er[u] >> 16) & 255u) - sx0 - x0, h = (int)(er[u] >> 24) - sy0 - y0;
            const uint32_t area = in_range ? (uint32_t)(w * h) : 0u;
            const bool big = area > kCoopFine;
            if (area != 0u && !big) {
                for (int y = y0; y < y0 + h; y++)
                    for (int x = x0; x < x0 + w; x++) {
                        const int f = y * kSuperTiles + x;
                        pairs[s_base[f] + atomicAdd(&s_cnt[f], 1u)] = make_uint2(ev[u], ed[u]);
                    }
            }
            unsigned m = __ballot_sync(0xffffffffu, big);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t a = __shfl_sync(0xffffffffu, area, src);
                const uint32_t pk = __shfl_sync(0xffffffffu, (uint32_t)x0 | (uint32_t)y0 << 8 | (uint32_t)w << 16, src);
                const uint32_t sv = __shfl_sync(0xffffffffu, ev[u], src), sd = __shfl_sync(0xffffffffu, ed[u], src);
                const int bx0 = pk & 255u, by0 = (pk >> 8) & 255u;
                const uint32_t bw = pk >> 16;                       // 1..8
                const uint32_t inv = (65536u + bw - 1u) / bw;       // q / bw == (q * inv) >> 16 for q < 64 (warp-uniform)
                for (uint32_t q = lane; q < a; q += 32) {
                    const uint32_t ry = (q * inv) >> 16, rx = q - ry * bw;
                    const int f = (by0 + (int)ry) * kSuperTiles + bx0 + (int)rx;
                    pairs[s_base[f] + atomicAdd(&s_cnt[f], 1u)] = make_uint2(sv, sd);
                }
            }
        }
    }
}

// D4.  Persistent CTAs pull tiles from a work list: list[item * list_step] for item < *n_list_ptr, items handed out
// through *cursor.
template <int T, int CAP, int LG_MAXB, int MIN_BLOCKS, bool BULK>
__global__ void __launch_bounds__(T, MIN_BLOCKS)
tile_bucket_sort_kernel(const uint2* __restrict__ pairs, const uint2* __restrict__ ranges, const uint32_t* __restrict__ list,
                        int list_step, const uint32_t* __restrict__ n_list_ptr, uint32_t* __restrict__ cursor,
                        uint32_t* __restrict__ vals_sorted, uint64_t* __restrict__ keys_sorted)
{
    constexpr int kSortCap = CAP, kMaxBuckets = 1 << LG_MAXB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* A_raw = reinterpret_cast<uint64_t*>(smem_raw);                                 // [kSortCap + 2] depth << 32 | value
    uint16_t* order = reinterpret_cast<uint16_t*>(smem_raw + (size_t)kSortCap * 8 + 16);      // [kSortCap] bucket-grouped indices
    uint32_t* bucket = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kSortCap * 10 + 16);    // [kMaxBuckets + 1]
    __shared__ __align__(8) uint64_t s_mbar;
    uint32_t* s_red = bucket + kMaxBuckets + 1;                                          // [64]: min and max per warp
    __shared__ uint32_t s_item;
    constexpr int kWarps = T / 32;
    static_assert(CAP < 65536 && (kWarps & (kWarps - 1)) == 0 && kWarps <= 32, "16-bit indices, power-of-two warps");
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t n_list = *n_list_ptr;
    if (BULK && t == 0) mbar_init(&s_mbar, 1u);
    uint32_t parity = 0;

    while (true) {
        if (t == 0) s_item = atomicAdd(cursor, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_list) break;
        const uint32_t tile = list[(ptrdiff_t)item * list_step];
        const uint2 range = ranges[tile];
        const int n = (int)(range.y - range.x);
        const uint2* src = pairs + range.x;
        // a stored pair (value, depth bits) read as one little-endian 64-bit word IS the key depth << 32 | value: staging is a
        // plain copy of a contiguous segment -> one 1-D bulk copy (TMA) instead of a load / store loop.  The copy engine
        // wants 16-byte alignment: it starts one pair early when the segment starts at an odd index
        const uint32_t lead = BULK ? (range.x & 1u) : 0u;
        uint64_t* A = A_raw + lead;
        if (n > kSortCap) {  // cannot happen: api.cu only takes this path when the longest tile fits
            __syncthreads();
            continue;
        }

        // number of buckets: a power of two in [n/2, n), 32..kMaxBuckets (n up to 5 kMaxBuckets in the L class)
        int lg_nb = 32 - __clz((unsigned)max(n - 1, 1)) - 1;
        lg_nb = min(max(lg_nb, 5), LG_MAXB);
        const int nb = 1 << lg_nb;

        // sweep 0: stage, min / max of the depth bits
        uint32_t dmin = 0xffffffffu, dmax = 0u;
        if (BULK) {
            if (t == 0) {
                const uint32_t bytes = (((uint32_t)n + lead + 1u) & ~1u) * 8u;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was last touched by ordinary loads / stores
                mbar_expect_tx(&s_mbar, bytes);
                bulk_load(A_raw, src - lead, bytes, &s_mbar);
            }
            mbar_wait(&s_mbar, parity);
            parity ^= 1u;
            for (int i = t; i < n; i += T) {
                const uint32_t d = (uint32_t)(A[i] >> 32);
                dmin = min(dmin, d);
                dmax = max(dmax, d);
            }
        } else {
            for (int i = t; i < n; i += T) {
                const uint2 p = src[i];
                A[i] = ((uint64_t)p.y << 32) | p.x;
                dmin = min(dmin, p.y);
                dmax = max(dmax, p.y);
            }
        }
        dmin = __reduce_min_sync(0xffffffffu, dmin);
        dmax = __reduce_max_sync(0xffffffffu, dmax);
        if (lane == 0) { s_red[warp] = dmin; s_red[kWarps + warp] = dmax; }
        for (int i = t; i <= nb; i += T) bucket[i] = 0u;
        __syncthreads();  // also orders the read of s_item above against the next iteration's write
        dmin = __reduce_min_sync(0xffffffffu, s_red[lane & (kWarps - 1)]);
        dmax = __reduce_max_sync(0xffffffffu, s_red[kWarps + (lane & (kWarps - 1))]);
        // monotone map of depth bits to buckets: floor((d - dmin) * mul / 2^32) with mul <= nb * 2^32 / (span + 1), so
        // that the largest key lands below nb; any smaller mul is still monotone, so a 32-bit quotient that rounds
        // the divisor up is enough (it leaves < 0.1 % of the buckets unused at the usual span / nb ~ 2^11).  d - dmin
        // itself when the span is below the bucket count.
        const uint32_t span = dmax - dmin;
        const bool direct = span < (uint32_t)nb;
        const uint32_t mul = direct ? 0u : 0xffffffffu / (((span + 1u) >> lg_nb) + 1u);
#define LGM_BUCKET(d) (direct ? ((d) - dmin) : __umulhi((d) - dmin, mul))

        // sweep 1: bucket sizes
        for (int i = t; i < n; i += T) atomicAdd(&bucket[LGM_BUCKET((uint32_t)(A[i] >> 32))], 1u);
        __syncthreads();

        // exclusive scan of the nb sizes (each thread owns nb / T consecutive buckets, or one when nb < T); the sum of the
        // squared sizes — the number of compares the rank loop of sweep 3 will make — is formed on the way
        uint32_t rank_work = 0;  // <= n^2 < 2^32
        {
            const int per = nb >= T ? nb / T : 1;
            const int b0 = t * per;
            uint32_t sum = 0;
            if (b0 < nb)
                for (int j = 0; j < per; j++) {
                    const uint32_t c = bucket[b0 + j];
                    sum += c;
                    rank_work += c * c;
                }
            const uint32_t incl = warp_incl_scan(sum, lane);
            rank_work = __reduce_add_sync(0xffffffffu, rank_work);
            if (lane == 31) { s_red[warp] = incl; s_red[kWarps + warp] = rank_work; }
            __syncthreads();
            rank_work = __reduce_add_sync(0xffffffffu, lane < kWarps ? s_red[kWarps + lane] : 0u);
            // comparators of the sorting network below: n / 2 per stage, L (L + 1) / 2 stages, L = ceil(log2 n); one of
            // them costs about twice a compare of the rank loop
            const uint32_t lg_n = 32u - (uint32_t)__clz((unsigned)max(n - 1, 1));
            if (rank_work > (uint32_t)n * (lg_n * (lg_n + 1u) / 2u)) {
                // Depth ties (a plane at constant view depth, duplicated points, positions clamped to the same value) put
                // many keys into one bucket, and the rank loop of sweep 3 is quadratic in the bucket size.  Such a tile is
                // ordered by a sorting network instead: bitonic merges in their all-ascending form (first step of a merge
                // pairs i with i ^ (k - 1)), which needs no padding — a partner beyond n is a virtual +infinity that no
                // comparator moves.  n log^2 n / 4 comparators whatever the keys are.
                __syncthreads();
                for (uint32_t k = 2; (k >> 1) < (uint32_t)n; k <<= 1) {
                    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                        const uint32_t flip = (j == (k >> 1)) ? (k - 1u) : j;
                        for (uint32_t i = t; i < (uint32_t)n; i += T) {
                            const uint32_t p = i ^ flip;
                            if (p > i && p < (uint32_t)n) {
                                const uint64_t a = A[i], b = A[p];
                                if (b < a) { A[i] = b; A[p] = a; }
                            }
                        }
                        __syncthreads();
                    }
                }
                const uint64_t hi_t = (uint64_t)tile << 32;
                for (int i = t; i < n; i += T) {
                    const uint64_t key = A[i];
                    vals_sorted[range.x + i] = (uint32_t)key;
                    if (keys_sorted) keys_sorted[range.x + i] = hi_t | (key >> 32);
                }
                __syncthreads();  // shared memory is reused by the next tile
                continue;
            }
            const uint32_t base = __reduce_add_sync(0xffffffffu, lane < warp ? s_red[lane] : 0u);  // warp < kWarps <= 32
            uint32_t run = base + incl - sum;
            if (b0 < nb)
                for (int j = 0; j < per; j++) {
                    const uint32_t c = bucket[b0 + j];
                    bucket[b0 + j] = run;
                    run += c;
                }
        }
        __syncthreads();

        // sweep 2: group the element indices by bucket; bucket[b] runs from the bucket's start to its end
        for (int i = t; i < n; i += T) {
            const uint32_t pos = atomicAdd(&bucket[LGM_BUCKET((uint32_t)(A[i] >> 32))], 1u);
            order[pos] = (uint16_t)i;
        }
        __syncthreads();

        // sweep 3: final slot = bucket start + number of smaller keys in the bucket
        const uint64_t hi = (uint64_t)tile << 32;
        for (int p = t; p < n; p += T) {
            const uint64_t key = A[order[p]];
            const uint32_t b = LGM_BUCKET((uint32_t)(key >> 32));
            const uint32_t s0 = b ? bucket[b - 1] : 0u, e0 = bucket[b];
            uint32_t rank = 0;
            for (uint32_t q = s0; q < e0; q++) rank += (A[order[q]] < key) ? 1u : 0u;
            const uint32_t out = range.x + s0 + rank;
            vals_sorted[out] = (uint32_t)key;
            if (keys_sorted) keys_sorted[out] = hi | (key >> 32);
        }
#undef LGM_BUCKET
        __syncthreads();  // shared memory is reused by the next tile
    }
}

// D4, second form ("grouped keys").  The first form stages the segment (A), groups 16-bit INDICES by bucket and ranks
// every element against A[order[q]] — three random shared-memory gathers per compare chain, and the kernel is bound by
// the bandwidth of the shared-memory pipe under bank conflicts (ncu: L1 82 %).  Here the segment is not staged at all: the
// three sweeps that need the keys in input order (min / max, bucket sizes, grouping) read them straight from global
// memory — coalesced 8-byte loads of a segment that stays in L1 / L2 between the sweeps — and the grouping sweep scatters
// the KEYS themselves into shared memory.  The rank sweep then reads its own key at its own index and scans its bucket,
// which neighbouring lanes share or adjoin: nearly conflict-free.  8 B of shared memory per instance instead of 10.
template <int T, int CAP, int LG_MAXB, int MIN_BLOCKS>
__global__ void __launch_bounds__(T, MIN_BLOCKS)
tile_group_sort_kernel(const uint2* __restrict__ pairs, const uint2* __restrict__ ranges, const uint32_t* __restrict__ list,
                       int list_step, const uint32_t* __restrict__ n_list_ptr, uint32_t* __restrict__ cursor,
                       uint32_t* __restrict__ vals_sorted, uint64_t* __restrict__ keys_sorted)
{
    constexpr int kSortCap = CAP, kMaxBuckets = 1 << LG_MAXB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* G = reinterpret_cast<uint64_t*>(smem_raw);                                  // [kSortCap] keys grouped by bucket
    uint32_t* bucket = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kSortCap * 8);      // [kMaxBuckets + 1]
    uint32_t* s_red = bucket + kMaxBuckets + 1;                                           // [64]
    __shared__ uint32_t s_item, s_tile;
    __shared__ uint2 s_range;
    constexpr int kWarps = T / 32;
    static_assert((kWarps & (kWarps - 1)) == 0 && kWarps <= 32, "power-of-two warps");
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t n_list = *n_list_ptr;

    // Work items are fetched ONE TILE AHEAD: a tile starts with a chain of dependent global round trips (cursor atomic ->
    // list entry -> range -> first touch of the segment, ~3 us) that only the other two or five CTAs of the SM could
    // hide.  Thread 0 draws the next item while the current tile's first sweeps run, resolves its tile and range during
    // the last sweeps, and asks L2 for the next segment (cp.async.bulk.prefetch.L2).
    if (t == 0) {
        const uint32_t item = atomicAdd(cursor, 1u);
        s_item = item;
        if (item < n_list) {
            const uint32_t tl = list[(ptrdiff_t)item * list_step];
            s_tile = tl;
            s_range = ranges[tl];
        }
    }
    __syncthreads();
    while (true) {
        const uint32_t item = s_item;
        if (item >= n_list) break;
        const uint32_t tile = s_tile;
        const uint2 range = s_range;
        uint32_t next_item = 0xffffffffu;
        if (t == 0) next_item = atomicAdd(cursor, 1u);  // in flight during the first sweeps
        const int n = (int)(range.y - range.x);
        // a stored pair (value, depth bits) read as one little-endian 64-bit word IS the key depth << 32 | value
        const uint64_t* __restrict__ src = reinterpret_cast<const uint64_t*>(pairs + range.x);
        uint32_t next_tile = 0;
        uint2 next_range = make_uint2(0u, 0u);
        if (n > kSortCap) {  // cannot happen: api.cu only takes this path when the longest tile fits
            __syncthreads();
            if (t == 0) {
                s_item = next_item;
                if (next_item < n_list) { s_tile = list[(ptrdiff_t)next_item * list_step]; s_range = ranges[s_tile]; }
            }
            __syncthreads();
            continue;
        }
        int lg_nb = 32 - __clz((unsigned)max(n - 1, 1)) - 1;
        lg_nb = min(max(lg_nb, 5), LG_MAXB);
        const int nb = 1 << lg_nb;

        // sweep 0: min / max of the depth bits
        uint32_t dmin = 0xffffffffu, dmax = 0u;
        for (int i = t; i < n; i += T) {
            const uint32_t d = (uint32_t)(src[i] >> 32);
            dmin = min(dmin, d);
            dmax = max(dmax, d);
        }
        dmin = __reduce_min_sync(0xffffffffu, dmin);
        dmax = __reduce_max_sync(0xffffffffu, dmax);
        if (lane == 0) { s_red[warp] = dmin; s_red[kWarps + warp] = dmax; }
        for (int i = t; i <= nb; i += T) bucket[i] = 0u;
        __syncthreads();  // also orders the read of s_item above against the next iteration's write
        dmin = __reduce_min_sync(0xffffffffu, s_red[lane & (kWarps - 1)]);
        dmax = __reduce_max_sync(0xffffffffu, s_red[kWarps + (lane & (kWarps - 1))]);
        const uint32_t span = dmax - dmin;
        const bool direct = span < (uint32_t)nb;
        const uint32_t mul = direct ? 0u : 0xffffffffu / (((span + 1u) >> lg_nb) + 1u);
#define LGM_BUCKET(d) (direct ? ((d) - dmin) : __umulhi((d) - dmin, mul))

        // sweep 1: bucket sizes (the barrier after it also protects s_red, which the scan reuses)
        for (int i = t; i < n; i += T) atomicAdd(&bucket[LGM_BUCKET((uint32_t)(src[i] >> 32))], 1u);
        __syncthreads();

        // exclusive scan of the sizes + the compares the rank sweep will make (sum of the squared sizes)
        uint32_t rank_work = 0;
        {
            const int per = nb >= T ? nb / T : 1;
            const int b0 = t * per;
            uint32_t sum = 0;
            if (b0 < nb)
                for (int j = 0; j < per; j++) {
                    const uint32_t c = bucket[b0 + j];
                    sum += c;
                    rank_work += c * c;
                }
            const uint32_t incl = warp_incl_scan(sum, lane);
            rank_work = __reduce_add_sync(0xffffffffu, rank_work);
            if (lane == 31) { s_red[warp] = incl; s_red[kWarps + warp] = rank_work; }
            __syncthreads();
            rank_work = __reduce_add_sync(0xffffffffu, lane < kWarps ? s_red[kWarps + lane] : 0u);
            const uint32_t lg_n = 32u - (uint32_t)__clz((unsigned)max(n - 1, 1));
            if (rank_work > (uint32_t)n * (lg_n * (lg_n + 1u) / 2u)) {
                // depth ties: the bitonic network of the first form, on a plain copy of the segment
                for (int i = t; i < n; i += T) G[i] = src[i];
                __syncthreads();
                for (uint32_t k = 2; (k >> 1) < (uint32_t)n; k <<= 1) {
                    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                        const uint32_t flip = (j == (k >> 1)) ? (k - 1u) : j;
                        for (uint32_t i = t; i < (uint32_t)n; i += T) {
                            const uint32_t p = i ^ flip;
                            if (p > i && p < (uint32_t)n) {
                                const uint64_t a = G[i], b = G[p];
                                if (b < a) { G[i] = b; G[p] = a; }
                            }
                        }
                        __syncthreads();
                    }
                }
                const uint64_t hi_t = (uint64_t)tile << 32;
                for (int i = t; i < n; i += T) {
                    const uint64_t key = G[i];
                    vals_sorted[range.x + i] = (uint32_t)key;
                    if (keys_sorted) keys_sorted[range.x + i] = hi_t | (key >> 32);
                }
                if (t == 0) {
                    s_item = next_item;
                    if (next_item < n_list) { s_tile = list[(ptrdiff_t)next_item * list_step]; s_range = ranges[s_tile]; }
                }
                __syncthreads();
                continue;
            }
            const uint32_t base = __reduce_add_sync(0xffffffffu, lane < warp ? s_red[lane] : 0u);
            uint32_t run = base + incl - sum;
            if (b0 < nb)
                for (int j = 0; j < per; j++) {
                    const uint32_t c = bucket[b0 + j];
                    bucket[b0 + j] = run;
                    run += c;
                }
        }
        __syncthreads();
        // the next item's tile and range: two dependent loads, in flight during sweep 2
        if (t == 0 && next_item < n_list) {
            next_tile = list[(ptrdiff_t)next_item * list_step];
            next_range = ranges[next_tile];
        }

        // sweep 2: the keys, grouped by bucket; bucket[b] runs from the bucket's start to its end.  (Keeping the arrival
        // index the counting atomic returns, so that this sweep needs no second atomic, was measured slower: 2 more bytes of
        // shared memory per instance cost a resident CTA, and the atomics are not what bounds the kernel.)
        for (int i = t; i < n; i += T) {
            const uint64_t key = src[i];
            G[atomicAdd(&bucket[LGM_BUCKET((uint32_t)(key >> 32))], 1u)] = key;
        }
        __syncthreads();
        if (t == 0 && next_item < n_list && next_range.y > next_range.x) {
            // ask L2 for the next segment (16-byte granules around it)
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(pairs + next_range.x) & ~(uintptr_t)15;
            const uintptr_t a1 = (reinterpret_cast<uintptr_t>(pairs + next_range.y) + 15) & ~(uintptr_t)15;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((uint32_t)(a1 - a0)) : "memory");
        }

        // sweep 3: final slot = bucket start + number of smaller keys in the bucket
        const uint64_t hi = (uint64_t)tile << 32;
        for (int p = t; p < n; p += T) {
            const uint64_t key = G[p];
            const uint32_t b = LGM_BUCKET((uint32_t)(key >> 32));
            const uint32_t s0 = b ? bucket[b - 1] : 0u, e0 = bucket[b];
            uint32_t rank = 0;
            for (uint32_t q = s0; q < e0; q++) rank += (G[q] < key) ? 1u : 0u;
            const uint32_t out = range.x + s0 + rank;
            vals_sorted[out] = (uint32_t)key;
            if (keys_sorted) keys_sorted[out] = hi | (key >> 32);
        }
#undef LGM_BUCKET
        if (t == 0) {
            s_item = next_item;
            s_tile = next_tile;
            s_range = next_range;
        }
        __syncthreads();  // shared memory is reused by the next tile; the next item is published
    }
}
constexpr size_t group_sort_smem(int cap, int lg_buckets) { return (size_t)cap * 8 + (size_t)((1 << lg_buckets) + 1) * 4 + 64 * 4; }

}  // namespace

int direct_bin_tile_cap() { return kSortCapL; }

// scratch of the direct path.  Zeroed at the head of every step: head [64 u32] | view totals [nv] | view entries [nv] |
// counts [n_ranges] | super counts [n_groups] | super cursor [n_groups]; not zeroed: list [n_ranges] | list_x [n_ranges] |
// super offsets [n_groups] | items [item_capacity x uint2]
struct DirectScratch {
    uint32_t *head, *view_totals, *view_entries, *counts, *super_counts, *super_cursor, *list, *list_x, *super_offsets;
    uint2* items;
    uint32_t item_capacity, n_super;
    bool coarse;  // whether the coarse grouping can be used at all for this shape
    size_t zero_bytes, total_bytes;
};
static DirectScratch direct_scratch(const RenderParams& prm, void* scratch)
{
    const size_t n_ranges = (size_t)prm.n_views * prm.n_tiles;
    const size_t nv = ((size_t)prm.n_views + 63) / 64 * 64;
    DirectScratch d;
    d.n_super = (uint32_t)(supers_x(prm) * supers_y(prm));
    // packed 8-bit rect coordinates and the per-CTA shared-memory stage bound the shapes the grouping serves
    d.coarse = prm.gx <= 255 && prm.gy <= 255 && d.n_super <= kCoarseMaxSupers && prm.n_tiles <= kEnumMaxTiles;
    const size_t n_groups = ((size_t)prm.n_views * d.n_super + 63) / 64 * 64;
    d.item_capacity = (uint32_t)(((size_t)prm.n_views * prm.P * 4) / kCoarseChunk + n_groups);
    d.head = static_cast<uint32_t*>(scratch);
    d.view_totals = d.head + 64;
    d.view_entries = d.view_totals + nv;
    d.counts = d.view_entries + nv;
    d.super_counts = d.counts + n_ranges;
    d.super_cursor = d.super_counts + n_groups;
    d.list = d.super_cursor + n_groups;
    d.list_x = d.list + n_ranges;
    d.super_offsets = d.list_x + n_ranges;
    uint32_t* end = d.super_offsets + n_groups;
    end += (8 - ((size_t)(end - d.head) & 7)) & 7;  // items are 8-byte aligned (the scratch itself is 256-byte aligned)
    d.items = reinterpret_cast<uint2*>(end);
    d.zero_bytes = (size_t)(d.list - d.head) * sizeof(uint32_t);
    d.total_bytes = (size_t)(end - d.head) * sizeof(uint32_t) + (size_t)d.item_capacity * sizeof(uint2) + 256;
    return d;
}

size_t direct_bin_scratch_bytes(const RenderParams& prm) { return direct_scratch(prm, nullptr).total_bytes; }

// D1 + D2: after this, ranges[] is final for the direct path, *longest_out (device, zeroed here) holds the longest
// tile and *entries_out the number of coarse entries (0 when the shape does not allow the grouping).  `scratch` as
// direct_bin_scratch_bytes; it must reach launch_direct_bin_sort untouched.
// minimum instances per coarse entry from which the grouping is used (lgm_set_tuning "coarse_ratio"; 0 = never)
static uint64_t coarse_ratio()
{
    const int t = tuning(kTuneCoarseRatio);
    return t >= 0 ? (uint64_t)t : 6;
}

cudaError_t launch_direct_bin_count(cudaStream_t stream, const RenderParams& prm, const int32_t* radii, const float2* xy,
                                    uint2* ranges, void* scratch, uint32_t* longest_out, uint32_t* entries_out,
                                    const unsigned long long* total_instances)
{
    const DirectScratch d = direct_scratch(prm, scratch);
    cudaError_t err = cudaMemsetAsync(scratch, 0, d.zero_bytes, stream);
    if (err != cudaSuccess) return err;
    if (longest_out && (err = cudaMemsetAsync(longest_out, 0, sizeof(uint32_t), stream)) != cudaSuccess) return err;
    if (entries_out && (err = cudaMemsetAsync(entries_out, 0, sizeof(uint32_t), stream)) != cudaSuccess) return err;
    const bool coarse = d.coarse && entries_out != nullptr && total_instances != nullptr && enumerate_in_smem(prm) && coarse_ratio() != 0;
    // device-side gate: entries <= pairs, so a step with fewer than (ratio - 1) instances per pair is unlikely to reach
    // `ratio` instances per entry
    if ((err = launch_tile_enumerate<false>(stream, prm, radii, xy, nullptr, d.counts, d.view_totals, nullptr, nullptr,
                                            coarse ? d.super_counts : nullptr, d.view_entries, total_instances,
                                            coarse_ratio() - (coarse_ratio() != 0))) != cudaSuccess)
        return err;
    tile_ranges_scan_kernel<<<prm.n_views, kBlock, 0, stream>>>(
        d.counts, d.view_totals, prm.n_tiles, (uint32_t)prm.n_views * (uint32_t)prm.n_tiles, ranges, d.list, d.list_x, d.head, longest_out,
        coarse ? d.super_counts : nullptr, d.view_entries, (int)d.n_super, d.super_offsets, d.items, d.item_capacity, entries_out);
    return cudaGetLastError();
}

// Whether a step with these counts takes the coarse grouping: large footprints (instances per entry) — the regime in
// which the plain scatter's 8-byte stores spread over thousands of segments — and the entries fit the buffer.
bool direct_bin_use_coarse(const RenderParams& prm, uint64_t n_instances, uint64_t coarse_entries)
{
    const DirectScratch d = direct_scratch(prm, nullptr);
    if (!d.coarse || !enumerate_in_smem(prm) || coarse_entries == 0) return false;
    if (coarse_entries > (uint64_t)prm.n_views * prm.P * 4) return false;  // the work-item list was sized for this bound
    const uint64_t ratio = coarse_ratio();
    return ratio != 0 && n_instances >= ratio * coarse_entries;
}

// D3 + D4.  pairs: L x 8 B of workspace.  keys_sorted may be null (keys not wanted).  longest_tile: the value read back
// after launch_direct_bin_count (decides which size classes are launched).  entries != null: coarse grouping first
// (entry buffer of coarse_entries x 16 B), then the fine scatter from the grouped entries.
cudaError_t launch_direct_bin_sort(cudaStream_t stream, const RenderParams& prm, const int32_t* radii, const float2* xy,
                                   const float* depth, const uint2* ranges, void* pairs, uint32_t* vals_sorted,
                                   uint64_t* keys_sorted, void* scratch, uint32_t longest_tile, void* entries,
                                   uint32_t instances_per_entry)
{
    const uint32_t n_ranges = (uint32_t)prm.n_views * (uint32_t)prm.n_tiles;
    const DirectScratch d = direct_scratch(prm, scratch);
    // lgm_set_tuning "sort_bulk": 2 (default) = the grouped-keys form (reads the segment from global memory, no staging);
    // 1 = first form, segment staged by one 1-D bulk copy (TMA) per tile; 0 = first form, staged by a load / store loop
    const int form = tuning(kTuneSortBulk) >= 0 ? tuning(kTuneSortBulk) : 2;
    const bool bulk = form != 0;
    using SortFn = void (*)(const uint2*, const uint2*, const uint32_t*, int, const uint32_t*, uint32_t*, uint32_t*, uint64_t*);
    SortFn sort_m, sort_x, sort_l, sort_s;
    size_t smem_m, smem_x, smem_l, smem_s;
    if (form >= 2) {
        sort_m = tile_group_sort_kernel<kSortThreadsM, kSortCapM, kLgBucketsM, 4>;
        sort_x = tile_group_sort_kernel<kSortThreadsX, kSortCapX, kLgBucketsX, 2>;
        sort_l = tile_group_sort_kernel<kSortThreadsL, kSortCapL, kLgBucketsL, 1>;
        sort_s = tile_group_sort_kernel<kSortThreadsS, kSortCapS, kLgBucketsS, 8>;
        smem_m = group_sort_smem(kSortCapM, kLgBucketsM); smem_x = group_sort_smem(kSortCapX, kLgBucketsX);
        smem_l = group_sort_smem(kSortCapL, kLgBucketsL); smem_s = group_sort_smem(kSortCapS, kLgBucketsS);
    } else {
        sort_m = bulk ? tile_bucket_sort_kernel<kSortThreadsM, kSortCapM, kLgBucketsM, 3, true>
                      : tile_bucket_sort_kernel<kSortThreadsM, kSortCapM, kLgBucketsM, 3, false>;
        sort_x = bulk ? tile_bucket_sort_kernel<kSortThreadsX, kSortCapX, kLgBucketsX, 2, true>
                      : tile_bucket_sort_kernel<kSortThreadsX, kSortCapX, kLgBucketsX, 2, false>;
        sort_l = bulk ? tile_bucket_sort_kernel<kSortThreadsL, kSortCapL, kLgBucketsL, 1, true>
                      : tile_bucket_sort_kernel<kSortThreadsL, kSortCapL, kLgBucketsL, 1, false>;
        sort_s = bulk ? tile_bucket_sort_kernel<kSortThreadsS, kSortCapS, kLgBucketsS, 6, true>
                      : tile_bucket_sort_kernel<kSortThreadsS, kSortCapS, kLgBucketsS, 6, false>;
        smem_m = sort_smem(kSortCapM, kLgBucketsM); smem_x = sort_smem(kSortCapX, kLgBucketsX);
        smem_l = sort_smem(kSortCapL, kLgBucketsL); smem_s = sort_smem(kSortCapS, kLgBucketsS);
    }
    static std::atomic<uint64_t> opted[3][4];
    if (cudaError_t e = opt_in_dynamic_smem(sort_m, smem_m, opted[form >= 2 ? 2 : form][0])) return e;
    if (cudaError_t e = opt_in_dynamic_smem(sort_x, smem_x, opted[form >= 2 ? 2 : form][1])) return e;
    if (cudaError_t e = opt_in_dynamic_smem(sort_l, smem_l, opted[form >= 2 ? 2 : form][2])) return e;
    if (cudaError_t e = opt_in_dynamic_smem(sort_s, smem_s, opted[form >= 2 ? 2 : form][3])) return e;
    const int n_sm = device_sm_count();
    cudaError_t err;
    if (entries) {
        dim3 grid(enum_ctas_per_view(prm), prm.n_views);
        coarse_scatter_kernel<<<grid, kBlock, (size_t)d.n_super * 8, stream>>>(prm, radii, xy, depth, d.super_offsets, d.super_cursor,
                                                                               static_cast<uint4*>(entries));
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
        tile_scatter_entries_kernel<<<4 * n_sm, kBlock, 0, stream>>>(prm, static_cast<const uint4*>(entries), d.super_offsets,
                                                                     d.super_counts, d.items, d.head, d.head + kHeadItemCursor, d.counts,
                                                                     ranges, static_cast<uint2*>(pairs),
                                                                     tuning(kTuneFineTileMajor) >= 0 ? tuning(kTuneFineTileMajor)
                                                                                                     : (instances_per_entry >= kTileMajorRatio));
        err = cudaGetLastError();
    } else {
        err = launch_tile_enumerate<true>(stream, prm, radii, xy, depth, d.counts, d.view_totals, ranges, static_cast<uint2*>(pairs));
    }
    if (err != cudaSuccess) return err;
    // the long tiles first: they are the critical path of the tail
    if (longest_tile > (uint32_t)kSortCapX) {
        const uint32_t n_cta = (uint32_t)min((unsigned)n_sm, n_ranges);
        sort_l<<<n_cta, kSortThreadsL, smem_l, stream>>>(static_cast<const uint2*>(pairs), ranges, d.list + (n_ranges - 1), -1,
                                                        d.head + kHeadL, d.head + kHeadLCursor, vals_sorted, keys_sorted);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    if (longest_tile > (uint32_t)kSortCapM) {
        const uint32_t n_cta = (uint32_t)min((unsigned)(2 * n_sm), n_ranges);
        sort_x<<<n_cta, kSortThreadsX, smem_x, stream>>>(static_cast<const uint2*>(pairs), ranges, d.list_x, 1, d.head + kHeadX,
                                                        d.head + kHeadXCursor, vals_sorted, keys_sorted);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    if (longest_tile > (uint32_t)kSortCapS) {
        const uint32_t n_cta = (uint32_t)min((unsigned)((form >= 2 ? 4 : 3) * n_sm), n_ranges);
        sort_m<<<n_cta, kSortThreadsM, smem_m, stream>>>(static_cast<const uint2*>(pairs), ranges, d.list, 1, d.head + kHeadM,
                                                        d.head + kHeadMCursor, vals_sorted, keys_sorted);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    const uint32_t n_cta = (uint32_t)min((unsigned)((form >= 2 ? 8 : 6) * n_sm), n_ranges);
    sort_s<<<n_cta, kSortThreadsS, smem_s, stream>>>(static_cast<const uint2*>(pairs), ranges, d.list_x + (n_ranges - 1), -1,
                                                    d.head + kHeadS, d.head + kHeadSCursor, vals_sorted, keys_sorted);
    return cudaGetLastError();
}

}  // namespace lgm
