// radix_sort.cu — K3: hand-written onesweep LSD radix sort of (u64 key, u32 value) pairs (stable), 8-bit digits.
// Replaces cub::DeviceRadixSort::SortPairs of the external rasterizer for the (view|tile|depth)
// keys of ALL views of a step in one sort.
//
// Structure (Adinets & Merrill "Onesweep", restated from the published algorithm, written from scratch):
//   1. one histogram kernel reads the keys once and builds the digit histograms of every pass;
//   2. per pass one kernel: a tile of 4096 pairs per CTA, dynamic tile ids (atomic ticket) so a CTA only ever waits
//      on CTAs that started before it; in-CTA ranking with warp match.any multisplit (stable), chained-scan
//      decoupled look-back across tiles per digit (flag+count packed in one 32-bit word, so no fences), staging
//      through shared memory so that global writes are contiguous per digit run.
// HBM traffic per pass = 12 B read + 12 B write per pair (+ 8 B per pair once for the histogram): the
// B_sort = n_pass*24*L + 8*L of SURVEY.md §8d.  No tensor cores: integer/byte work, HBM-bound.
#include "common.cuh"

namespace lgm {
namespace {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItems = 16;                           // pairs per thread
constexpr int kTileItems = kSortThreads * kItems;    // 4096 pairs per CTA
constexpr int kWarpItems = kItems * 32;
constexpr int kMaxPasses = 8;

constexpr uint32_t kFlagAgg = 1u << 30;   // tile aggregate available
constexpr uint32_t kFlagInc = 2u << 30;   // inclusive prefix available
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValMask = ~kFlagMask;

__global__ void __launch_bounds__(kSortThreads)
histogram_kernel(const uint64_t* __restrict__ keys, uint32_t n, int n_pass, int end_bit, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_hist[kMaxPasses * kRadix];
    for (int i = threadIdx.x; i < n_pass * kRadix; i += kSortThreads) s_hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // whole warps iterate together (match.any needs converged lanes); invalid lanes use an impossible marker
    const uint32_t stride = gridDim.x * kSortThreads;
    const uint32_t n_round = (n + stride - 1) / stride;
    for (uint32_t r = 0; r < n_round; r++) {
        const uint32_t i = r * stride + blockIdx.x * kSortThreads + threadIdx.x;
        const bool valid = i < n;
        const uint64_t k = valid ? keys[i] : 0ull;
        for (int p = 0; p < n_pass; p++) {
            // the last pass may cover fewer than 8 significant bits: bits at and above end_bit are ignored
            const uint32_t dmask = (1u << min(kRadixBits, end_bit - p * kRadixBits)) - 1u;
            const uint32_t d = valid ? (uint32_t)(k >> (p * kRadixBits)) & dmask : 0x100u + lane;
            const uint32_t m = __match_any_sync(0xffffffffu, d);
            if (valid && lane == (__ffs(m) - 1)) atomicAdd(&s_hist[p * kRadix + d], (uint32_t)__popc(m));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pass * kRadix; i += kSortThreads) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

struct OnesweepSmem {
    uint64_t keys[kTileItems];
    uint32_t vals[kTileItems];
    uint32_t whist[kSortWarps * kRadix];  // per-warp digit counts, then per-warp exclusive offsets
    uint32_t bin_start[kRadix];           // first slot of each digit inside the CTA's staged tile
    uint32_t goff[kRadix];                // global destination of slot j of digit d = goff[d] + j   (mod 2^32)
    uint32_t warp_tot[kSortWarps];
    uint32_t tile;
};

__global__ void __launch_bounds__(kSortThreads, 2)
onesweep_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                uint32_t dmask /* (1 << significant bits of this digit) - 1 */, const uint32_t* __restrict__ hist /*[256] of this pass*/, uint32_t* lookback /*[tiles][256], zeroed*/,
                uint32_t* ticket /*zeroed*/)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OnesweepSmem& sm = *reinterpret_cast<OnesweepSmem*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    if (t == 0) sm.tile = atomicAdd(ticket, 1u);
    for (int i = t; i < kSortWarps * kRadix; i += kSortThreads) sm.whist[i] = 0;
    __syncthreads();
    const uint32_t tile = sm.tile;
    const uint32_t base = tile * (uint32_t)kTileItems;
    const uint32_t n_valid = min((uint32_t)kTileItems, n - base);

    // ---- load (warp-striped: item i of lane l of warp w = base + w*512 + i*32 + l); pads (all-ones keys) have the
    // largest digit (dmask) in every pass and, being last in load order, rank after every real key of that digit ----
    uint64_t k[kItems];
    uint32_t v[kItems];
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const uint32_t loc = warp * kWarpItems + i * 32 + lane;
        const bool ok = loc < n_valid;
        k[i] = ok ? keys_in[base + loc] : ~0ull;
        v[i] = ok ? vals_in[base + loc] : 0u;
    }

    // ---- rank inside the warp: stable in (item, lane) order ----
    uint32_t rank[kItems];
    uint32_t* wh = sm.whist + warp * kRadix;
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const uint32_t d = (uint32_t)(k[i] >> shift) & dmask;
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(m) - 1;
        uint32_t pre = 0;
        if (lane == leader) {
            pre = wh[d];
            wh[d] = pre + (uint32_t)__popc(m);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[i] = pre + (uint32_t)__popc(m & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // ---- thread d: exclusive prefix of digit d over the warps; CTA count of digit d ----
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) {
        const uint32_t c = sm.whist[w * kRadix + t];
        sm.whist[w * kRadix + t] = cnt;
        cnt += c;
    }
    const uint32_t cnt_valid = ((uint32_t)t == dmask) ? cnt - ((uint32_t)kTileItems - n_valid) : cnt;

    // ---- publish the aggregate as early as possible, then the CTA-local scans ----
    volatile uint32_t* lb = lookback;
    if (tile != 0) lb[(size_t)tile * kRadix + t] = kFlagAgg | cnt_valid;

    // exclusive scan over digits of (a) the CTA counts -> bin_start, (b) the global histogram -> global bin base
    const uint32_t h = hist[t];
    const uint32_t incl_c = warp_incl_scan(cnt, lane);
    const uint32_t incl_h = warp_incl_scan(h, lane);
    if (lane == 31) {
        sm.warp_tot[warp] = incl_c;
        sm.goff[warp] = incl_h;  // temporary use of goff[0..7] as warp totals of the histogram scan
    }
    __syncthreads();
    uint32_t base_c = 0, base_h = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) {
        if (w < warp) {
            base_c += sm.warp_tot[w];
            base_h += sm.goff[w];
        }
    }
    const uint32_t bin_start = base_c + incl_c - cnt;
    const uint32_t bin_global = base_h + incl_h - h;
    __syncthreads();  // everyone has read goff[0..7] before it is overwritten below
    sm.bin_start[t] = bin_start;

    // ---- decoupled look-back for digit t ----
    uint32_t excl_prev = 0;
    if (tile == 0) {
        lb[t] = kFlagInc | cnt_valid;
    } else {
        int p = (int)tile - 1;
        while (true) {
            const uint32_t w = lb[(size_t)p * kRadix + t];
            if ((w & kFlagMask) == 0) continue;  // predecessor has not published yet (it started before us)
            excl_prev += w & kValMask;
            if (w & kFlagInc) break;
            p--;
        }
        lb[(size_t)tile * kRadix + t] = kFlagInc | (excl_prev + cnt_valid);
    }
    sm.goff[t] = bin_global + excl_prev - bin_start;
    __syncthreads();

    // ---- scatter into the staged tile (sorted by digit, stable) ----
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const uint32_t d = (uint32_t)(k[i] >> shift) & dmask;
        const uint32_t pos = sm.bin_start[d] + wh[d] + rank[i];
        sm.keys[pos] = k[i];
        sm.vals[pos] = v[i];
    }
    __syncthreads();

    // ---- write out: slot j -> goff[digit] + j ; consecutive slots of a digit are consecutive in global memory ----
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const uint32_t j = i * kSortThreads + t;
        if (j < n_valid) {
            const uint64_t key = sm.keys[j];
            const uint32_t d = (uint32_t)(key >> shift) & dmask;
            const uint32_t dst = sm.goff[d] + j;
            keys_out[dst] = key;
            vals_out[dst] = sm.vals[j];
        }
    }
}

}  // namespace

int sort_num_passes(int end_bit) { return (end_bit + kRadixBits - 1) / kRadixBits; }
// pass p reads buffer (p even ? start : other); the result of the last pass must be keys_out.
bool sort_input_is_tmp(int end_bit) { return (sort_num_passes(end_bit) & 1) != 0; }

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t sort_scratch_bytes(uint32_t n, int end_bit)
{
    const int np = sort_num_passes(end_bit);
    const size_t tiles = (n + kTileItems - 1) / kTileItems;
    // [hist: np*256 u32][tickets: np u32 (padded)][lookback: np * tiles * 256 u32]
    return align_up((size_t)np * kRadix * 4, 256) + 256 + (size_t)np * tiles * kRadix * 4;
}

cudaError_t launch_onesweep_sort(cudaStream_t stream, uint64_t* keys_out, uint32_t* vals_out, uint64_t* keys_tmp,
                                 uint32_t* vals_tmp, uint32_t n, int end_bit, void* scratch, size_t scratch_bytes)
{
    const int np = sort_num_passes(end_bit);
    if (np > kMaxPasses || np < 1) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    const size_t need = sort_scratch_bytes(n, end_bit);
    if (scratch_bytes < need) return cudaErrorInvalidValue;
    const uint32_t tiles = (n + kTileItems - 1) / kTileItems;
    unsigned char* sp = static_cast<unsigned char*>(scratch);
    uint32_t* hist = reinterpret_cast<uint32_t*>(sp);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(sp + align_up((size_t)np * kRadix * 4, 256));
    uint32_t* lookback = tickets + 64;
    cudaError_t err = cudaMemsetAsync(scratch, 0, need, stream);
    if (err != cudaSuccess) return err;

    uint64_t* kin = sort_input_is_tmp(end_bit) ? keys_tmp : keys_out;
    uint32_t* vin = sort_input_is_tmp(end_bit) ? vals_tmp : vals_out;
    uint64_t* kalt = sort_input_is_tmp(end_bit) ? keys_out : keys_tmp;
    uint32_t* valt = sort_input_is_tmp(end_bit) ? vals_out : vals_tmp;

    const int hist_grid = (int)min((size_t)148 * 8, (size_t)(n + kSortThreads * 8 - 1) / (kSortThreads * 8));
    histogram_kernel<<<hist_grid, kSortThreads, 0, stream>>>(kin, n, np, end_bit, hist);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;

    static bool attr_set = false;
    if (!attr_set) {
        err = cudaFuncSetAttribute(onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OnesweepSmem));
        if (err != cudaSuccess) return err;
        attr_set = true;
    }
    for (int p = 0; p < np; p++) {
        onesweep_kernel<<<tiles, kSortThreads, sizeof(OnesweepSmem), stream>>>(
            kin, vin, kalt, valt, n, p * kRadixBits, (1u << (end_bit - p * kRadixBits < kRadixBits ? end_bit - p * kRadixBits : kRadixBits)) - 1u,
            hist + p * kRadix, lookback + (size_t)p * tiles * kRadix, tickets + p);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        uint64_t* tk = kin; kin = kalt; kalt = tk;
        uint32_t* tv = vin; vin = valt; valt = tv;
    }
    return cudaSuccess;
}

}  // namespace lgm
