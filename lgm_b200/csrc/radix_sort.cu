// radix_sort.cu — K3: hand-written onesweep LSD radix sort of (u64 key, u32 value) pairs (stable), 8-bit digits.
// Replaces cub::DeviceRadixSort::SortPairs of the external rasterizer for the (view|tile|depth)
// keys of ALL views of a step in one sort.
//
// Structure (Adinets & Merrill "Onesweep", restated from the published algorithm, written from scratch):
//   1. one histogram kernel reads the keys once and builds the digit histograms of every pass;
//   2. per pass one kernel: a tile of pairs per CTA, dynamic tile ids (atomic ticket) so a CTA only ever waits
//      on CTAs that started before it; in-CTA ranking with warp match.any multisplit (stable), chained-scan
//      decoupled look-back across tiles per digit (flag+count packed in one 32-bit word, so no fences), staging
//      through shared memory so that global writes are contiguous per digit run.
// HBM traffic per pass = 12 B read + 12 B write per pair (+ 8 B per pair once for the histogram): the
// B_sort = n_pass*24*L + 8*L of SURVEY.md §8d.  No tensor cores: integer/byte work, HBM-bound.
//
// "Compressed" keys: the renderer's keys carry a float depth > 0 in bits [0,32), so bit 31 is always 0.  With
// compress = 1 the digits are taken from ck = key[30:0] | key[63:32] << 31 (computed on the fly; the stored key is
// untouched), which saves a whole pass whenever 31 + tile bits <= 8 k < 32 + tile bits (e.g. 208 views x 400 tiles:
// 48 bits -> 6 passes instead of 7).
#include "common.cuh"
#include <stdlib.h>

namespace lgm {
namespace {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 8;
constexpr int kHistThreads = 256;
constexpr int kHistItems = 8;

constexpr uint32_t kFlagAgg = 1u << 30;   // tile aggregate available
constexpr uint32_t kFlagInc = 2u << 30;   // inclusive prefix available
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValMask = ~kFlagMask;

__device__ __forceinline__ uint64_t sort_bits(uint64_t k, int compress)
{
    return compress ? ((k & 0x7fffffffull) | ((k >> 32) << 31)) : k;
}

// Digit histograms of all passes in one read of the keys.  Keys arrive in emit order, so the upper digits
// (view, tile row, depth exponent) are usually identical across a warp: a warp-uniform digit costs two REDUX and one
// shared-memory atomic; otherwise every lane adds 1 (distinct digits -> distinct banks, few conflicts).
__global__ void __launch_bounds__(kHistThreads)
histogram_kernel(const uint64_t* __restrict__ keys, uint32_t n, int n_pass, int begin_bit, int end_bit, int compress,
                 int n_low /* leading passes whose digits are lane-random (depth mantissa bytes) */,
                 uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_hist[kMaxPasses * kRadix];
    for (int i = threadIdx.x; i < n_pass * kRadix; i += kHistThreads) s_hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t chunk = kHistThreads * kHistItems;
    const uint32_t n_chunks = (n + chunk - 1) / chunk;
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        uint64_t k[kHistItems];
        bool ok[kHistItems];
#pragma unroll
        for (int j = 0; j < kHistItems; j++) {
            const uint32_t i = c * chunk + j * kHistThreads + threadIdx.x;
            ok[j] = i < n;
            k[j] = ok[j] ? sort_bits(keys[i], compress) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kHistItems; j++) {
            const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, ok[j]));
            if (nvalid == 0) continue;  // warp-uniform
            // bits at and above end_bit are ignored (the last pass may cover fewer than 8 significant bits)
            const uint64_t kk = (end_bit < 64 ? (k[j] & ((1ull << end_bit) - 1ull)) : k[j]) >> begin_bit;
            // passes 0..2 (low mantissa bytes of the depth): digits differ across lanes, plain shared atomics
            if (ok[j])
                for (int p = 0; p < n_low; p++) atomicAdd(&s_hist[p * kRadix + ((uint32_t)(kk >> (p * kRadixBits)) & 0xffu)], 1u);
            if (n_pass > n_low) {
                // passes 3.. : one uniformity test (two REDUX each on the two 32-bit halves of the upper bits) covers them all
                const uint64_t hi = kk >> (n_low * kRadixBits);
                const uint32_t h0 = (uint32_t)hi, h1 = (uint32_t)(hi >> 32);
                const bool uni = __reduce_min_sync(0xffffffffu, ok[j] ? h0 : 0xffffffffu) == __reduce_max_sync(0xffffffffu, ok[j] ? h0 : 0u) &&
                                 __reduce_min_sync(0xffffffffu, ok[j] ? h1 : 0xffffffffu) == __reduce_max_sync(0xffffffffu, ok[j] ? h1 : 0u);
                const uint64_t hv = __shfl_sync(0xffffffffu, hi, __ffs(__ballot_sync(0xffffffffu, ok[j])) - 1);
                if (uni) {
                    if (lane == 0)
                        for (int p = n_low; p < n_pass; p++) atomicAdd(&s_hist[p * kRadix + ((uint32_t)(hv >> ((p - n_low) * kRadixBits)) & 0xffu)], nvalid);
                } else if (ok[j]) {
                    for (int p = n_low; p < n_pass; p++) atomicAdd(&s_hist[p * kRadix + ((uint32_t)(hi >> ((p - n_low) * kRadixBits)) & 0xffu)], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pass * kRadix; i += kHistThreads) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

template <int THREADS, int ITEMS>
struct OnesweepSmem {
    static constexpr int kTileItems = THREADS * ITEMS;
    static constexpr int kWarps = THREADS / 32;
    uint64_t keys[kTileItems];
    uint32_t vals[kTileItems];
    uint32_t whist[kWarps * kRadix];  // per-warp digit counts, then per-warp exclusive offsets
    uint32_t bin_start[kRadix];       // first slot of each digit inside the CTA's staged tile
    uint32_t goff[kRadix];            // global destination of slot j of digit d = goff[d] + j   (mod 2^32)
    uint32_t warp_tot[8], warp_hist_tot[8];
    uint32_t tile;
};

// Decoupled look-back for one digit: sum of the counts of tiles [0, tile) of digit t.  Walks back over the
// predecessors' status words, adding aggregates until an inclusive prefix is met.  The walk is the latency chain of the
// whole pass (hundreds of tiles are in flight and most have only their aggregate out), so kLookAhead status words are
// fetched per round trip instead of one; a word that is not published yet is simply polled again.
constexpr int kLookAhead = 16;  // status words per round trip in the walk
constexpr int kLookFirst = 8;   // status words of the first round trip, issued before the scatter
template <int N>
__device__ __forceinline__ void lookback_issue(volatile uint32_t* lb, int p, int t, uint32_t (&w)[N])
{
#pragma unroll
    for (int u = 0; u < N; u++) {
        w[u] = uint32_t(kFlagInc);  // before tile 0: an inclusive prefix of zero
        if (p - u >= 0) w[u] = lb[(size_t)(p - u) * kRadix + t];
    }
}
// consume status words of tiles p, p-1, ...: returns true when an inclusive prefix ended the walk; a word that is not
// published yet stops the batch (that tile is polled again)
template <int N>
__device__ __forceinline__ bool lookback_consume(const uint32_t (&w)[N], int& p, uint32_t& excl)
{
#pragma unroll
    for (int u = 0; u < N; u++) {
        if ((w[u] & kFlagMask) == 0) return false;
        excl += w[u] & kValMask;
        p--;
        if (w[u] & kFlagInc) return true;
    }
    return false;
}
__device__ __forceinline__ uint32_t lookback_finish(volatile uint32_t* lb, uint32_t tile, int t, const uint32_t (&w0)[kLookFirst])
{
    uint32_t excl = 0;
    int p = (int)tile - 1;
    if (lookback_consume<kLookFirst>(w0, p, excl)) return excl;
    while (p >= 0) {
        uint32_t w[kLookAhead];
        lookback_issue<kLookAhead>(lb, p, t, w);
        if (lookback_consume<kLookAhead>(w, p, excl)) break;
    }
    return excl;
}
__device__ __forceinline__ uint32_t lookback_exclusive(volatile uint32_t* lb, uint32_t tile, int t)
{
    uint32_t excl = 0;
    int p = (int)tile - 1;
    while (p >= 0) {
        uint32_t w[kLookAhead];
        lookback_issue<kLookAhead>(lb, p, t, w);
        if (lookback_consume<kLookAhead>(w, p, excl)) break;
    }
    return excl;
}

// lanes of the warp holding the same digit as this lane.  MATCH = the match.any instruction; otherwise one ballot per
// digit bit (8 VOTEs + logic: more instructions, but no dependence on the long-latency MATCH unit).
template <bool MATCH>
__device__ __forceinline__ uint32_t peers_with_same_digit(uint32_t d)
{
    if (MATCH) return __match_any_sync(0xffffffffu, d);
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < kRadixBits; b++) {
        const bool bit = (d >> b) & 1u;
        const uint32_t vote = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? vote : ~vote;
    }
    return peers;
}

template <int THREADS, int ITEMS, int MIN_BLOCKS, bool MATCH>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                uint32_t dmask /* (1 << significant bits of this digit) - 1 */, int compress,
                const uint32_t* __restrict__ hist /*[256] of this pass*/, uint32_t* lookback /*[tiles][256], zeroed*/,
                uint32_t* ticket /*zeroed*/)
{
    using Smem = OnesweepSmem<THREADS, ITEMS>;
    constexpr int kTileItems = Smem::kTileItems;
    constexpr int kWarps = Smem::kWarps;
    constexpr int kWarpItems = ITEMS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    if (t == 0) sm.tile = atomicAdd(ticket, 1u);
    const uint32_t h_digit = t < kRadix ? hist[t] : 0u;  // global count of digit t in this pass (needed after ranking)
    for (int i = t; i < kWarps * kRadix; i += THREADS) sm.whist[i] = 0;
    __syncthreads();
    const uint32_t tile = sm.tile;
    const uint32_t base = tile * (uint32_t)kTileItems;
    const uint32_t n_valid = min((uint32_t)kTileItems, n - base);

    // ---- load (warp-striped: item i of lane l of warp w = base + w*ITEMS*32 + i*32 + l); pads (all-ones keys) have
    // the largest digit (dmask) in every pass and, being last in load order, rank after every real key of that digit
    uint64_t k[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t loc = warp * kWarpItems + i * 32 + lane;
        k[i] = loc < n_valid ? keys_in[base + loc] : ~0ull;
    }

    // ---- rank inside the warp: stable in (item, lane) order ----
    uint32_t rank[ITEMS];
    uint32_t* wh = sm.whist + warp * kRadix;
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t d = (uint32_t)(sort_bits(k[i], compress) >> shift) & dmask;
        const uint32_t m = peers_with_same_digit<MATCH>(d);
        const int leader = __ffs(m) - 1;
        // the group's leader reserves the group's slots with one shared-memory atomic (returns the digit's count over
        // the warp's earlier items); atomics of one warp execute in program order, so the ITEMS rounds pipeline with
        // no warp barrier between them
        uint32_t pre = 0;
        if (lane == leader) pre = atomicAdd(&wh[d], (uint32_t)__popc(m));
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[i] = pre + (uint32_t)__popc(m & lt_mask);
    }
    __syncthreads();

    // ---- threads 0..255: exclusive prefix of digit t over the warps; CTA count -> aggregate out; local bin scan ----
    volatile uint32_t* lb = lookback;
    uint32_t cnt_valid = 0, bin_start = 0, bin_global = 0;
    if (t < kRadix) {
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            const uint32_t c = sm.whist[w * kRadix + t];
            sm.whist[w * kRadix + t] = cnt;
            cnt += c;
        }
        cnt_valid = ((uint32_t)t == dmask) ? cnt - ((uint32_t)kTileItems - n_valid) : cnt;
        // publish the aggregate as early as possible
        if (tile != 0) lb[(size_t)tile * kRadix + t] = kFlagAgg | cnt_valid;

        // exclusive scan over digits of (a) the CTA counts -> bin_start, (b) the global histogram -> global bin base
        const uint32_t incl_c = warp_incl_scan(cnt, lane);
        const uint32_t incl_h = warp_incl_scan(h_digit, lane);
        if (lane == 31) {
            sm.warp_tot[warp] = incl_c;
            sm.warp_hist_tot[warp] = incl_h;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the first 8 warps only
        uint32_t base_c = 0, base_h = 0;
#pragma unroll
        for (int w = 0; w < kRadix / 32; w++) {
            if (w < warp) {
                base_c += sm.warp_tot[w];
                base_h += sm.warp_hist_tot[w];
            }
        }
        bin_start = base_c + incl_c - cnt;
        bin_global = base_h + incl_h - h_digit;
        sm.bin_start[t] = bin_start;
    }
    // values are loaded only now: their registers are not live during ranking
    uint32_t v[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t loc = warp * kWarpItems + i * 32 + lane;
        v[i] = loc < n_valid ? vals_in[base + loc] : 0u;
    }
    __syncthreads();

    // ---- look-back, first round trip in flight while the tile is scattered into shared memory ----
    uint32_t w0[kLookFirst];
    if (t < kRadix && tile != 0) lookback_issue<kLookFirst>(lb, (int)tile - 1, t, w0);

    // scatter into the staged tile (sorted by digit, stable)
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t d = (uint32_t)(sort_bits(k[i], compress) >> shift) & dmask;
        const uint32_t pos = sm.bin_start[d] + wh[d] + rank[i];
        sm.keys[pos] = k[i];
        sm.vals[pos] = v[i];
    }

    if (t < kRadix) {
        uint32_t excl_prev = 0;
        if (tile == 0) {
            lb[t] = kFlagInc | cnt_valid;
        } else {
            excl_prev = lookback_finish(lb, tile, t, w0);
            lb[(size_t)tile * kRadix + t] = kFlagInc | (excl_prev + cnt_valid);
        }
        sm.goff[t] = bin_global + excl_prev - bin_start;
    }
    __syncthreads();

    // ---- write out: slot j -> goff[digit] + j ; consecutive slots of a digit are consecutive in global memory ----
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t j = i * THREADS + t;
        if (j < n_valid) {
            const uint64_t key = sm.keys[j];
            const uint32_t d = (uint32_t)(sort_bits(key, compress) >> shift) & dmask;
            const uint32_t dst = sm.goff[d] + j;
            keys_out[dst] = key;
            vals_out[dst] = sm.vals[j];
        }
    }
}

// Persistent, software-pipelined form of the same pass: one CTA per resident slot loops over ticketed tiles and issues
// the key loads of its NEXT tile before ranking the current one, so HBM reads stay in flight during ranking, look-back
// and scatter (the one-tile-per-CTA form above keeps the memory system busy only ~1/4 of a CTA's life: measured 2.5 TB/s).
// Dead-lock freedom is unchanged: a CTA works on its tickets in increasing order and a tile only ever waits on lower
// tickets, each of which is held by a resident CTA that needs nothing from higher tickets.
template <int THREADS, int ITEMS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_persistent_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                           uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, uint32_t n_tiles,
                           int shift, uint32_t dmask, int compress, const uint32_t* __restrict__ hist,
                           uint32_t* lookback /*[tiles][256], zeroed*/, uint32_t* ticket /*zeroed*/)
{
    using Smem = OnesweepSmem<THREADS, ITEMS>;
    constexpr int kTileItems = Smem::kTileItems;
    constexpr int kWarps = Smem::kWarps;
    constexpr int kWarpItems = ITEMS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    __shared__ uint32_t s_next;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* wh = sm.whist + warp * kRadix;
    volatile uint32_t* lb = lookback;

    // global bin bases of this pass (exclusive scan of the histogram), kept in a register by threads 0..255
    uint32_t bin_global = 0;
    if (t == 0) sm.tile = atomicAdd(ticket, 1u);
    if (t < kRadix) {
        const uint32_t h = hist[t];
        const uint32_t incl_h = warp_incl_scan(h, lane);
        if (lane == 31) sm.warp_hist_tot[warp] = incl_h;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t base_h = 0;
#pragma unroll
        for (int w = 0; w < kRadix / 32; w++)
            if (w < warp) base_h += sm.warp_hist_tot[w];
        bin_global = base_h + incl_h - h;
    }
    __syncthreads();
    uint32_t tile = sm.tile;

    uint64_t kn[ITEMS];  // keys of the tile this CTA will process next (in flight)
    if (tile < n_tiles) {
        const uint32_t base = tile * (uint32_t)kTileItems;
        const uint32_t n_valid = min((uint32_t)kTileItems, n - base);
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t loc = warp * kWarpItems + i * 32 + lane;
            kn[i] = loc < n_valid ? keys_in[base + loc] : ~0ull;
        }
    }

    while (tile < n_tiles) {
        const uint32_t base = tile * (uint32_t)kTileItems;
        const uint32_t n_valid = min((uint32_t)kTileItems, n - base);
        uint64_t k[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) k[i] = kn[i];

        __syncthreads();  // everyone is done with the previous tile's shared state
        uint32_t my_next = 0;
        if (t == 0) my_next = atomicAdd(ticket, 1u);  // stays in a register: the round trip overlaps the ranking
        for (int i = t; i < kWarps * kRadix; i += THREADS) sm.whist[i] = 0;
        __syncthreads();

        // values of this tile: needed only at the scatter
        uint32_t v[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t loc = warp * kWarpItems + i * 32 + lane;
            v[i] = loc < n_valid ? vals_in[base + loc] : 0u;
        }

        // ---- rank inside the warp: stable in (item, lane) order ----
        uint32_t rank[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t d = (uint32_t)(sort_bits(k[i], compress) >> shift) & dmask;
            const uint32_t m = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(m) - 1;
            uint32_t pre = 0;
            if (lane == leader) pre = atomicAdd(&wh[d], (uint32_t)__popc(m));
            pre = __shfl_sync(0xffffffffu, pre, leader);
            rank[i] = pre + (uint32_t)__popc(m & lt_mask);
        }
        if (t == 0) s_next = my_next;
        __syncthreads();

        // ---- prefetch the next tile's keys (ticket has arrived by now) ----
        const uint32_t next = s_next;
        if (next < n_tiles) {
            const uint32_t nbase = next * (uint32_t)kTileItems;
            const uint32_t nn = min((uint32_t)kTileItems, n - nbase);
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const uint32_t loc = warp * kWarpItems + i * 32 + lane;
                kn[i] = loc < nn ? keys_in[nbase + loc] : ~0ull;
            }
        }

        // ---- threads 0..255: per-digit prefix over warps, CTA count, look-back, offsets ----
        if (t < kRadix) {
            uint32_t cnt = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) {
                const uint32_t c = sm.whist[w * kRadix + t];
                sm.whist[w * kRadix + t] = cnt;
                cnt += c;
            }
            const uint32_t cnt_valid = ((uint32_t)t == dmask) ? cnt - ((uint32_t)kTileItems - n_valid) : cnt;
            if (tile != 0) lb[(size_t)tile * kRadix + t] = kFlagAgg | cnt_valid;
            const uint32_t incl_c = warp_incl_scan(cnt, lane);
            if (lane == 31) sm.warp_tot[warp] = incl_c;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            uint32_t base_c = 0;
#pragma unroll
            for (int w = 0; w < kRadix / 32; w++)
                if (w < warp) base_c += sm.warp_tot[w];
            const uint32_t bin_start = base_c + incl_c - cnt;
            sm.bin_start[t] = bin_start;
            uint32_t excl_prev = 0;
            if (tile == 0) {
                lb[t] = kFlagInc | cnt_valid;
            } else {
                excl_prev = lookback_exclusive(lb, tile, t);
                lb[(size_t)tile * kRadix + t] = kFlagInc | (excl_prev + cnt_valid);
            }
            sm.goff[t] = bin_global + excl_prev - bin_start;
        }
        __syncthreads();

        // ---- scatter into the staged tile (sorted by digit, stable), then write out ----
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t d = (uint32_t)(sort_bits(k[i], compress) >> shift) & dmask;
            const uint32_t pos = sm.bin_start[d] + wh[d] + rank[i];
            sm.keys[pos] = k[i];
            sm.vals[pos] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t j = i * THREADS + t;
            if (j < n_valid) {
                const uint64_t key = sm.keys[j];
                const uint32_t d = (uint32_t)(sort_bits(key, compress) >> shift) & dmask;
                const uint32_t dst = sm.goff[d] + j;
                keys_out[dst] = key;
                vals_out[dst] = sm.vals[j];
            }
        }
        tile = next;
    }
}

struct Variant {
    int threads, items, min_blocks;
    bool match, persistent;
};
// Tunable launch shapes (LGM_SORT_VARIANT selects; default chosen from B200 measurements)
constexpr Variant kVariants[] = {{256, 16, 2, true, false}, {512, 8, 2, true, false},  {256, 8, 4, true, false},
                                 {256, 16, 2, false, false}, {512, 8, 2, false, false}, {256, 8, 4, false, false},
                                 {384, 8, 2, true, true},   {384, 8, 3, false, false}, {384, 8, 3, true, false},
                                 {256, 12, 3, false, false}, {384, 6, 4, false, false}, {1024, 8, 1, false, false}};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kDefaultVariant = 4;

int variant_index()
{
    const int v = tuning(kTuneSortVariant);  // lgm_set_tuning "sort_variant"
    return (v >= 0 && v < kNumVariants) ? v : kDefaultVariant;
}

template <int THREADS, int ITEMS, int MIN_BLOCKS, bool MATCH>
cudaError_t launch_pass(cudaStream_t stream, uint32_t tiles, const uint64_t* kin, const uint32_t* vin, uint64_t* kout,
                        uint32_t* vout, uint32_t n, int shift, uint32_t dmask, int compress, const uint32_t* hist,
                        uint32_t* lookback, uint32_t* ticket)
{
    using Smem = OnesweepSmem<THREADS, ITEMS>;
    auto kern = onesweep_kernel<THREADS, ITEMS, MIN_BLOCKS, MATCH>;
    static std::atomic<uint64_t> opted{0};
    if (cudaError_t e = opt_in_dynamic_smem(kern, sizeof(Smem), opted)) return e;
    kern<<<tiles, THREADS, sizeof(Smem), stream>>>(kin, vin, kout, vout, n, shift, dmask, compress, hist, lookback, ticket);
    return cudaGetLastError();
}

template <int THREADS, int ITEMS, int MIN_BLOCKS>
cudaError_t launch_pass_persistent(cudaStream_t stream, uint32_t tiles, const uint64_t* kin, const uint32_t* vin,
                                   uint64_t* kout, uint32_t* vout, uint32_t n, int shift, uint32_t dmask, int compress,
                                   const uint32_t* hist, uint32_t* lookback, uint32_t* ticket)
{
    using Smem = OnesweepSmem<THREADS, ITEMS>;
    auto kern = onesweep_persistent_kernel<THREADS, ITEMS, MIN_BLOCKS>;
    static std::atomic<uint64_t> opted{0};
    static std::atomic<int> per_sm_cached{0};  // occupancy of this kernel: the same on every B200 of the box
    if (cudaError_t e = opt_in_dynamic_smem(kern, sizeof(Smem), opted)) return e;
    int per_sm = per_sm_cached.load(std::memory_order_relaxed);
    if (!per_sm) {
        if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, sizeof(Smem))) return e;
        per_sm = per_sm > 0 ? per_sm : 1;
        per_sm_cached.store(per_sm, std::memory_order_relaxed);
    }
    const int resident = device_sm_count() * per_sm;  // one CTA per resident slot: a multiple of the SM count (148)
    const uint32_t grid = tiles < (uint32_t)resident ? tiles : (uint32_t)resident;
    kern<<<grid, THREADS, sizeof(Smem), stream>>>(kin, vin, kout, vout, n, tiles, shift, dmask, compress, hist, lookback, ticket);
    return cudaGetLastError();
}

}  // namespace

int sort_num_passes(int begin_bit, int end_bit) { return (end_bit - begin_bit + kRadixBits - 1) / kRadixBits; }
// pass p reads buffer (p even ? start : other); the result of the last pass must be keys_out.
bool sort_input_is_tmp(int begin_bit, int end_bit) { return (sort_num_passes(begin_bit, end_bit) & 1) != 0; }

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static uint32_t tile_items() { const Variant v = kVariants[variant_index()]; return (uint32_t)(v.threads * v.items); }

size_t sort_scratch_bytes(uint32_t n, int begin_bit, int end_bit)
{
    const int np = sort_num_passes(begin_bit, end_bit);
    const size_t tiles = (n + tile_items() - 1) / tile_items();
    // [hist: np*256 u32][tickets: np u32 (padded)][lookback: np * tiles * 256 u32]
    return align_up((size_t)np * kRadix * 4, 256) + 256 + (size_t)np * tiles * kRadix * 4;
}

cudaError_t launch_onesweep_sort(cudaStream_t stream, uint64_t* keys_out, uint32_t* vals_out, uint64_t* keys_tmp,
                                 uint32_t* vals_tmp, uint32_t n, int begin_bit, int end_bit, int compress, void* scratch,
                                 size_t scratch_bytes)
{
    const int np = sort_num_passes(begin_bit, end_bit);
    if (np > kMaxPasses || np < 1) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    const size_t need = sort_scratch_bytes(n, begin_bit, end_bit);
    if (scratch_bytes < need) return cudaErrorInvalidValue;
    const uint32_t tiles = (n + tile_items() - 1) / tile_items();
    unsigned char* sp = static_cast<unsigned char*>(scratch);
    uint32_t* hist = reinterpret_cast<uint32_t*>(sp);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(sp + align_up((size_t)np * kRadix * 4, 256));
    uint32_t* lookback = tickets + 64;
    cudaError_t err = cudaMemsetAsync(scratch, 0, need, stream);
    if (err != cudaSuccess) return err;

    const bool in_tmp = sort_input_is_tmp(begin_bit, end_bit);
    uint64_t* kin = in_tmp ? keys_tmp : keys_out;
    uint32_t* vin = in_tmp ? vals_tmp : vals_out;
    uint64_t* kalt = in_tmp ? keys_out : keys_tmp;
    uint32_t* valt = in_tmp ? vals_out : vals_tmp;

    const uint32_t chunk = kHistThreads * kHistItems;
    const int hist_grid = (int)min((size_t)148 * 8, (size_t)(n + chunk - 1) / chunk);
    // passes over the low 24 bits (depth mantissa) see lane-random digits; higher digits are near-uniform per warp
    int n_low = (24 - begin_bit + kRadixBits - 1) / kRadixBits;
    n_low = n_low < 0 ? 0 : (n_low > np ? np : n_low);
    histogram_kernel<<<hist_grid, kHistThreads, 0, stream>>>(kin, n, np, begin_bit, end_bit, compress, n_low, hist);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;

    const int vi = variant_index();
    for (int p = 0; p < np; p++) {
        const int shift = begin_bit + p * kRadixBits;
        const int sig = end_bit - shift < kRadixBits ? end_bit - shift : kRadixBits;
        const uint32_t dmask = (1u << sig) - 1u;
        const uint32_t* h = hist + p * kRadix;
        uint32_t* lb = lookback + (size_t)p * tiles * kRadix;
        uint32_t* tk = tickets + p;
#define LGM_PASS(T, I, B, M) launch_pass<T, I, B, M>(stream, tiles, kin, vin, kalt, valt, n, shift, dmask, compress, h, lb, tk)
        switch (vi) {
            case 0: err = LGM_PASS(256, 16, 2, true); break;
            case 1: err = LGM_PASS(512, 8, 2, true); break;
            case 2: err = LGM_PASS(256, 8, 4, true); break;
            case 3: err = LGM_PASS(256, 16, 2, false); break;
            case 4: err = LGM_PASS(512, 8, 2, false); break;
            case 5: err = LGM_PASS(256, 8, 4, false); break;
#define LGM_PPASS(T, I, B) launch_pass_persistent<T, I, B>(stream, tiles, kin, vin, kalt, valt, n, shift, dmask, compress, h, lb, tk)
            case 6: err = LGM_PPASS(384, 8, 2); break;
#undef LGM_PPASS
            case 7: err = LGM_PASS(384, 8, 3, false); break;
            case 8: err = LGM_PASS(384, 8, 3, true); break;
            case 9: err = LGM_PASS(256, 12, 3, false); break;
            case 10: err = LGM_PASS(384, 6, 4, false); break;
            default: err = LGM_PASS(1024, 8, 1, false); break;
        }
#undef LGM_PASS
        if (err != cudaSuccess) return err;
        uint64_t* tk2 = kin; kin = kalt; kalt = tk2;
        uint32_t* tv = vin; vin = valt; valt = tv;
    }
    return cudaSuccess;
}

}  // namespace lgm
