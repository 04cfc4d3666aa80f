// tile_sort.cu — second half of the hybrid (view|tile|depth) sort: after the onesweep passes have grouped the
// instances by global tile (stable, so each tile still holds its Gaussians in emit order), every tile's segment is
// sorted on the 31 depth bits in SHARED MEMORY by one CTA: up to 4 stable 8-bit LSD passes that never touch HBM.
//
// The result is bit-identical to sorting the whole 64-bit key with LSD onesweep passes (same stable order: tile,
// depth bits, emit order), but the depth bits cost 12 B read + 4..12 B write per instance instead of 4 x 24 B.
// A 16x16 tile holds ~750 instances in a trained scene; B200's 227 KB of shared memory per CTA also keeps the rare
// long tiles on chip:
//   short tiles  n <=  2048 : one CTA per tile (grid = all tiles), 256 threads, 40 KB, 5 CTAs per SM
//   long tiles   n <= 12288 : collected in a list by the short-tile launch; a persistent grid of one 512-thread,
//                             208 KB CTA per SM pulls them from the list
//   beyond that             : the same passes by that CTA through the global alternate buffers (rare, correct, slow)
// A pass whose digit is the same for the whole segment (the depth exponent byte, nearly always) is skipped.
#include "common.cuh"

namespace lgm {
namespace {

constexpr int kBits = 8;
constexpr int kBins = 1 << kBits;
constexpr int kSmallCap = 2048, kSmallThreads = 256;
constexpr int kLargeCap = 12288, kLargeThreads = 512;

// lanes of the warp holding the same digit as this lane (8 ballots; cf. radix_sort.cu)
__device__ __forceinline__ uint32_t same_digit_peers(uint32_t d, uint32_t active)
{
    uint32_t peers = active;
#pragma unroll
    for (int b = 0; b < kBits; b++) {
        const bool bit = (d >> b) & 1u;
        const uint32_t vote = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? vote : ~vote;
    }
    return peers;
}

// One stable 8-bit pass over n (key32, val32) pairs: src -> dst; returns false (and leaves dst untouched) when every
// key has the same digit.  Each warp owns a contiguous chunk of the segment (so "warp order, then position" is the
// original order); sweep 1 counts digits per warp, a per-digit prefix over the warps turns the counts into start
// offsets, sweep 2 re-derives the peer groups and lets each group's leader reserve its slots with one shared-memory
// atomic — the slots of a group are handed out in lane (= position) order.
template <int THREADS>
__device__ __forceinline__ bool block_radix_pass(const uint32_t* __restrict__ src_k, const uint32_t* __restrict__ src_v,
                                                 uint32_t* __restrict__ dst_k, uint32_t* __restrict__ dst_v, int n, int shift,
                                                 uint32_t dmask, uint32_t* whist /*[THREADS/32][256]*/, uint32_t* s_misc /*[9]*/)
{
    constexpr int kWarps = THREADS / 32;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int chunk = ((n + kWarps * 32 - 1) / (kWarps * 32)) * 32;  // per-warp chunk, multiple of 32
    const int w0 = warp * chunk, w1 = min(n, w0 + chunk);
    uint32_t* wh = whist + warp * kBins;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (int i = t; i < kWarps * kBins; i += THREADS) whist[i] = 0;
    if (t == 0) s_misc[8] = 0;
    __syncthreads();
    // sweep 1: per-warp digit counts
    for (int i = w0; i < w1; i += 32) {
        const int j = i + lane;
        const bool ok = j < w1;
        const uint32_t d = ok ? (src_k[j] >> shift) & dmask : 0u;
        const uint32_t act = __ballot_sync(0xffffffffu, ok);
        const uint32_t m = same_digit_peers(d, act);
        if (ok && lane == __ffs(m) - 1) atomicAdd(&wh[d], (uint32_t)__popc(m));
    }
    __syncthreads();
    // per-digit exclusive prefix over the warps, then exclusive scan over the digits (threads 0..255)
    if (t < kBins) {
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            const uint32_t c = whist[w * kBins + t];
            whist[w * kBins + t] = cnt;
            cnt += c;
        }
        if (cnt == (uint32_t)n) s_misc[8] = 1;  // the whole segment shares this digit: nothing to do
        const uint32_t incl = warp_incl_scan(cnt, lane);
        if (lane == 31) s_misc[warp] = incl;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t base = 0;
#pragma unroll
        for (int w = 0; w < kBins / 32; w++)
            if (w < warp) base += s_misc[w];
        const uint32_t bin_start = base + incl - cnt;
#pragma unroll
        for (int w = 0; w < kWarps; w++) whist[w * kBins + t] += bin_start;
    }
    __syncthreads();
    if (s_misc[8]) {
        __syncthreads();  // everyone has read the flag before the next pass resets it
        return false;
    }
    // sweep 2: rank and scatter
    for (int i = w0; i < w1; i += 32) {
        const int j = i + lane;
        const bool ok = j < w1;
        const uint32_t k = ok ? src_k[j] : 0u;
        const uint32_t v = ok ? src_v[j] : 0u;
        const uint32_t d = (k >> shift) & dmask;
        const uint32_t act = __ballot_sync(0xffffffffu, ok);
        const uint32_t m = same_digit_peers(d, act);
        const int leader = __ffs(m) - 1;
        uint32_t pos = 0;
        if (ok && lane == leader) pos = atomicAdd(&wh[d], (uint32_t)__popc(m));
        pos = __shfl_sync(0xffffffffu, pos, leader < 0 ? 0 : leader);
        if (ok) {
            pos += (uint32_t)__popc(m & lt_mask);
            dst_k[pos] = k;
            dst_v[pos] = v;
        }
    }
    __syncthreads();
    return true;
}

// Sort one segment.  n <= cap: in shared memory (planes a/b); else through the global alternate buffers.
template <int THREADS>
__device__ __forceinline__ void sort_segment(uint64_t* __restrict__ gk, uint32_t* __restrict__ gv, uint64_t* __restrict__ gk_tmp,
                                             uint32_t* __restrict__ gv_tmp, int n, int cap, uint32_t* whist, uint32_t* planes,
                                             uint32_t* s_misc, int write_keys)
{
    const int t = threadIdx.x;
    uint32_t *a_k, *a_v, *b_k, *b_v;
    if (n <= cap) {
        a_k = planes; a_v = a_k + cap; b_k = a_v + cap; b_v = b_k + cap;
        for (int i = t; i < n; i += THREADS) {
            a_k[i] = (uint32_t)gk[i];  // the low word: depth bits (the high word, the tile, is the same for all)
            a_v[i] = gv[i];
        }
    } else {
        // two u32 key planes carved from this segment's slice of the alternate key buffer (n u64 = 2n u32); the
        // values ping-pong between the value array itself and the alternate value buffer
        a_k = reinterpret_cast<uint32_t*>(gk_tmp); b_k = a_k + n; a_v = gv; b_v = gv_tmp;
        for (int i = t; i < n; i += THREADS) a_k[i] = (uint32_t)gk[i];
    }
    __syncthreads();
    const uint32_t masks[4] = {0xffu, 0xffu, 0xffu, 0x7fu};
#pragma unroll
    for (int p = 0; p < 4; p++) {
        if (block_radix_pass<THREADS>(a_k, a_v, b_k, b_v, n, 8 * p, masks[p], whist, s_misc)) {
            uint32_t* tk = a_k; a_k = b_k; b_k = tk;
            uint32_t* tv = a_v; a_v = b_v; b_v = tv;
        }
    }
    // the sorted segment is in (a_k, a_v)
    const uint64_t hi = gk[0] & 0xffffffff00000000ull;
    if (a_v != gv)
        for (int i = t; i < n; i += THREADS) gv[i] = a_v[i];
    if (write_keys) {
        __syncthreads();
        for (int i = t; i < n; i += THREADS) gk[i] = hi | a_k[i];
    }
    __syncthreads();
}

// short tiles: one CTA per tile; tiles longer than kSmallCap are appended to `long_list` (count in long_count[0])
__global__ void __launch_bounds__(kSmallThreads)
tile_sort_short_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, const uint2* __restrict__ ranges,
                       uint32_t* __restrict__ long_list, uint32_t* __restrict__ long_count, int write_keys)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_misc[9];
    const uint2 range = ranges[blockIdx.x];
    const int n = (int)(range.y - range.x);
    if (n <= 1) return;
    if (n > kSmallCap) {
        if (threadIdx.x == 0) long_list[atomicAdd(long_count, 1u)] = blockIdx.x;
        return;
    }
    uint32_t* whist = reinterpret_cast<uint32_t*>(smem_raw);
    sort_segment<kSmallThreads>(keys + range.x, vals + range.x, nullptr, nullptr, n, kSmallCap, whist,
                                whist + (kSmallThreads / 32) * kBins, s_misc, write_keys);
}

// long tiles: persistent CTAs (one per SM: 208 KB of shared memory each) pull tile ids from the list
__global__ void __launch_bounds__(kLargeThreads)
tile_sort_long_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t* __restrict__ keys_tmp,
                      uint32_t* __restrict__ vals_tmp, const uint2* __restrict__ ranges,
                      const uint32_t* __restrict__ long_list, const uint32_t* __restrict__ long_count,
                      uint32_t* __restrict__ cursor, int write_keys)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_misc[9];
    __shared__ uint32_t s_item;
    uint32_t* whist = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t count = *long_count;
    while (true) {
        if (threadIdx.x == 0) s_item = atomicAdd(cursor, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        __syncthreads();
        if (item >= count) break;
        const uint2 range = ranges[long_list[item]];
        const int n = (int)(range.y - range.x);
        sort_segment<kLargeThreads>(keys + range.x, vals + range.x, keys_tmp + range.x, vals_tmp + range.x, n, kLargeCap, whist,
                                    whist + (kLargeThreads / 32) * kBins, s_misc, write_keys);
    }
}

}  // namespace

size_t tile_sort_scratch_bytes(uint32_t n_ranges) { return ((size_t)n_ranges + 64) * sizeof(uint32_t); }

// scratch: [long_count, cursor, pad.. (64 u32)] [long_list: n_ranges u32]
cudaError_t launch_tile_depth_sort(cudaStream_t stream, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                                   const uint2* ranges, uint32_t n_ranges, int write_keys, void* scratch)
{
    if (n_ranges == 0) return cudaSuccess;
    constexpr size_t small_smem = (size_t)(kSmallThreads / 32) * kBins * 4 + (size_t)kSmallCap * 16;
    constexpr size_t large_smem = (size_t)(kLargeThreads / 32) * kBins * 4 + (size_t)kLargeCap * 16;
    static std::atomic<uint64_t> opted_s{0}, opted_l{0};
    if (cudaError_t e = opt_in_dynamic_smem(tile_sort_short_kernel, small_smem, opted_s)) return e;
    if (cudaError_t e = opt_in_dynamic_smem(tile_sort_long_kernel, large_smem, opted_l)) return e;
    const int n_sm = device_sm_count();
    uint32_t* head = static_cast<uint32_t*>(scratch);
    cudaError_t err = cudaMemsetAsync(head, 0, 64 * sizeof(uint32_t), stream);
    if (err != cudaSuccess) return err;
    tile_sort_short_kernel<<<n_ranges, kSmallThreads, small_smem, stream>>>(keys, vals, ranges, head + 64, head, write_keys);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    tile_sort_long_kernel<<<n_sm, kLargeThreads, large_smem, stream>>>(keys, vals, keys_tmp, vals_tmp, ranges, head + 64, head,
                                                                     head + 1, write_keys);
    return cudaGetLastError();
}

}  // namespace lgm
