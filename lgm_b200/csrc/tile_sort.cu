// tile_sort.cu — second half of the hybrid (view|tile|depth) sort: after the onesweep passes have grouped the
// instances by global tile (stable, so each tile still holds its Gaussians in emit order), every tile's segment is
// sorted on the 31 depth bits in SHARED MEMORY by one CTA: 4 stable 8-bit LSD passes that never touch HBM.
//
// The result is bit-identical to sorting the whole 64-bit key with LSD onesweep passes (same stable order: tile,
// depth bits, emit order), but the depth bits cost 12 B read + 4..12 B write per instance instead of 4 x 24 B.
// A 16x16 tile holds ~750 instances in a trained scene and ~10 k at initialisation; B200's 227 KB of shared memory
// per CTA keeps both kinds on chip:
//   small class  n <=  2048 : 256 threads,  40 KB shared memory, 5 CTAs per SM
//   large class  n <= 12288 : 512 threads, 208 KB shared memory, 1 CTA  per SM
//   beyond that             : the same passes by one CTA through the global alternate buffers (rare, correct, slow)
// Each class is one launch over all tiles; a CTA whose tile belongs to another class exits at once.
#include "common.cuh"

namespace lgm {
namespace {

constexpr int kBits = 8;
constexpr int kBins = 1 << kBits;
constexpr int kDepthBits = 31;  // a depth > 0.2 has a clear sign bit
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

// One stable 8-bit pass over n (key32, val32) pairs: src -> dst.  Each warp owns a contiguous chunk of the segment
// (so "warp order, then position" is the original order); sweep 1 counts digits per warp, a per-digit prefix over the
// warps turns the counts into start offsets, sweep 2 re-derives the peer groups and lets each group's leader reserve
// its slots with one shared-memory atomic — the slots of a group are handed out in lane (= position) order.
template <int THREADS>
__device__ __forceinline__ void block_radix_pass(const uint32_t* __restrict__ src_k, const uint32_t* __restrict__ src_v,
                                                 uint32_t* __restrict__ dst_k, uint32_t* __restrict__ dst_v, int n, int shift,
                                                 uint32_t dmask, uint32_t* whist /*[THREADS/32][256]*/, uint32_t* s_warp /*[8]*/)
{
    constexpr int kWarps = THREADS / 32;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int chunk = ((n + kWarps * 32 - 1) / (kWarps * 32)) * 32;  // per-warp chunk, multiple of 32
    const int w0 = warp * chunk, w1 = min(n, w0 + chunk);
    uint32_t* wh = whist + warp * kBins;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (int i = t; i < kWarps * kBins; i += THREADS) whist[i] = 0;
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
        const uint32_t incl = warp_incl_scan(cnt, lane);
        if (lane == 31) s_warp[warp] = incl;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t base = 0;
#pragma unroll
        for (int w = 0; w < kBins / 32; w++)
            if (w < warp) base += s_warp[w];
        const uint32_t bin_start = base + incl - cnt;
#pragma unroll
        for (int w = 0; w < kWarps; w++) whist[w * kBins + t] += bin_start;
    }
    __syncthreads();
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
}

// CAP > 0: tiles with lo < n <= CAP are sorted in shared memory.  CAP == 0 (only in the large-class launch): tiles
// longer than kLargeCap run the same passes through the global alternate buffers.
template <int THREADS, int CAP>
__global__ void __launch_bounds__(THREADS)
tile_depth_sort_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t* __restrict__ keys_tmp,
                       uint32_t* __restrict__ vals_tmp, const uint2* __restrict__ ranges, int lo, int write_keys)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kWarps = THREADS / 32;
    __shared__ uint32_t s_warp[8];
    const uint2 range = ranges[blockIdx.x];
    const int n = (int)(range.y - range.x);
    if (n <= lo || n <= 1) return;  // another class, or nothing to sort
    const int t = threadIdx.x;
    uint32_t* whist = reinterpret_cast<uint32_t*>(smem_raw);
    uint64_t* gk = keys + range.x;
    uint32_t* gv = vals + range.x;

    if (n <= CAP) {
        uint32_t* a_k = whist + kWarps * kBins;
        uint32_t* a_v = a_k + CAP;
        uint32_t* b_k = a_v + CAP;
        uint32_t* b_v = b_k + CAP;
        for (int i = t; i < n; i += THREADS) {
            a_k[i] = (uint32_t)gk[i];  // the low word: depth bits (the high word, the tile, is the same for all)
            a_v[i] = gv[i];
        }
        __syncthreads();
        block_radix_pass<THREADS>(a_k, a_v, b_k, b_v, n, 0, 0xffu, whist, s_warp);
        block_radix_pass<THREADS>(b_k, b_v, a_k, a_v, n, 8, 0xffu, whist, s_warp);
        block_radix_pass<THREADS>(a_k, a_v, b_k, b_v, n, 16, 0xffu, whist, s_warp);
        block_radix_pass<THREADS>(b_k, b_v, a_k, a_v, n, 24, 0x7fu, whist, s_warp);
        const uint64_t hi = gk[0] & 0xffffffff00000000ull;
        for (int i = t; i < n; i += THREADS) {
            gv[i] = a_v[i];
            if (write_keys) gk[i] = hi | a_k[i];
        }
    } else if (CAP >= kLargeCap) {
        // longer than shared memory allows: identical passes through global memory (two u32 planes carved from the
        // alternate key buffer of this segment, values ping-pong with the alternate value buffer)
        uint32_t* p_k = reinterpret_cast<uint32_t*>(keys_tmp + range.x);  // 2n u32 available: planes A and B
        uint32_t* q_k = p_k + n;
        uint32_t* q_v = vals_tmp + range.x;
        for (int i = t; i < n; i += THREADS) p_k[i] = (uint32_t)gk[i];
        __syncthreads();
        block_radix_pass<THREADS>(p_k, gv, q_k, q_v, n, 0, 0xffu, whist, s_warp);
        block_radix_pass<THREADS>(q_k, q_v, p_k, gv, n, 8, 0xffu, whist, s_warp);
        block_radix_pass<THREADS>(p_k, gv, q_k, q_v, n, 16, 0xffu, whist, s_warp);
        block_radix_pass<THREADS>(q_k, q_v, p_k, gv, n, 24, 0x7fu, whist, s_warp);
        if (write_keys) {
            const uint64_t hi = gk[0] & 0xffffffff00000000ull;
            __syncthreads();
            for (int i = t; i < n; i += THREADS) gk[i] = hi | p_k[i];
        }
    }
}

}  // namespace

cudaError_t launch_tile_depth_sort(cudaStream_t stream, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                                   const uint2* ranges, uint32_t n_ranges, int write_keys)
{
    if (n_ranges == 0) return cudaSuccess;
    constexpr size_t small_smem = (size_t)(kSmallThreads / 32) * kBins * 4 + (size_t)kSmallCap * 16;
    constexpr size_t large_smem = (size_t)(kLargeThreads / 32) * kBins * 4 + (size_t)kLargeCap * 16;
    auto small = tile_depth_sort_kernel<kSmallThreads, kSmallCap>;
    auto large = tile_depth_sort_kernel<kLargeThreads, kLargeCap>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)large_smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    small<<<n_ranges, kSmallThreads, small_smem, stream>>>(keys, vals, keys_tmp, vals_tmp, ranges, 0, write_keys);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    large<<<n_ranges, kLargeThreads, large_smem, stream>>>(keys, vals, keys_tmp, vals_tmp, ranges, kSmallCap, write_keys);
    return cudaGetLastError();
}

}  // namespace lgm
