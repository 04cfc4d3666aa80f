// composite2.cu — K5 / K6 with TWO PIXELS PER LANE and packed fp32 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2).
//
// Blackwell issues `fma.rn.f32x2 / mul.rn.f32x2 / add.rn.f32x2` — two correctly rounded fp32 operations on a register
// pair — in ONE issue slot, and a scalar operand is broadcast to both halves for free (SASS `R5.F32`).  The compositing
// kernels are bound by warp-instruction issue (profiles/: 80 % of the 4 x 148 schedulers busy, 5 % DRAM), so the lever
// is instructions per pixel.  Here a lane owns the pixels (x, y) and (x, y + 4) of its warp's 8x8 patch (4 warps = one
// 16x16 tile): everything that depends on the pixel — offset, power, alpha, transmittance, the colour / weight
// accumulation, and in the backward the whole gradient body — runs as ONE packed instruction for both pixels, with the
// Gaussian's parameters as the broadcast scalar operand; what depends on the Gaussian only (dx, cx dx, cy dx) is computed
// once.  Each packed half is the same correctly rounded operation, in the same order, as the one-pixel-per-lane kernels
// (composite.cu) and as the pinned sequence of splat_math.cuh, so images, n_contrib and the skip decisions are
// bit-identical (tests/test_gpu_parity.py::test_composite_variants_agree).  In the backward the 14-shuffle warp
// reduction is paid once per 64 pixels instead of once per 32.
//
// Same culling as composite.cu (a per-Gaussian mask of the patches it can reach, computed once at staging), at the
// granularity of the 8x8 patches.
#include "composite_common.cuh"

namespace lgm {
namespace {

constexpr int kBlock2 = 128;      // 4 warps x (8x8 pixels)
// Gaussians staged per block barrier (lgm_set_tuning fwd_batch / bwd_batch override).  Measured on B200, 208 views x 98,304
// Gaussians: 256 / 384 / 512 / 640 / 768 / 1024 -> fwd 2.76 / 2.79 / 3.31 / 3.43 / 3.63 / 4.11 ms, bwd 4.55 / 4.52 / 4.82 /
// 4.96 / 5.25 / 6.43 ms (larger batches cost resident CTAs: 48 B of shared memory per staged Gaussian)
constexpr int kFwdBatch2 = 256;
constexpr int kBwdBatch2 = 384;
constexpr int kSparseLanes2 = 14;  // backward: hits with at most this many contributing lanes use direct vector reductions
// (measured, headline step: 0 / 2 / 4 / 8 / 12 / 16 / 24 / 32 -> bwd 4.41 / 4.22 / 4.10 / 3.93 / 3.83 / 3.83 / 4.37 / 6.14 ms)
constexpr int kFwdOcc2 = 10, kBwdOcc2 = 8;  // resident CTAs per SM the kernels are compiled for (launch bounds)
// A pixel that has stopped (or lies outside the image) is parked at row 1e18: its dy is astronomically large, so the
// Gaussian's exponent is hugely negative (conics are >= ~1e-7), alpha underflows to 0 and the reference's own
// "alpha < 1/255 -> skip" test rejects the pair — no `done` flag in the inner loop.  (A non-positive-definite conic gives
// power > 0 or -inf: skipped as well.)
constexpr float kParked = 1e18f;

#ifdef LGM_STATS
// Developer build only (scripts/composite_stats.py, -DLGM_STATS): per-hit statistics of the compositing kernels.
// [0] fwd candidates  [1] fwd hits (some pixel composites)  [2] fwd composited pixels  [3] fwd staged instances
// [8] bwd candidates  [9] bwd hits  [10] bwd valid pixels  [11] bwd staged instances
// [16 + b] bwd hits whose number of lanes with a valid pixel falls in bucket b: 1, 2, 3-4, 5-8, 9-16, 17-32
__device__ unsigned long long g_stats[32];
// the value is evaluated by every lane (it may contain warp votes)
#define LGM_STAT(i, v) do { const unsigned long long sv_ = (v); if (lane == 0) st[i] += sv_; } while (0)
#else
#define LGM_STAT(i, v) do { } while (0)
#endif

// ---- packed fp32 (PTX ISA 8.6, sm_100+): both halves are .rn operations, never contracted or re-associated ----
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }          // broadcast (free: an .F32 operand)
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }  // folds into the operand's sign

// pair_power (splat_math.cuh) for the pixels (dx, dy.x) and (dx, dy.y): the same operations in the same order
__device__ __forceinline__ float2 pair_power2(float cx, float cy, float cz, float dx, float2 dy)
{
    const float t1 = LGM_MUL(cx, dx);
    const float2 t2 = mul2(mul2(bc(cz), dy), dy);
    const float2 s = fma2(bc(t1), bc(dx), t2);
    const float2 u = mul2(bc(LGM_MUL(cy, dx)), dy);
    return fma2(s, bc(-0.5f), neg2(u));
}
// exp_fast for both halves
__device__ __forceinline__ float2 exp_fast2(float2 x)
{
    const float2 e = mul2(x, bc(1.4426950408889634f));
    float2 r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(e.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(e.y));
    return r;
}

// Staging buffer as four arrays (conflict-free 16-byte stores: a 48-byte record stride would serialise every STS.128
// four ways), 52 B per staged Gaussian: p0[], p1[], rgbd[] as the fields of Staged, mask[] = the patch mask on its own.
constexpr int kStagedBytes2 = 3 * 16 + 4;
// BATCH is a compile-time constant so that the three other arrays are immediate offsets from one base register.
template <int BATCH>
struct Staging2 {
    float4* p0;
    float4* const p1;
    float4* const rgbd;
    uint32_t* const mask;
    __device__ __forceinline__ Staging2(unsigned char* base)
        : p0(reinterpret_cast<float4*>(base)), p1(p0 + BATCH), rgbd(p0 + 2 * BATCH), mask(reinterpret_cast<uint32_t*>(p0 + 3 * BATCH)) {}
    template <bool DEPTH>
    __device__ __forceinline__ void stage(int k, uint32_t g, uint32_t view_base, const float* __restrict__ scene_g,
                                          const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                                          const float* __restrict__ depth, float tile_x0, float tile_y0) const
    {
        Staged r;
        stage_one<64, DEPTH>(r, g, view_base, scene_g, xy, conic_opacity, depth, tile_x0, tile_y0);
        p0[k] = r.p0;
        p1[k] = r.p1;
        rgbd[k] = r.rgbd;
        mask[k] = __float_as_uint(r.p1.w);
    }
};

// this lane's pixels: column x, rows y0 and y0 + 4 of the warp's 8x8 patch
__device__ __forceinline__ void pixels_of_lane(int tile_x, int tile_y, int& px, int& py0)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    px = tile_x * kTile + (warp & 1) * 8 + (lane & 7);
    py0 = tile_y * kTile + (warp >> 1) * 8 + (lane >> 3);
}

template <bool DEPTH, int MINB, int BATCH>
__global__ void __launch_bounds__(kBlock2, MINB)
composite2_fwd_kernel(const RenderParams prm, const float* __restrict__ gaussians, const int32_t* __restrict__ view_scene,
                      const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                      const float* __restrict__ depth, const uint32_t* __restrict__ vals,
                      const uint2* __restrict__ ranges, const float* __restrict__ bg, int clamp_image,
                      float* __restrict__ image, float* __restrict__ alpha_img, float* __restrict__ depth_img,
                      uint32_t* __restrict__ n_contrib)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Staging2<BATCH> sb(smem_raw);
    constexpr int batch = BATCH;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gt = blockIdx.x;
    const int view = gt / prm.n_tiles;
    const int tile = gt - view * prm.n_tiles;
    const int tile_y = tile / prm.gx, tile_x = tile - tile_y * prm.gx;
    const int scene = view_scene[view];
    int px, py0;
    pixels_of_lane(tile_x, tile_y, px, py0);
    const int py1 = py0 + 4;
    const bool in0 = px < prm.W && py0 < prm.H, in1 = px < prm.W && py1 < prm.H;
    const float pfx = (float)px;
    float2 npfy = make_float2(in0 ? -(float)py0 : -kParked, in1 ? -(float)py1 : -kParked);
    const float tile_x0 = (float)(tile_x * kTile), tile_y0 = (float)(tile_y * kTile);

    const uint2 range = ranges[gt];
    const int todo = (int)(range.y - range.x);
    const uint32_t view_base = (uint32_t)view * (uint32_t)prm.P;
    const float* scene_g = gaussians + (size_t)scene * prm.P * 14;

    float2 T = bc(1.0f), C0 = bc(0.f), C1 = bc(0.f), C2 = bc(0.f), Wt = bc(0.f), D = bc(0.f);
    uint32_t last0 = 0, last1 = 0;
#ifdef LGM_STATS
    unsigned long long st[4] = {0, 0, 0, 0};
#endif
#define LGM_BOTH_PARKED (npfy.x < -0.5f * kParked && npfy.y < -0.5f * kParked)

    for (int r0 = 0; r0 < todo; r0 += batch) {
        if (__syncthreads_count(LGM_BOTH_PARKED) == kBlock2) break;  // also the barrier that protects the staging buffer
        const int nb = min(batch, todo - r0);
        {
            // the instance indices of this thread's slots first (independent loads), then the gathers they address
            constexpr int kPer = BATCH / kBlock2;
            uint32_t gi[kPer];
#pragma unroll
            for (int u = 0; u < kPer; u++) {
                const int k = threadIdx.x + u * kBlock2;
                gi[u] = k < nb ? vals[range.x + r0 + k] : 0u;
            }
#pragma unroll
            for (int u = 0; u < kPer; u++) {
                const int k = threadIdx.x + u * kBlock2;
                if (k < nb) sb.template stage<DEPTH>(k, gi[u], view_base, scene_g, xy, conic_opacity, depth, tile_x0, tile_y0);
            }
        }
        __syncthreads();
        if (warp == 0) LGM_STAT(3, nb);
        for (int base = 0; base < nb; base += 32) {
            if (__all_sync(0xffffffffu, LGM_BOTH_PARKED)) break;  // every pixel of the warp's patch is saturated (or outside)
            const int jl = base + lane;
            unsigned m = __ballot_sync(0xffffffffu, jl < nb && ((sb.mask[jl] >> warp) & 1u));
            while (m != 0u) {
                const int j = base + __ffs(m) - 1;
                m &= m - 1;
                const float4 p0 = sb.p0[j];
                const float4 p1 = sb.p1[j];
                const float dx = LGM_SUB(p0.x, pfx);
                const float2 dy = add2(bc(p0.y), npfy);
                const float2 power = pair_power2(p0.z, p0.w, p1.x, dx, dy);
                const float2 ar = mul2(bc(p1.y), exp_fast2(power));
                const float2 a = make_float2(fminf(kAlphaMax, ar.x), fminf(kAlphaMax, ar.y));
                const float2 test_T = mul2(T, add2(bc(1.0f), neg2(a)));
                // A.4 in predicate form, per pixel: skip if power > 0 or alpha < 1/255; stop (without compositing) if T
                // would fall below 1e-4; otherwise composite.  Pixels that do not composite add exact zeros.
                const bool cand0 = !(power.x > 0.0f) && !(a.x < kAlphaMin);   // (a parked pixel has alpha = 0)
                const bool cand1 = !(power.y > 0.0f) && !(a.y < kAlphaMin);
                const bool stop0 = cand0 && (test_T.x < kTEps), stop1 = cand1 && (test_T.y < kTEps);
                const bool comp0 = cand0 && !stop0, comp1 = cand1 && !stop1;
                npfy.x = stop0 ? -kParked : npfy.x;
                npfy.y = stop1 ? -kParked : npfy.y;
                LGM_STAT(0, 1);
                if (!__any_sync(0xffffffffu, comp0 || comp1)) continue;
#ifdef LGM_STATS
                LGM_STAT(1, 1);
                LGM_STAT(2, __popc(__ballot_sync(0xffffffffu, comp0)) + __popc(__ballot_sync(0xffffffffu, comp1)));
#endif
                const float4 cd = sb.rgbd[j];
                const float2 ae = make_float2(comp0 ? a.x : 0.0f, comp1 ? a.y : 0.0f);
                C0 = fma2(mul2(bc(cd.x), ae), T, C0);
                C1 = fma2(mul2(bc(cd.y), ae), T, C1);
                C2 = fma2(mul2(bc(cd.z), ae), T, C2);
                Wt = fma2(ae, T, Wt);
                if (DEPTH) D = fma2(mul2(bc(cd.w), ae), T, D);
                T = make_float2(comp0 ? test_T.x : T.x, comp1 ? test_T.y : T.y);
                const uint32_t pos1 = (uint32_t)(r0 + j + 1);  // 1-based position in the tile's list (A.4 "contributor")
                last0 = comp0 ? pos1 : last0;
                last1 = comp1 ? pos1 : last1;
            }
        }
    }
#undef LGM_BOTH_PARKED
#ifdef LGM_STATS
    if (lane == 0)
        for (int i = 0; i < 4; i++) atomicAdd(&g_stats[i], st[i]);
#endif
    const size_t hw = (size_t)prm.H * prm.W;
    const float b0 = __ldg(bg), b1 = __ldg(bg + 1), b2 = __ldg(bg + 2);
    const float2 o0 = fma2(T, bc(b0), C0), o1 = fma2(T, bc(b1), C1), o2 = fma2(T, bc(b2), C2);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (!(h ? in1 : in0)) continue;
        float v0 = h ? o0.y : o0.x, v1 = h ? o1.y : o1.x, v2 = h ? o2.y : o2.x;
        uint32_t last = h ? last1 : last0;
        if (clamp_image) {
            // the renderer's clamp(0, 1) (/root/reference/core/gs.py:87) fused into the store; the channels whose
            // gradient the clamp blocks (value outside [0,1], or NaN) are flagged in bits 29..31 of n_contrib
            last |= (!(v0 >= 0.0f && v0 <= 1.0f) ? kClampFlag0 : 0u) | (!(v1 >= 0.0f && v1 <= 1.0f) ? kClampFlag0 << 1 : 0u) |
                    (!(v2 >= 0.0f && v2 <= 1.0f) ? kClampFlag0 << 2 : 0u);
            v0 = fminf(fmaxf(v0, 0.0f), 1.0f);
            v1 = fminf(fmaxf(v1, 0.0f), 1.0f);
            v2 = fminf(fmaxf(v2, 0.0f), 1.0f);
        }
        const size_t pix = (size_t)(h ? py1 : py0) * prm.W + px;
        n_contrib[(size_t)view * hw + pix] = last;
        float* img = image + (size_t)view * 3 * hw + pix;
        img[0] = v0;
        img[hw] = v1;
        img[2 * hw] = v2;
        alpha_img[(size_t)view * hw + pix] = h ? Wt.y : Wt.x;
        if (DEPTH) depth_img[(size_t)view * hw + pix] = h ? D.y : D.x;
    }
}

// group_reduce_10<32> (composite_common.cuh) with the additions of its first two levels packed (FADD2): the same sums on
// the same lanes.
template <bool DEPTH>
__device__ __forceinline__ void warp_reduce_10_packed(const float (&a)[8], const float (&b)[2], int lane, float& A, float& Bv)
{
    const unsigned full = 0xffffffffu;
    const bool h1 = lane & 16, h2 = lane & 8, h3 = lane & 4;
    float2 k01 = make_float2(h1 ? a[4] : a[0], h1 ? a[5] : a[1]), k23 = make_float2(h1 ? a[6] : a[2], h1 ? a[7] : a[3]);
    const float s0 = h1 ? a[0] : a[4], s1 = h1 ? a[1] : a[5], s2 = h1 ? a[2] : a[6], s3 = h1 ? a[3] : a[7];
    float bk = DEPTH ? (h1 ? b[1] : b[0]) : b[0];
    const float bs = DEPTH ? (h1 ? b[0] : b[1]) : b[0];
    k01 = add2(k01, make_float2(__shfl_xor_sync(full, s0, 16), __shfl_xor_sync(full, s1, 16)));
    k23 = add2(k23, make_float2(__shfl_xor_sync(full, s2, 16), __shfl_xor_sync(full, s3, 16)));
    bk += __shfl_xor_sync(full, bs, 16);  // DEPTH = false: a plain butterfly, all lanes end with sum b[0]
    float2 m = h2 ? k23 : k01;
    const float2 t = h2 ? k01 : k23;
    m = add2(m, make_float2(__shfl_xor_sync(full, t.x, 8), __shfl_xor_sync(full, t.y, 8)));
    bk += __shfl_xor_sync(full, bk, 8);
    A = h3 ? m.y : m.x;
    const float u = h3 ? m.x : m.y;
    A += __shfl_xor_sync(full, u, 4);
    bk += __shfl_xor_sync(full, bk, 4);
    A += __shfl_xor_sync(full, A, 2);
    bk += __shfl_xor_sync(full, bk, 2);
    A += __shfl_xor_sync(full, A, 1);
    bk += __shfl_xor_sync(full, bk, 1);
    Bv = bk;
}

// Vector reductions into a 16-byte aligned gradient row (sm_90+: REDG.E.ADD.F32x4 / .F32x2), fire and forget.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

template <bool DEPTH, int MINB, int BATCH>
__global__ void __launch_bounds__(kBlock2, MINB)
composite2_bwd_kernel(const RenderParams prm, const float* __restrict__ gaussians, const int32_t* __restrict__ view_scene,
                      const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                      const float* __restrict__ depth, const uint32_t* __restrict__ vals,
                      const uint2* __restrict__ ranges, const float* __restrict__ bg,
                      const float* __restrict__ alpha_img, const uint32_t* __restrict__ n_contrib,
                      const float* __restrict__ dL_dimage, const float* __restrict__ dL_dalpha_img,
                      const float* __restrict__ dL_ddepth_img, float* __restrict__ grad_rows, int sparse_lanes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Staging2<BATCH> sb(smem_raw);
    constexpr int batch = BATCH;
    __shared__ uint32_t s_max[kBlock2 / 32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gt = blockIdx.x;
    const int view = gt / prm.n_tiles;
    const int tile = gt - view * prm.n_tiles;
    const int tile_y = tile / prm.gx, tile_x = tile - tile_y * prm.gx;
    const int scene = view_scene[view];
    int px, py0;
    pixels_of_lane(tile_x, tile_y, px, py0);
    const int py1 = py0 + 4;
    const bool in0 = px < prm.W && py0 < prm.H, in1 = px < prm.W && py1 < prm.H;
    const float pfx = (float)px;
    const float2 npfy = make_float2(-(float)py0, -(float)py1);
    const float tile_x0 = (float)(tile_x * kTile), tile_y0 = (float)(tile_y * kTile);
    const size_t hw = (size_t)prm.H * prm.W;

    const uint2 range = ranges[gt];
    const uint32_t view_base = (uint32_t)view * (uint32_t)prm.P;
    const float* scene_g = gaussians + (size_t)scene * prm.P * 14;

    uint32_t lc0 = 0, lc1 = 0;  // last contributor of the two pixels
    float2 T = bc(0.f), dC0 = bc(0.f), dC1 = bc(0.f), dC2 = bc(0.f), dD = bc(0.f), dA = bc(0.f);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (!(h ? in1 : in0)) continue;
        const size_t pix = (size_t)(h ? py1 : py0) * prm.W + px;
        const uint32_t nc = n_contrib[(size_t)view * hw + pix];
        const float* dimg = dL_dimage + (size_t)view * 3 * hw + pix;
        const float tf = 1.0f - alpha_img[(size_t)view * hw + pix];
        const float c0 = (nc & kClampFlag0) ? 0.0f : dimg[0];  // clamp's gradient mask (set by the forward when it clamps)
        const float c1 = (nc & (kClampFlag0 << 1)) ? 0.0f : dimg[hw];
        const float c2 = (nc & (kClampFlag0 << 2)) ? 0.0f : dimg[2 * hw];
        const float da = dL_dalpha_img[(size_t)view * hw + pix];
        const float dd = DEPTH ? dL_ddepth_img[(size_t)view * hw + pix] : 0.0f;
        if (h) { lc1 = nc & kContribMask; T.y = tf; dC0.y = c0; dC1.y = c1; dC2.y = c2; dA.y = da; dD.y = dd; }
        else   { lc0 = nc & kContribMask; T.x = tf; dC0.x = c0; dC1.x = c1; dC2.x = c2; dA.x = da; dD.x = dd; }
    }
    const float b0 = __ldg(bg), b1 = __ldg(bg + 1), b2 = __ldg(bg + 2);
    // bgT = -T_final (bg . dC)
    const float2 bgT = mul2(neg2(T), fma2(bc(b2), dC2, fma2(bc(b1), dC1, mul2(bc(b0), dC0))));

    // Only list positions below the largest n_contrib of the tile (of the warp's patch) can contribute.
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, max(lc0, lc1));
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    uint32_t bmax = 0;
#pragma unroll
    for (int w = 0; w < kBlock2 / 32; w++) bmax = max(bmax, s_max[w]);
    const int todo = (int)min(range.y - range.x, bmax);

#ifdef LGM_STATS
    unsigned long long st[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    float2 U = bc(0.f);  // (colour, depth, alpha) accumulated behind the current Gaussian, dotted with (dC, dD, dA)
    // which of the ten reduced sums this lane sends to the gradient row (group_reduce_10<32>): lanes 0, 4, .., 28 hold the
    // eight "a" sums (slots 0..7), lanes 1 and 17 the two "b" sums (slots 8, 9)
    const bool a_sender = (lane & 3) == 0;
    const bool b_sender = (lane & 15) == 1 && (DEPTH || lane == 1);
    const int a_slot = lane >> 2, b_slot = 8 + (lane >> 4);

    for (int r0 = 0; r0 < todo; r0 += batch) {
        __syncthreads();  // the staging buffer is free again
        const int nb = min(batch, todo - r0);
        // slot k holds list position todo-1-(r0+k): the walk is back to front
        // (loading the thread's instance indices first, as the forward kernel does, measured slower here: 3.95 vs 3.81 ms)
        for (int k = threadIdx.x; k < nb; k += kBlock2)
            sb.template stage<DEPTH>(k, vals[range.x + (uint32_t)(todo - 1 - (r0 + k))], view_base, scene_g, xy, conic_opacity, depth,
                                     tile_x0, tile_y0);
        __syncthreads();
        if (warp == 0) LGM_STAT(3, nb);
        for (int base = 0; base < nb; base += 32) {
            const int jl = base + lane;
            // positions >= wmax were never reached by this warp's patch in the forward
            bool cand = false;
            if (jl < nb) {
                const uint32_t pos_l = (uint32_t)(todo - 1 - (r0 + jl));
                cand = ((sb.mask[jl] >> warp) & 1u) && pos_l < wmax;
            }
            unsigned m = __ballot_sync(0xffffffffu, cand);
            while (m != 0u) {
                const int j = base + __ffs(m) - 1;
                m &= m - 1;
                const uint32_t pos = (uint32_t)(todo - 1 - (r0 + j));
                const float4 p0 = sb.p0[j];
                const float4 p1 = sb.p1[j];
                const float dx = LGM_SUB(p0.x, pfx);
                const float2 dy = add2(bc(p0.y), npfy);
                const float2 power = pair_power2(p0.z, p0.w, p1.x, dx, dy);  // the forward's pinned decisions
                const float2 G = exp_fast2(power);
                const float2 ar = mul2(bc(p1.y), G);
                const float2 a = make_float2(fminf(kAlphaMax, ar.x), fminf(kAlphaMax, ar.y));
                const bool valid0 = (pos < lc0) && !(power.x > 0.0f) && !(a.x < kAlphaMin);
                const bool valid1 = (pos < lc1) && !(power.y > 0.0f) && !(a.y < kAlphaMin);
                LGM_STAT(0, 1);
                const unsigned vm = __ballot_sync(0xffffffffu, valid0 || valid1);  // lanes with a contributing pixel
                if (vm == 0u) continue;
#ifdef LGM_STATS
                {
                    const int nl = __popc(vm);
                    LGM_STAT(1, 1);
                    LGM_STAT(2, __popc(__ballot_sync(0xffffffffu, valid0)) + __popc(__ballot_sync(0xffffffffu, valid1)));
                    LGM_STAT(8 + (nl <= 1 ? 0 : nl <= 2 ? 1 : nl <= 4 ? 2 : nl <= 8 ? 3 : nl <= 16 ? 4 : 5), 1);
                }
#endif

                // Evaluated for both pixels of every lane, no divergent region: a pixel that does not contribute runs with
                // alpha = 0 and G = 0, which leaves its running state untouched and makes its ten terms exact zeros.  One
                // scalar per pixel carries A.5's "colour behind" recursions (see composite.cu):
                //   U <- U + a (c . dC + depth dD + dA - U),  dL/dalpha = (c . dC + depth dD + dA - U) T + bg-term,  T <- T / (1 - a)
                const float4 cd = sb.rgbd[j];
                const float2 ae = make_float2(valid0 ? a.x : 0.0f, valid1 ? a.y : 0.0f);
                const float2 Gv = make_float2(valid0 ? G.x : 0.0f, valid1 ? G.y : 0.0f);
                const float2 om = add2(bc(1.0f), neg2(ae));
                float2 rcp;  // 1 / (1 - alpha), alpha <= 0.99: MUFU.RCP (the backward is tolerance-checked)
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp.x) : "f"(om.x));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp.y) : "f"(om.y));
                T = mul2(T, rcp);
                const float2 w = mul2(ae, T);
                float2 cdot = fma2(bc(cd.x), dC0, dA);
                cdot = fma2(bc(cd.y), dC1, cdot);
                cdot = fma2(bc(cd.z), dC2, cdot);
                if (DEPTH) cdot = fma2(bc(cd.w), dD, cdot);
                const float2 e = add2(cdot, neg2(U));
                const float2 dL_da = fma2(e, T, mul2(bgT, rcp));
                U = fma2(ae, e, U);
                // moments of q = G dL/dalpha about the Gaussian's centre (see composite.cu / moments_to_gradients)
                const float2 q = mul2(Gv, dL_da);
                const float2 qx = mul2(q, bc(dx)), qy = mul2(q, dy);
                const float2 qxx = mul2(qx, bc(dx)), qxy = mul2(qx, dy), qyy = mul2(qy, dy);
                const float2 g0 = mul2(w, dC0), g1 = mul2(w, dC1), g2 = mul2(w, dC2);
                float va[8], vb[2];
                va[0] = qx.x + qx.y;
                va[1] = qy.x + qy.y;
                va[2] = qxx.x + qxx.y;
                va[3] = qxy.x + qxy.y;
                va[4] = qyy.x + qyy.y;
                va[5] = q.x + q.y;
                va[6] = g0.x + g0.y;
                va[7] = g1.x + g1.y;
                vb[0] = g2.x + g2.y;
                if (DEPTH) {
                    const float2 gd = mul2(w, dD);
                    vb[1] = gd.x + gd.y;
                } else {
                    vb[1] = 0.0f;
                }
                float* row = grad_rows + (size_t)__float_as_uint(p1.z) * kGradRow;
                if (__popc(vm) <= sparse_lanes) {
                    // A hit that only a few lanes contribute to (a Gaussian clipping the patch: 18 % of the hits of the
                    // headline step have one or two such lanes): those lanes send their own terms with three vector
                    // reductions instead of the whole warp running the 14-shuffle reduction for them.
                    if (valid0 || valid1) {
                        red_add_v4(row, va[0], va[1], va[2], va[3]);
                        red_add_v4(row + 4, va[4], va[5], va[6], va[7]);
                        if (DEPTH) red_add_v2(row + 8, vb[0], vb[1]);
                        else atomicAdd(row + 8, vb[0]);
                    }
                    continue;
                }
                float A, Bv;
                warp_reduce_10_packed<DEPTH>(va, vb, lane, A, Bv);
                // ten lanes hold the ten sums: fire-and-forget fp32 reductions (RED) into the Gaussian's gradient row
                if (a_sender || b_sender) atomicAdd(row + (a_sender ? a_slot : b_slot), a_sender ? A : Bv);
            }
        }
    }
#ifdef LGM_STATS
    if (lane == 0)
        for (int i = 0; i < 16; i++) atomicAdd(&g_stats[8 + i], st[i]);
#endif
}

template <bool DEPTH, int MINB, int BATCH>
cudaError_t run2_fwd(cudaStream_t stream, unsigned blocks, const RenderParams& prm, const float* gaussians,
                     const int32_t* view_scene, const float2* xy, const float4* conic_opacity, const float* depth,
                     const uint32_t* vals, const uint2* ranges, const float* bg, int clamp_image, float* image, float* alpha,
                     float* depth_img, uint32_t* n_contrib)
{
    composite2_fwd_kernel<DEPTH, MINB, BATCH><<<blocks, kBlock2, BATCH * kStagedBytes2, stream>>>(
        prm, gaussians, view_scene, xy, conic_opacity, depth, vals, ranges, bg, clamp_image, image, alpha, depth_img, n_contrib);
    return cudaGetLastError();
}

template <bool DEPTH, int MINB, int BATCH>
cudaError_t run2_bwd(cudaStream_t stream, unsigned blocks, const RenderParams& prm, const float* gaussians,
                     const int32_t* view_scene, const float2* xy, const float4* conic_opacity, const float* depth,
                     const uint32_t* vals, const uint2* ranges, const float* bg, const float* alpha, const uint32_t* n_contrib,
                     const float* dL_dimage, const float* dL_dalpha, const float* dL_ddepth, float* grad_rows)
{
    // lgm_set_tuning "sparse_lanes": hits with at most this many contributing lanes skip the warp reduction (0 = never)
    const int sparse_lanes = tuning(kTuneSparseLanes) >= 0 ? tuning(kTuneSparseLanes) : kSparseLanes2;
    composite2_bwd_kernel<DEPTH, MINB, BATCH><<<blocks, kBlock2, BATCH * kStagedBytes2, stream>>>(
        prm, gaussians, view_scene, xy, conic_opacity, depth, vals, ranges, bg, alpha, n_contrib, dL_dimage, dL_dalpha, dL_ddepth,
        grad_rows, sparse_lanes);
    return cudaGetLastError();
}

}  // namespace

// Launch shapes.  The staging batch and the resident CTAs per SM (launch bounds) are template parameters; lgm_set_tuning
// "fwd_batch" / "bwd_batch" (256 | 384; 20 KB fit the default shared-memory limit) and "c2_occ" (100 x forward + backward
// CTAs per SM) select among the instantiations.
cudaError_t launch_composite2_fwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                  const int32_t* view_scene, const float2* xy, const float4* conic_opacity,
                                  const float* depth, const uint32_t* vals, const uint2* ranges, const float* bg,
                                  int clamp_image, float* image, float* alpha, float* depth_img, uint32_t* n_contrib)
{
    const size_t blocks = (size_t)prm.n_views * prm.n_tiles;
    if (blocks == 0) return cudaSuccess;
    const int batch = tuning(kTuneFwdBatch) == 384 ? 384 : kFwdBatch2;
    const int occ = tuning(kTuneC2Occ) >= 0 ? tuning(kTuneC2Occ) / 100 : kFwdOcc2;
#define LGM_F(D, M, B) run2_fwd<D, M, B>(stream, (unsigned)blocks, prm, gaussians, view_scene, xy, conic_opacity, depth, vals, ranges, bg, \
                                         clamp_image, image, alpha, depth_img, n_contrib)
#define LGM_FB(D, M) (batch == 384 ? LGM_F(D, M, 384) : LGM_F(D, M, 256))
    if (depth_img) return occ == 12 ? LGM_FB(true, 12) : (occ == 8 ? LGM_FB(true, 8) : LGM_FB(true, 10));
    return occ == 12 ? LGM_FB(false, 12) : (occ == 8 ? LGM_FB(false, 8) : LGM_FB(false, 10));
#undef LGM_FB
#undef LGM_F
}

cudaError_t launch_composite2_bwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                  const int32_t* view_scene, const float2* xy, const float4* conic_opacity,
                                  const float* depth, const uint32_t* vals, const uint2* ranges, const float* bg,
                                  const float* alpha, const uint32_t* n_contrib, const float* dL_dimage,
                                  const float* dL_dalpha, const float* dL_ddepth, float* grad_rows)
{
    const size_t blocks = (size_t)prm.n_views * prm.n_tiles;
    if (blocks == 0) return cudaSuccess;
    const int batch = tuning(kTuneBwdBatch) == 256 ? 256 : kBwdBatch2;
    const int occ = tuning(kTuneC2Occ) >= 0 ? tuning(kTuneC2Occ) % 100 : kBwdOcc2;
#define LGM_B(D, M, B) run2_bwd<D, M, B>(stream, (unsigned)blocks, prm, gaussians, view_scene, xy, conic_opacity, depth, vals, ranges, bg, \
                                         alpha, n_contrib, dL_dimage, dL_dalpha, dL_ddepth, grad_rows)
#define LGM_BB(D, M) (batch == 256 ? LGM_B(D, M, 256) : LGM_B(D, M, 384))
    if (dL_ddepth) return occ == 10 ? LGM_BB(true, 10) : (occ == 6 ? LGM_BB(true, 6) : LGM_BB(true, 8));
    return occ == 10 ? LGM_BB(false, 10) : (occ == 6 ? LGM_BB(false, 6) : LGM_BB(false, 8));
#undef LGM_BB
#undef LGM_B
}

#ifdef LGM_STATS
cudaError_t debug_stats(unsigned long long* host_out, int reset)
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && host_out) e = cudaMemcpyFromSymbol(host_out, g_stats, sizeof(g_stats));
    if (e == cudaSuccess && reset) {
        const unsigned long long z[32] = {};
        e = cudaMemcpyToSymbol(g_stats, z, sizeof(z));
    }
    return e;
}
#endif

}  // namespace lgm

#ifdef LGM_STATS
extern "C" int lgm_debug_stats(unsigned long long* host_out, int reset) { return (int)lgm::debug_stats(host_out, reset); }
#endif
