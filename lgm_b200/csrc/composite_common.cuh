// composite_common.cuh — pieces shared by the compositing kernels (composite.cu: one pixel per lane; composite2.cu: two
// pixels per lane with packed fp32): the staged record, the culling mask, the staging gather, the warp reduction.
#pragma once
#include "common.cuh"
#include "splat_math.cuh"

namespace lgm {
namespace {

// Gaussians staged per block barrier (LGM_FWD_BATCH / LGM_BWD_BATCH override).  Measured on B200, 208 views x 98,304
// Gaussians, shipped kernels: fwd 2.91 / 2.87 / 2.88 / 2.96 / 3.68 / 3.31 ms and bwd 5.25 / 5.17 / 5.10 / 5.08 / 5.15 / 6.10 ms
// at 256 / 384 / 512 / 640 / 768 / 1024.
constexpr int kFwdBatch = 512;
constexpr int kBwdBatch = 640;
constexpr int kPatchLanes = 32;  // lanes per pixel patch (LGM_PATCH_LANES overrides): 32 = 8x4, 16 = 4x4, 8 = 4x2 pixels
constexpr uint32_t kClampFlag0 = 1u << 29;      // n_contrib bits 29..31: colour channel 0..2 was clamped
constexpr uint32_t kContribMask = kClampFlag0 - 1u;
constexpr float kCullScale = 1.002f;  // safety margins of the alpha >= 1/255 test (fp32 rounding of power / expf / logf)
constexpr float kCullPad = 2e-3f;
constexpr float kCullPix = 0.02f;     // pixels

// exp(x) for the compositing kernels: one multiply and MUFU.EX2 (relative error ~3e-7 at |x| <= 5.5, the range in which
// alpha can pass the 1/255 test) instead of libdevice's 8-instruction sequence.  Forward and backward use the same
// function, so the backward re-derives exactly the alpha (and the skip decisions) of the forward; against the oracle
// the images move by < 1e-6, far inside the 1e-4 bar.  The pinned arithmetic (splat_math.cuh) is untouched: it ends at
// `power`, everything that feeds radii, tiles and keys is upstream of this.
__device__ __forceinline__ float exp_fast(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}

struct __align__(16) Staged {
    float4 p0;    // px, py, conic xx, conic xy
    float4 p1;    // conic yy, opacity, row index (view * P + idx) as bits, patch mask as bits
    float4 rgbd;  // r, g, b, depth
};
static_assert(sizeof(Staged) == 48, "staged record is three 16-byte vectors");

// Geometry of the patches.  Warp w owns the 8x4 region at column (w & 1), row (w >> 1) of the tile; its lanes are
// split into NSUB groups of LANES, group s owning a PW x PH patch of the region.  Patch bits are numbered row-major
// over the tile's grid of (16 / PW) x (16 / PH) patches.
// LANES = 64 names the two-pixels-per-lane layout of composite2.cu: one 8x8 patch per warp, four warps per tile.
template <int LANES>
struct Patch {
    static constexpr int NSUB = LANES >= 32 ? 1 : 32 / LANES;
    static constexpr int PW = LANES >= 32 ? 8 : 4, PH = LANES == 64 ? 8 : (LANES == 8 ? 2 : 4);
    static constexpr int NCOLS = kTile / PW, NROWS = kTile / PH;
    // offset of patch s inside the warp's 8x4 region
    static __device__ __forceinline__ int sub_x(int s) { return LANES == 32 ? 0 : (LANES == 16 ? s * 4 : (s & 1) * 4); }
    static __device__ __forceinline__ int sub_y(int s) { return LANES == 8 ? (s >> 1) * 2 : 0; }
    static __device__ __forceinline__ int bit(int warp, int s)
    {
        const int x = (warp & 1) * 8 + sub_x(s), y = (warp >> 1) * 4 + sub_y(s);
        return (y / PH) * NCOLS + x / PW;
    }
};

template <int LANES>
__device__ __forceinline__ void pixel_of_thread(int tile_x, int tile_y, int& px, int& py)
{
    using PT = Patch<LANES>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = lane / LANES, sl = lane % LANES;
    px = tile_x * kTile + (warp & 1) * 8 + PT::sub_x(s) + (sl % PT::PW);
    py = tile_y * kTile + (warp >> 1) * 4 + PT::sub_y(s) + (sl / PT::PW);
}

// Which of the tile's patches can this Gaussian reach with alpha >= 1/255 ?
// alpha = min(0.99, o * exp(power)) >= 1/255 requires power >= -ln(255 o), i.e. q(d) = d^T Q d <= 2 tau, tau = ln(255 o),
// with Q = [[cx, cy], [cy, cz]] and d = centre - pixel: an ellipse around the centre.
// Computed ONCE per staged Gaussian by its staging thread; each patch then only tests its bit.
// 255 o <= 1: never visible (mask 0).  Q not positive definite (or NaN): the region is unbounded, all patches.
//
// LANES >= 32 (the 8-pixel-wide patches of the default kernels): the EXACT test — the minimum of q over the rectangle
// spanned by the patch's pixel centres against 2 tau.  q is convex with its minimum at d = 0, so over a box it is
// attained on the faces that separate the box from the origin: with (X, Y) = the box point closest to 0 per coordinate,
//   min q = min( min_dy q(X, dy),  min_dx q(dx, Y) ),  each a clamped 1-D minimisation (dy* = clamp(-cy X / cz), ...).
// Measured on the headline step (scripts/composite_stats.py): with the axis-aligned extents of the ellipse alone 28 % of
// the evaluated (patch, Gaussian) candidates had no pixel above the threshold — the Gaussians are anisotropic and
// randomly oriented, so the ellipse's bounding box is loose.
// Narrower patches (the cross-check variants LGM_PATCH_LANES = 16 / 8) keep the bounding-box test
//   hx = sqrt(2 tau cz / det Q),  hy = sqrt(2 tau cx / det Q).
// Safety margins for fp32 rounding (of this test, of the kernels' pinned `power`, of ex2.approx / __logf): the threshold
// is scaled and padded, the rectangle grown by kCullPix.
template <int LANES>
__device__ __forceinline__ uint32_t patch_mask(float px, float py, const float4 co, float tile_x0, float tile_y0)
{
    using PT = Patch<LANES>;
    constexpr uint32_t kAll = (PT::NCOLS * PT::NROWS == 32) ? 0xffffffffu : ((1u << (PT::NCOLS * PT::NROWS)) - 1u);
    const float k = 255.0f * co.w;
    if (k <= 1.0f) return 0u;
    const float det = co.x * co.z - co.y * co.y;
    if (!(det > 0.0f) || !(co.x > 0.0f) || !(co.z > 0.0f)) return kAll;
    const float t2 = 2.0f * __logf(k) * kCullScale + kCullPad;
    if (LANES >= 32) {
        const float gx = px - tile_x0, gy = py - tile_y0;  // centre relative to the tile origin
        const float kx = -__fdividef(co.y, co.x), ky = -__fdividef(co.y, co.z), cy2 = co.y + co.y;
        // fp32 rounding of q (here and in the kernels' `power`): its three terms are each <= cx dx^2 + cz dy^2 in magnitude
        // and cancel for elongated ellipses — a pad proportional to their bound over the tile
        const float ax = fabsf(gx) + (float)kTile, ay = fabsf(gy) + (float)kTile;
        const float t2p = fmaf(4e-6f, fmaf(co.x * ax, ax, co.z * ay * ay), t2);
        uint32_t m = 0;
#pragma unroll
        for (int r = 0; r < PT::NROWS; r++) {
            // d = centre - pixel over the patch's pixel centres [r PH, r PH + PH - 1] (grown by kCullPix)
            const float y_lo = gy - (float)(r * PT::PH + PT::PH - 1) - kCullPix, y_hi = gy - (float)(r * PT::PH) + kCullPix;
            const float Y = fminf(fmaxf(0.0f, y_lo), y_hi);
            const float czY2 = co.z * Y * Y, cy2Y = cy2 * Y;
#pragma unroll
            for (int c = 0; c < PT::NCOLS; c++) {
                const float x_lo = gx - (float)(c * PT::PW + PT::PW - 1) - kCullPix, x_hi = gx - (float)(c * PT::PW) + kCullPix;
                const float X = fminf(fmaxf(0.0f, x_lo), x_hi);
                const float dy = fminf(fmaxf(X * ky, y_lo), y_hi);          // minimiser on the face dx = X
                const float q1 = fmaf(fmaf(co.z, dy, cy2 * X), dy, co.x * X * X);
                const float dx = fminf(fmaxf(Y * kx, x_lo), x_hi);          // minimiser on the face dy = Y
                const float q2 = fmaf(fmaf(co.x, dx, cy2Y), dx, czY2);
                m |= (!(q1 > t2p) || !(q2 > t2p)) ? (1u << (PT::NCOLS * r + c)) : 0u;  // NaN compares false -> keep
            }
        }
        return m;
    }
    const float inv = __fdividef(t2, det);
    const float hx = sqrtf(co.z * inv) * kCullScale + kCullPix, hy = sqrtf(co.x * inv) * kCullScale + kCullPix;
    const float lo_x = px - hx - tile_x0, hi_x = px + hx - tile_x0;  // extent relative to the tile origin
    const float lo_y = py - hy - tile_y0, hi_y = py + hy - tile_y0;
    // patch column c covers pixel centres [c PW, c PW + PW - 1], row r [r PH, r PH + PH - 1].  NaN compares false -> keep.
    uint32_t cols = 0;
#pragma unroll
    for (int c = 0; c < PT::NCOLS; c++)
        cols |= (!(hi_x < (float)(c * PT::PW)) && !(lo_x > (float)(c * PT::PW + PT::PW - 1))) ? (1u << c) : 0u;
    uint32_t m = 0;
#pragma unroll
    for (int r = 0; r < PT::NROWS; r++) {
        const bool row = !(hi_y < (float)(r * PT::PH)) && !(lo_y > (float)(r * PT::PH + PT::PH - 1));
        m |= row ? (cols << (PT::NCOLS * r)) : 0u;
    }
    return m;
}

// Gather one instance (row g of the per-(view,Gaussian) arrays + its colour) into a staged record.  DEPTH = false: the
// caller uses neither the depth image nor its gradient, the depth of the instance is not fetched.
template <int LANES, bool DEPTH>
__device__ __forceinline__ void stage_one(Staged& dst, uint32_t g, uint32_t view_base, const float* __restrict__ scene_g,
                                          const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                                          const float* __restrict__ depth, float tile_x0, float tile_y0)
{
    const float2 p = xy[g];
    const float4 co = conic_opacity[g];
    const float* col = scene_g + (size_t)(g - view_base) * 14 + 11;
    dst.p0 = make_float4(p.x, p.y, co.x, co.y);
    dst.p1 = make_float4(co.z, co.w, __uint_as_float(g), __uint_as_float(patch_mask<LANES>(p.x, p.y, co, tile_x0, tile_y0)));
    dst.rgbd = make_float4(__ldg(col), __ldg(col + 1), __ldg(col + 2), DEPTH ? depth[g] : 0.0f);
}

// The per-patch hit lists of one group of 32 staged records: lane jl offers the mask of record jl (0 = not a
// candidate); the lanes of patch s get the ballot of that patch's bit.
template <int LANES>
__device__ __forceinline__ unsigned patch_hits(uint32_t mk, int warp, int sub)
{
    using PT = Patch<LANES>;
    unsigned m = 0;
#pragma unroll
    for (int s = 0; s < PT::NSUB; s++) {
        const unsigned bs = __ballot_sync(0xffffffffu, (mk >> PT::bit(warp, s)) & 1u);
        m = (sub == s) ? bs : m;
    }
    return m;
}

// Sum 10 per-lane values over each group of LANES lanes by recursive halving: at each of the first three steps a lane
// keeps half of its values and hands the other half to its partner (14 / 12 / 10 shuffles at LANES = 32 / 16 / 8).
// With h1, h2, h3 = the lane's bits LANES/2, LANES/4, LANES/8, on return the lane holds
//   A  = sum over the group of a[4 h1 + 2 h2 + h3]      (all lanes that agree on h1..h3 hold the same sum)
//   Bv = sum over the group of b[h1]                    (DEPTH = false: b[1] is identically zero, Bv = sum of b[0])
template <int LANES, bool DEPTH>
__device__ __forceinline__ void group_reduce_10(const float (&a)[8], const float (&b)[2], int lane, float& A, float& Bv)
{
    const unsigned full = 0xffffffffu;
    const bool h1 = lane & (LANES / 2), h2 = lane & (LANES / 4), h3 = lane & (LANES / 8);
    float k0 = h1 ? a[4] : a[0], k1 = h1 ? a[5] : a[1], k2 = h1 ? a[6] : a[2], k3 = h1 ? a[7] : a[3];
    const float s0 = h1 ? a[0] : a[4], s1 = h1 ? a[1] : a[5], s2 = h1 ? a[2] : a[6], s3 = h1 ? a[3] : a[7];
    float bk = DEPTH ? (h1 ? b[1] : b[0]) : b[0];
    const float bs = DEPTH ? (h1 ? b[0] : b[1]) : b[0];
    k0 += __shfl_xor_sync(full, s0, LANES / 2);
    k1 += __shfl_xor_sync(full, s1, LANES / 2);
    k2 += __shfl_xor_sync(full, s2, LANES / 2);
    k3 += __shfl_xor_sync(full, s3, LANES / 2);
    bk += __shfl_xor_sync(full, bs, LANES / 2);  // DEPTH = false: a plain butterfly, all lanes end with sum b[0]
    float m0 = h2 ? k2 : k0, m1 = h2 ? k3 : k1;
    const float t0 = h2 ? k0 : k2, t1 = h2 ? k1 : k3;
    m0 += __shfl_xor_sync(full, t0, LANES / 4);
    m1 += __shfl_xor_sync(full, t1, LANES / 4);
    bk += __shfl_xor_sync(full, bk, LANES / 4);
    A = h3 ? m1 : m0;
    const float u = h3 ? m0 : m1;
    A += __shfl_xor_sync(full, u, LANES / 8);
    bk += __shfl_xor_sync(full, bk, LANES / 8);
#pragma unroll
    for (int o = LANES / 16; o >= 1; o >>= 1) {
        A += __shfl_xor_sync(full, A, o);
        bk += __shfl_xor_sync(full, bk, o);
    }
    Bv = bk;
}

}  // namespace
}  // namespace lgm
