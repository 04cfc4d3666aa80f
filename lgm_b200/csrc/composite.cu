// composite.cu — K5 front-to-back alpha / depth / colour compositing per 16x16 tile, and K6 its backward
// (back-to-front walk, warp-reduced gradients, vector-free fp32 reductions straight into the gradient rows).
// Replaces renderCUDA fwd/bwd of the external rasterizer; every view of the step
// is rendered by ONE launch (grid = views x tiles).
//
// Work skipping that cannot change a result: while Gaussians are staged in shared memory, the staging thread also
// derives, ONCE per Gaussian, a conservative screen-space box outside of which alpha = min(0.99, o * exp(power)) is
// certainly < 1/255 (the reference's skip threshold) and from it a bit mask of the tile's pixel patches the Gaussian can
// reach.  A warp owns an 8x4-pixel region made of NSUB = 32 / LANES patches (LANES = 32: one 8x4 patch, 16: two 4x4,
// 8: four 4x2); per 32 staged Gaussians one ballot per patch selects the Gaussians that patch must evaluate, and the
// patches of a warp walk their own lists side by side (each lane group reads its own staged record) — with the
// reference's exact, pinned per-pair arithmetic (splat_math.cuh).  A skipped pair is one the reference evaluates and
// then discards.  Smaller patches waste fewer lanes on the small Gaussians of a trained scene: on the 208-view step the
// warp iterations drop to 0.79x (4x4) / 0.68x (4x2) of the 8x4 count.
//
// Scheduling: after culling, the work of the 8 warps of a tile is uneven, so block barriers are the enemy.  512 (fwd)
// / 640 (bwd) Gaussians (24 / 30 KB of the SM's 227 KB shared memory) are staged per barrier — most tiles need one or
// two — and the warps then run independently to the end of the batch.
//
// Not HBM-bound: FP32 issue + MUFU.EX2 + shared-memory broadcast reads (and shuffles / L2 reductions in the backward).
#include "composite_common.cuh"

namespace lgm {
namespace {

// DEPTH: whether the depth image is wanted (LGM computes it and drops it, /root/reference/core/gs.py:76)
template <int LANES, bool DEPTH>
__global__ void __launch_bounds__(kBlock, 6)
composite_fwd_kernel(const RenderParams prm, const float* __restrict__ gaussians, const int32_t* __restrict__ view_scene,
                     const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                     const float* __restrict__ depth, const uint32_t* __restrict__ vals,
                     const uint2* __restrict__ ranges, const float* __restrict__ bg, int clamp_image, int batch,
                     float* __restrict__ image, float* __restrict__ alpha_img, float* __restrict__ depth_img,
                     uint32_t* __restrict__ n_contrib)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Staged* s_rec = reinterpret_cast<Staged*>(smem_raw);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gt = blockIdx.x;
    const int view = gt / prm.n_tiles;
    const int tile = gt - view * prm.n_tiles;
    const int tile_y = tile / prm.gx, tile_x = tile - tile_y * prm.gx;
    const int scene = view_scene[view];
    int px, py;
    pixel_of_thread<LANES>(tile_x, tile_y, px, py);
    const int sub = lane / LANES;
    const bool inside = px < prm.W && py < prm.H;
    const float pfx = (float)px, pfy = (float)py;
    const float tile_x0 = (float)(tile_x * kTile), tile_y0 = (float)(tile_y * kTile);

    const uint2 range = ranges[gt];
    const int todo = (int)(range.y - range.x);
    const uint32_t view_base = (uint32_t)view * (uint32_t)prm.P;
    const float* scene_g = gaussians + (size_t)scene * prm.P * 14;

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Wt = 0.f, D = 0.f;
    uint32_t last = 0;
    bool done = !inside;

    for (int r0 = 0; r0 < todo; r0 += batch) {
        if (__syncthreads_count(done) == kBlock) break;  // also the barrier that protects the staging buffer
        const int nb = min(batch, todo - r0);
        for (int k = threadIdx.x; k < nb; k += kBlock)
            stage_one<LANES, DEPTH>(s_rec[k], vals[range.x + r0 + k], view_base, scene_g, xy, conic_opacity, depth, tile_x0, tile_y0);
        __syncthreads();
        for (int base = 0; base < nb; base += 32) {
            if (__all_sync(0xffffffffu, done)) break;  // every pixel of the warp's region is saturated (or outside)
            const int jl = base + lane;
            unsigned m = patch_hits<LANES>(jl < nb ? __float_as_uint(s_rec[jl].p1.w) : 0u, warp, sub);
            // the patches walk their lists side by side; a patch whose list is exhausted idles (act = false)
            while (LANES == 32 ? (m != 0u) : __any_sync(0xffffffffu, m != 0u)) {
                const bool act = m != 0u;
                const int j = base + (act ? __ffs(m) - 1 : 0);
                m &= m - 1;
                const float4 p0 = s_rec[j].p0;
                const float4 p1 = s_rec[j].p1;
                const float dx = LGM_SUB(p0.x, pfx), dy = LGM_SUB(p0.y, pfy);
                const float power = pair_power(p0.z, p0.w, p1.x, dx, dy);
                const float a = fminf(kAlphaMax, LGM_MUL(p1.y, exp_fast(power)));
                const float test_T = LGM_MUL(T, LGM_SUB(1.0f, a));
                // A.4 in predicate form: skip if power > 0 or alpha < 1/255; stop (without compositing) if T would
                // fall below 1e-4; otherwise composite.  Lanes that do not composite add exact zeros.
                const bool cand = act && !done && !(power > 0.0f) && !(a < kAlphaMin);
                const bool stop = cand && (test_T < kTEps);
                const bool comp = cand && !stop;
                done = done || stop;
                if (!__any_sync(0xffffffffu, comp)) continue;
                const float4 cd = s_rec[j].rgbd;
                const float ae = comp ? a : 0.0f;
                C0 = LGM_FMA(LGM_MUL(cd.x, ae), T, C0);
                C1 = LGM_FMA(LGM_MUL(cd.y, ae), T, C1);
                C2 = LGM_FMA(LGM_MUL(cd.z, ae), T, C2);
                Wt = LGM_FMA(ae, T, Wt);
                if (DEPTH) D = LGM_FMA(LGM_MUL(cd.w, ae), T, D);
                T = comp ? test_T : T;
                last = comp ? (uint32_t)(r0 + j + 1) : last;  // 1-based position in the tile's list (A.4 "contributor")
            }
        }
    }
    if (inside) {
        const size_t hw = (size_t)prm.H * prm.W;
        const size_t pix = (size_t)py * prm.W + px;
        float o0 = LGM_FMA(T, __ldg(bg), C0), o1 = LGM_FMA(T, __ldg(bg + 1), C1), o2 = LGM_FMA(T, __ldg(bg + 2), C2);
        if (clamp_image) {
            // the renderer's clamp(0, 1) (/root/reference/core/gs.py:87) fused into the store; the channels whose
            // gradient the clamp blocks (value outside [0,1], or NaN) are flagged in bits 29..31 of n_contrib
            last |= (!(o0 >= 0.0f && o0 <= 1.0f) ? kClampFlag0 : 0u) | (!(o1 >= 0.0f && o1 <= 1.0f) ? kClampFlag0 << 1 : 0u) |
                    (!(o2 >= 0.0f && o2 <= 1.0f) ? kClampFlag0 << 2 : 0u);
            o0 = fminf(fmaxf(o0, 0.0f), 1.0f);
            o1 = fminf(fmaxf(o1, 0.0f), 1.0f);
            o2 = fminf(fmaxf(o2, 0.0f), 1.0f);
        }
        n_contrib[(size_t)view * hw + pix] = last;
        float* img = image + (size_t)view * 3 * hw + pix;
        img[0] = o0;
        img[hw] = o1;
        img[2 * hw] = o2;
        alpha_img[(size_t)view * hw + pix] = Wt;
        if (DEPTH) depth_img[(size_t)view * hw + pix] = D;
    }
}

// DEPTH: whether a gradient w.r.t. the depth image is given (LGM never uses the depth output, so its training step
// runs the cheaper DEPTH = false instantiation: one value less to reduce, no depth recursion).
template <int LANES, bool DEPTH>
__global__ void __launch_bounds__(kBlock, 5)
composite_bwd_kernel(const RenderParams prm, const float* __restrict__ gaussians, const int32_t* __restrict__ view_scene,
                     const float2* __restrict__ xy, const float4* __restrict__ conic_opacity,
                     const float* __restrict__ depth, const uint32_t* __restrict__ vals,
                     const uint2* __restrict__ ranges, const float* __restrict__ bg,
                     const float* __restrict__ alpha_img, const uint32_t* __restrict__ n_contrib,
                     const float* __restrict__ dL_dimage, const float* __restrict__ dL_dalpha_img,
                     const float* __restrict__ dL_ddepth_img, float* __restrict__ grad_rows, int batch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Staged* s_rec = reinterpret_cast<Staged*>(smem_raw);
    __shared__ uint32_t s_max[kBlock / 32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gt = blockIdx.x;
    const int view = gt / prm.n_tiles;
    const int tile = gt - view * prm.n_tiles;
    const int tile_y = tile / prm.gx, tile_x = tile - tile_y * prm.gx;
    const int scene = view_scene[view];
    int px, py;
    pixel_of_thread<LANES>(tile_x, tile_y, px, py);
    const int sub = lane / LANES, sl = lane % LANES;
    const bool inside = px < prm.W && py < prm.H;
    const float pfx = (float)px, pfy = (float)py;
    const float tile_x0 = (float)(tile_x * kTile), tile_y0 = (float)(tile_y * kTile);
    const size_t hw = (size_t)prm.H * prm.W;
    const size_t pix = (size_t)py * prm.W + px;

    const uint2 range = ranges[gt];
    const uint32_t view_base = (uint32_t)view * (uint32_t)prm.P;
    const float* scene_g = gaussians + (size_t)scene * prm.P * 14;

    uint32_t last_contributor = 0;
    float T_final = 0.f, dC0 = 0.f, dC1 = 0.f, dC2 = 0.f, dD = 0.f, dA = 0.f;
    if (inside) {
        const uint32_t nc = n_contrib[(size_t)view * hw + pix];
        last_contributor = nc & kContribMask;
        T_final = 1.0f - alpha_img[(size_t)view * hw + pix];
        const float* dimg = dL_dimage + (size_t)view * 3 * hw + pix;
        dC0 = (nc & kClampFlag0) ? 0.0f : dimg[0];  // clamp's gradient mask (set by the forward when it clamps)
        dC1 = (nc & (kClampFlag0 << 1)) ? 0.0f : dimg[hw];
        dC2 = (nc & (kClampFlag0 << 2)) ? 0.0f : dimg[2 * hw];
        if (DEPTH) dD = dL_ddepth_img[(size_t)view * hw + pix];
        dA = dL_dalpha_img[(size_t)view * hw + pix];
    }
    const float bgT = -T_final * (__ldg(bg) * dC0 + __ldg(bg + 1) * dC1 + __ldg(bg + 2) * dC2);

    // Only list positions below the largest n_contrib of the tile (of the patch) can contribute.
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, last_contributor);
    uint32_t pmax[Patch<LANES>::NSUB];  // per patch of this warp
    {
        uint32_t v = last_contributor;
#pragma unroll
        for (int o = LANES / 2; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
#pragma unroll
        for (int s = 0; s < Patch<LANES>::NSUB; s++) pmax[s] = __shfl_sync(0xffffffffu, v, s * LANES);
    }
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    uint32_t bmax = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; w++) bmax = max(bmax, s_max[w]);
    const int todo = (int)min(range.y - range.x, bmax);

    float T = T_final;
    float U = 0.f;  // (colour, depth, alpha) accumulated behind the current Gaussian, dotted with (dC, dD, dA)
    // which of the ten reduced sums this lane sends to the gradient row (group_reduce_10): the lanes of a group whose
    // bits below LANES/8 are zero hold the eight "a" sums (slots 0..7).  The two "b" sums (slots 8, 9) are sent by the
    // lanes sl = 1 and sl = LANES/2 + 1 where such lanes are free (LANES >= 16), else by a second reduction
    // instruction from the lanes sl = 0 and sl = LANES/2.
    const bool a_sender = (sl & (LANES / 8 - 1)) == 0;
    const int b_lane = LANES >= 16 ? 1 : 0;
    const bool b_sender = (sl & (LANES / 2 - 1)) == b_lane && (DEPTH || sl == b_lane);
    const int a_slot = sl / (LANES / 8), b_slot = 8 + sl / (LANES / 2);

    for (int r0 = 0; r0 < todo; r0 += batch) {
        __syncthreads();  // the staging buffer is free again
        const int nb = min(batch, todo - r0);
        // slot k holds list position todo-1-(r0+k): the walk is back to front
        for (int k = threadIdx.x; k < nb; k += kBlock)
            stage_one<LANES, DEPTH>(s_rec[k], vals[range.x + (uint32_t)(todo - 1 - (r0 + k))], view_base, scene_g, xy,
                                    conic_opacity, depth, tile_x0, tile_y0);
        __syncthreads();
        for (int base = 0; base < nb; base += 32) {
            const int jl = base + lane;
            // positions >= pmax[s] were never reached by patch s in the forward
            uint32_t mk = 0;
            if (jl < nb) {
                const uint32_t pos_l = (uint32_t)(todo - 1 - (r0 + jl));
                mk = __float_as_uint(s_rec[jl].p1.w);
#pragma unroll
                for (int s = 0; s < Patch<LANES>::NSUB; s++)
                    if (!(pos_l < pmax[s])) mk &= ~(1u << Patch<LANES>::bit(warp, s));
            }
            unsigned m = patch_hits<LANES>(mk, warp, sub);
            while (LANES == 32 ? (m != 0u) : __any_sync(0xffffffffu, m != 0u)) {
                const bool act = m != 0u;
                const int j = base + (act ? __ffs(m) - 1 : 0);
                m &= m - 1;
                const uint32_t pos = (uint32_t)(todo - 1 - (r0 + j));
                const float4 p0 = s_rec[j].p0;
                const float4 p1 = s_rec[j].p1;
                const float dx = LGM_SUB(p0.x, pfx), dy = LGM_SUB(p0.y, pfy);
                const float power = pair_power(p0.z, p0.w, p1.x, dx, dy);  // the forward's pinned decisions
                const float G = exp_fast(power);
                const float a = fminf(kAlphaMax, LGM_MUL(p1.y, G));
                const bool valid = act && (pos < last_contributor) && !(power > 0.0f) && !(a < kAlphaMin);
                if (!__any_sync(0xffffffffu, valid)) continue;

                float va[8], vb[2];
                {
                    // Evaluated by every lane, no divergent region: a lane that does not contribute runs with
                    // alpha = 0 and G = 0, which leaves its running state untouched and makes its ten terms exact zeros.
                    // A.5's "colour behind" recursions  acc <- a c + (1 - a) acc  (colour, depth, alpha) only ever enter
                    // through their dot product with this pixel's upstream gradient, so ONE scalar is carried:
                    //   U = acc . dC + acc_depth dD + acc_alpha dA,   U <- U + a (c . dC + depth dD + dA - U),
                    // and dL/dalpha = (c . dC + depth dD + dA - U) T + bg-term.  T <- T / (1 - a).
                    const float4 cd = s_rec[j].rgbd;
                    const float ae = valid ? a : 0.0f;
                    const float Gv = valid ? G : 0.0f;
                    float rcp;  // 1 / (1 - alpha), alpha <= 0.99: one MUFU.RCP (the backward is tolerance-checked)
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(1.0f - ae));
                    T *= rcp;
                    const float w = ae * T;
                    float cdot = fmaf(cd.x, dC0, dA);
                    cdot = fmaf(cd.y, dC1, cdot);
                    cdot = fmaf(cd.z, dC2, cdot);
                    if (DEPTH) cdot = fmaf(cd.w, dD, cdot);
                    const float e = cdot - U;
                    const float dL_da = fmaf(e, T, bgT * rcp);
                    U = fmaf(ae, e, U);
                    // The geometry gradients are linear in the moments of q = G dL/dalpha about the Gaussian's centre
                    // (dx, dy are centre - pixel, the same origin in every tile), with coefficients that depend on the
                    // Gaussian only; the row accumulates the moments and preprocess_bwd applies the coefficients
                    // (moments_to_gradients, splat_math.cuh) — 9 multiplies here instead of 21.
                    const float q = Gv * dL_da;
                    const float qx = q * dx, qy = q * dy;
                    va[0] = qx;
                    va[1] = qy;
                    va[2] = qx * dx;
                    va[3] = qx * dy;
                    va[4] = qy * dy;
                    va[5] = q;
                    va[6] = w * dC0;
                    va[7] = w * dC1;
                    vb[0] = w * dC2;
                    vb[1] = DEPTH ? w * dD : 0.0f;
                }
                float A, Bv;
                group_reduce_10<LANES, DEPTH>(va, vb, lane, A, Bv);
                // ten lanes per patch hold the ten sums: fire-and-forget fp32 reductions (RED) into the Gaussian's
                // gradient row (an idle patch sends nothing)
                float* row = grad_rows + (size_t)__float_as_uint(p1.z) * kGradRow;
                if (LANES >= 16) {
                    if (act && (a_sender || b_sender)) atomicAdd(row + (a_sender ? a_slot : b_slot), a_sender ? A : Bv);
                } else {
                    if (act) atomicAdd(row + a_slot, A);
                    if (act && b_sender) atomicAdd(row + b_slot, Bv);
                }
            }
        }
    }
}

constexpr int kMaxBatch = 2048;
// tuning hook (lgm_set_tuning "fwd_batch" / "bwd_batch"): staged Gaussians per barrier (multiple of 32, 32..2048)
int batch_or_default(Tuning which, int dflt)
{
    const int v = tuning(which);
    return (v >= 32 && v <= kMaxBatch && v % 32 == 0) ? v : dflt;
}

// lgm_set_tuning "patch_lanes": 64 (default) = two pixels per lane, 8x8 patches, packed fp32 (composite2.cu);
// 32 | 16 | 8 = the one-pixel-per-lane kernels of this file with 8x4 / 4x4 / 4x2 patches
int patch_lanes()
{
    const int v = tuning(kTunePatchLanes);
    return (v == 32 || v == 16 || v == 8) ? v : 64;
}

template <int LANES, bool DEPTH>
cudaError_t run_fwd(cudaStream_t stream, unsigned blocks, int smem, const RenderParams& prm, const float* gaussians,
                    const int32_t* view_scene, const float2* xy, const float4* conic_opacity, const float* depth,
                    const uint32_t* vals, const uint2* ranges, const float* bg, int clamp_image, int batch, float* image,
                    float* alpha, float* depth_img, uint32_t* n_contrib)
{
    static std::atomic<uint64_t> opted{0};
    if (cudaError_t e = opt_in_dynamic_smem(composite_fwd_kernel<LANES, DEPTH>, kMaxBatch * sizeof(Staged), opted)) return e;
    composite_fwd_kernel<LANES, DEPTH><<<blocks, kBlock, smem, stream>>>(prm, gaussians, view_scene, xy, conic_opacity, depth, vals, ranges,
                                                                  bg, clamp_image, batch, image, alpha, depth_img, n_contrib);
    return cudaGetLastError();
}

template <int LANES, bool DEPTH>
cudaError_t run_bwd(cudaStream_t stream, unsigned blocks, int smem, const RenderParams& prm, const float* gaussians,
                    const int32_t* view_scene, const float2* xy, const float4* conic_opacity, const float* depth,
                    const uint32_t* vals, const uint2* ranges, const float* bg, const float* alpha, const uint32_t* n_contrib,
                    const float* dL_dimage, const float* dL_dalpha, const float* dL_ddepth, float* grad_rows, int batch)
{
    static std::atomic<uint64_t> opted{0};
    if (cudaError_t e = opt_in_dynamic_smem(composite_bwd_kernel<LANES, DEPTH>, kMaxBatch * sizeof(Staged), opted)) return e;
    composite_bwd_kernel<LANES, DEPTH><<<blocks, kBlock, smem, stream>>>(prm, gaussians, view_scene, xy, conic_opacity, depth, vals,
                                                                         ranges, bg, alpha, n_contrib, dL_dimage, dL_dalpha,
                                                                         dL_ddepth, grad_rows, batch);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_composite_fwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                 const int32_t* view_scene, const float2* xy, const float4* conic_opacity,
                                 const float* depth, const uint32_t* vals, const uint2* ranges, const float* bg,
                                 int clamp_image, float* image, float* alpha, float* depth_img, uint32_t* n_contrib)
{
    const size_t blocks = (size_t)prm.n_views * prm.n_tiles;
    if (blocks == 0) return cudaSuccess;
    if (patch_lanes() == 64)
        return launch_composite2_fwd(stream, prm, gaussians, view_scene, xy, conic_opacity, depth, vals, ranges, bg, clamp_image, image,
                                     alpha, depth_img, n_contrib);
    const int batch = batch_or_default(kTuneFwdBatch, kFwdBatch);
    const int smem = batch * (int)sizeof(Staged);
#define LGM_FWD(L, D) run_fwd<L, D>(stream, (unsigned)blocks, smem, prm, gaussians, view_scene, xy, conic_opacity, depth, vals, \
                                    ranges, bg, clamp_image, batch, image, alpha, depth_img, n_contrib)
    const int lanes = patch_lanes();
    if (depth_img) {
        switch (lanes) {
            case 32: return LGM_FWD(32, true);
            case 16: return LGM_FWD(16, true);
            default: return LGM_FWD(8, true);
        }
    }
    switch (lanes) {
        case 32: return LGM_FWD(32, false);
        case 16: return LGM_FWD(16, false);
        default: return LGM_FWD(8, false);
    }
#undef LGM_FWD
}

cudaError_t launch_composite_bwd(cudaStream_t stream, const RenderParams& prm, const float* gaussians,
                                 const int32_t* view_scene, const float2* xy, const float4* conic_opacity,
                                 const float* depth, const uint32_t* vals, const uint2* ranges, const float* bg,
                                 const float* alpha, const uint32_t* n_contrib, const float* dL_dimage,
                                 const float* dL_dalpha, const float* dL_ddepth, float* grad_rows)
{
    const size_t blocks = (size_t)prm.n_views * prm.n_tiles;
    if (blocks == 0) return cudaSuccess;
    if (patch_lanes() == 64)
        return launch_composite2_bwd(stream, prm, gaussians, view_scene, xy, conic_opacity, depth, vals, ranges, bg, alpha, n_contrib,
                                     dL_dimage, dL_dalpha, dL_ddepth, grad_rows);
    const int batch = batch_or_default(kTuneBwdBatch, kBwdBatch);
    const int smem = batch * (int)sizeof(Staged);
#define LGM_BWD(L, D) run_bwd<L, D>(stream, (unsigned)blocks, smem, prm, gaussians, view_scene, xy, conic_opacity, depth, vals, \
                                    ranges, bg, alpha, n_contrib, dL_dimage, dL_dalpha, dL_ddepth, grad_rows, batch)
    const int lanes = patch_lanes();
    if (dL_ddepth) {
        switch (lanes) {
            case 32: return LGM_BWD(32, true);
            case 16: return LGM_BWD(16, true);
            default: return LGM_BWD(8, true);
        }
    }
    switch (lanes) {
        case 32: return LGM_BWD(32, false);
        case 16: return LGM_BWD(16, false);
        default: return LGM_BWD(8, false);
    }
#undef LGM_BWD
}

}  // namespace lgm
