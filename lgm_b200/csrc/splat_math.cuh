// splat_math.cuh — per-Gaussian / per-pair arithmetic of the splat-render path, pinned operation by operation.
//
// Arithmetic contract (DESIGN.md §"arithmetic contract"): every expression that feeds radii, tile rects, sort keys
// or the alpha thresholds is spelled as an explicit sequence of correctly-rounded fp32 operations
// (__fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn / __fsqrt_rn), which neither nvcc nor ptxas may re-associate or
// contract further.  The sequence is the one nvcc 12.9's default contraction produces for the upstream source
// form of ashawkey/diff-gaussian-rasterization:  for  X*Y (+|-) Z*W  the LEFT product
// is fused, the right one is rounded first; sums run left to right.
//
// The functions are __host__ __device__ so that tests/emul can run exactly this code on the CPU (test
// infrastructure only; -ffp-contract=off) and compare it bit for bit with the independent C oracle without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define LGM_HD __host__ __device__ __forceinline__
#else
#define LGM_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define LGM_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define LGM_MUL(a, b) __fmul_rn((a), (b))
#define LGM_ADD(a, b) __fadd_rn((a), (b))
#define LGM_SUB(a, b) __fsub_rn((a), (b))
#define LGM_DIV(a, b) __fdiv_rn((a), (b))
#define LGM_SQRT(a) __fsqrt_rn((a))
#define LGM_F2I(a) __float2int_rz((a))
#define LGM_DMUL(a, b) __dmul_rn((a), (b))
#define LGM_DADD(a, b) __dadd_rn((a), (b))
#else
#define LGM_FMA(a, b, c) fmaf((a), (b), (c))
#define LGM_MUL(a, b) ((float)(a) * (float)(b))
#define LGM_ADD(a, b) ((float)(a) + (float)(b))
#define LGM_SUB(a, b) ((float)(a) - (float)(b))
#define LGM_DIV(a, b) ((float)(a) / (float)(b))
#define LGM_SQRT(a) sqrtf((a))
static inline int lgm_host_f2i(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}
#define LGM_F2I(a) lgm_host_f2i((a))
#define LGM_DMUL(a, b) ((double)(a) * (double)(b))
#define LGM_DADD(a, b) ((double)(a) + (double)(b))
#endif

namespace lgm {

constexpr int kTile = 16;           // 16x16 pixel tiles (A.0)
constexpr float kNearCull = 0.2f;   // A.1
constexpr float kLowPass = 0.3f;    // A.1
constexpr float kAlphaMax = 0.99f;  // A.4
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kTEps = 0.0001f;
constexpr float kWEps = 0.0000001f;

struct Geom {       // what preprocess keeps per (view, Gaussian)
    float depth;    // p_view.z
    int radius;     // 0 = culled
    float px, py;   // pixel centre
    float cx, cy, cz;  // conic
    int rx0, ry0, rx1, ry1;  // tile rect [min, max)
    uint32_t tiles;
};

// m[i]*x + m[i+4]*y + m[i+8]*z + m[i+12]
LGM_HD float affine_row(const float* m, int i, float x, float y, float z)
{
    float t = LGM_MUL(m[i + 4], y);
    t = LGM_FMA(m[i], x, t);
    t = LGM_FMA(m[i + 8], z, t);
    return LGM_ADD(t, m[i + 12]);
}
// a0*b0 + a1*b1 + a2*b2
LGM_HD float dot3p(float a0, float b0, float a1, float b1, float a2, float b2)
{
    float t = LGM_MUL(a1, b1);
    t = LGM_FMA(a0, b0, t);
    return LGM_FMA(a2, b2, t);
}
LGM_HD int imin_(int a, int b) { return a < b ? a : b; }
LGM_HD int imax_(int a, int b) { return a > b ? a : b; }

// Rows of Rq*S (M[c][k] = s_k * Rq[c][k]) and the 6 unique entries of Sigma = (Rq S)(Rq S)^T.   A.1 cov3D
LGM_HD void cov3d_from_scale_rot(float sx, float sy, float sz, float mod, float r, float x, float y, float z,
                                 float* cov6, float* M /*9, row-major [c][k]*/)
{
    const float s[3] = {LGM_MUL(mod, sx), LGM_MUL(mod, sy), LGM_MUL(mod, sz)};
    float Rq[9];
    Rq[0] = LGM_SUB(1.0f, LGM_MUL(2.0f, LGM_FMA(y, y, LGM_MUL(z, z))));
    Rq[1] = LGM_MUL(2.0f, LGM_FMA(x, y, -LGM_MUL(r, z)));
    Rq[2] = LGM_MUL(2.0f, LGM_FMA(x, z, LGM_MUL(r, y)));
    Rq[3] = LGM_MUL(2.0f, LGM_FMA(x, y, LGM_MUL(r, z)));
    Rq[4] = LGM_SUB(1.0f, LGM_MUL(2.0f, LGM_FMA(x, x, LGM_MUL(z, z))));
    Rq[5] = LGM_MUL(2.0f, LGM_FMA(y, z, -LGM_MUL(r, x)));
    Rq[6] = LGM_MUL(2.0f, LGM_FMA(x, z, -LGM_MUL(r, y)));
    Rq[7] = LGM_MUL(2.0f, LGM_FMA(y, z, LGM_MUL(r, x)));
    Rq[8] = LGM_SUB(1.0f, LGM_MUL(2.0f, LGM_FMA(x, x, LGM_MUL(y, y))));
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int k = 0; k < 3; k++) M[3 * c + k] = LGM_MUL(s[k], Rq[3 * c + k]);
    cov6[0] = dot3p(M[0], M[0], M[1], M[1], M[2], M[2]);
    cov6[1] = dot3p(M[3], M[0], M[4], M[1], M[5], M[2]);
    cov6[2] = dot3p(M[6], M[0], M[7], M[1], M[8], M[2]);
    cov6[3] = dot3p(M[3], M[3], M[4], M[4], M[5], M[5]);
    cov6[4] = dot3p(M[6], M[3], M[7], M[4], M[8], M[5]);
    cov6[5] = dot3p(M[6], M[6], M[7], M[7], M[8], M[8]);
}

// EWA projection of the 3D covariance (A.1 cov2D).  Tm = J * W_view (2x3, row-major), t = clamped view point.
LGM_HD void cov2d_ewa(float pvx, float pvy, float pvz, float fx, float fy, float tanx, float tany,
                      const float* cov6, const float* mv, float* abc, float* Tm /*6*/, float* t /*3*/,
                      float* txtz_out, float* tytz_out)
{
    const float limx = LGM_MUL(1.3f, tanx), limy = LGM_MUL(1.3f, tany);
    const float txtz = LGM_DIV(pvx, pvz), tytz = LGM_DIV(pvy, pvz);
    const float tx = LGM_MUL(fminf(limx, fmaxf(-limx, txtz)), pvz);
    const float ty = LGM_MUL(fminf(limy, fmaxf(-limy, tytz)), pvz);
    const float tz = pvz;
    const float tz2 = LGM_MUL(tz, tz);
    const float J00 = LGM_DIV(fx, tz), J02 = LGM_DIV(-LGM_MUL(fx, tx), tz2);
    const float J11 = LGM_DIV(fy, tz), J12 = LGM_DIV(-LGM_MUL(fy, ty), tz2);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Tm[i] = LGM_FMA(mv[2 + 4 * i], J02, LGM_MUL(mv[0 + 4 * i], J00));
        Tm[3 + i] = LGM_FMA(mv[2 + 4 * i], J12, LGM_MUL(mv[1 + 4 * i], J11));
    }
    const float V[9] = {cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]};
    float A[6];  // A[j][i], i = 0,1
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 2; i++)
            A[2 * j + i] = dot3p(Tm[3 * i + 0], V[0 + j], Tm[3 * i + 1], V[3 + j], Tm[3 * i + 2], V[6 + j]);
    const float c00 = dot3p(A[0], Tm[0], A[2], Tm[1], A[4], Tm[2]);
    const float c01 = dot3p(A[1], Tm[0], A[3], Tm[1], A[5], Tm[2]);
    const float c11 = dot3p(A[1], Tm[3], A[3], Tm[4], A[5], Tm[5]);
    abc[0] = LGM_ADD(c00, kLowPass);
    abc[1] = c01;
    abc[2] = LGM_ADD(c11, kLowPass);
    t[0] = tx; t[1] = ty; t[2] = tz;
    *txtz_out = txtz;
    *tytz_out = tytz;
}

// upstream ndc2Pix: ((v + 1.0) * S - 1.0) * 0.5 with double literals, narrowed to float once.
LGM_HD float ndc2pix(float v, int S)
{
    double d = LGM_DADD((double)v, 1.0);
    d = LGM_DMUL(d, (double)S);
    d = LGM_DADD(d, -1.0);
    d = LGM_DMUL(d, 0.5);
    return (float)d;
}

LGM_HD void tile_rect(float px, float py, int radius, int gx, int gy, int& x0, int& y0, int& x1, int& y1)
{
    // upstream divides by BLOCK_X = 16; a correctly rounded division by a power of two IS the multiplication by its
    // reciprocal (exact scaling, no underflow for pixel coordinates), at a fraction of the instructions — this function
    // runs in K1, the count, the scatter and the emit kernels
    const float r = (float)radius;
    constexpr float kInvTile = 1.0f / 16.0f;
    x0 = imin_(gx, imax_(0, LGM_F2I(LGM_MUL(LGM_SUB(px, r), kInvTile))));
    y0 = imin_(gy, imax_(0, LGM_F2I(LGM_MUL(LGM_SUB(py, r), kInvTile))));
    x1 = imin_(gx, imax_(0, LGM_F2I(LGM_MUL(LGM_SUB(LGM_ADD(LGM_ADD(px, r), 16.0f), 1.0f), kInvTile))));
    y1 = imin_(gy, imax_(0, LGM_F2I(LGM_MUL(LGM_SUB(LGM_ADD(LGM_ADD(py, r), 16.0f), 1.0f), kInvTile))));
}

// A.1 for one (view, Gaussian).  g = the 14 floats of /root/reference/core/gs.py:45-49 minus colour:
// pos(3) opacity(1) scale(3) rot(4).  Returns a zeroed Geom (radius 0) when culled.
// cov6_in (optional): a precomputed 3D covariance (xx, xy, xz, yy, yz, zz) — upstream's cov3D_precomp — used instead of
// the one built from scale / rotation / mod.
LGM_HD Geom preprocess_point(const float* pos, const float* scale, const float* rot, float mod, const float* mv,
                             const float* mp, int W, int H, float tanx, float tany, float fx, float fy, int gx, int gy,
                             const float* cov6_in = nullptr)
{
    Geom o;
    o.depth = 0.f; o.radius = 0; o.px = o.py = 0.f; o.cx = o.cy = o.cz = 0.f;
    o.rx0 = o.ry0 = o.rx1 = o.ry1 = 0; o.tiles = 0;
    const float x = pos[0], y = pos[1], z = pos[2];
    const float pvx = affine_row(mv, 0, x, y, z), pvy = affine_row(mv, 1, x, y, z), pvz = affine_row(mv, 2, x, y, z);
    if (pvz <= kNearCull) return o;
    const float hx = affine_row(mp, 0, x, y, z), hy = affine_row(mp, 1, x, y, z), hw = affine_row(mp, 3, x, y, z);
    const float pw = LGM_DIV(1.0f, LGM_ADD(hw, kWEps));
    const float projx = LGM_MUL(hx, pw), projy = LGM_MUL(hy, pw);
    float cov6[6], M[9], abc[3], Tm[6], t[3], txtz, tytz;
    if (cov6_in) {
#pragma unroll
        for (int k = 0; k < 6; k++) cov6[k] = cov6_in[k];
    } else {
        cov3d_from_scale_rot(scale[0], scale[1], scale[2], mod, rot[0], rot[1], rot[2], rot[3], cov6, M);
    }
    cov2d_ewa(pvx, pvy, pvz, fx, fy, tanx, tany, cov6, mv, abc, Tm, t, &txtz, &tytz);
    const float a = abc[0], b = abc[1], c = abc[2];
    const float det = LGM_FMA(a, c, -LGM_MUL(b, b));
    if (det == 0.0f) return o;
    const float det_inv = LGM_DIV(1.0f, det);
    const float mid = LGM_MUL(0.5f, LGM_ADD(a, c));
    const float sq = LGM_SQRT(fmaxf(0.1f, LGM_FMA(mid, mid, -det)));
    const float l1 = LGM_ADD(mid, sq), l2 = LGM_SUB(mid, sq);
    const int rad = LGM_F2I(ceilf(LGM_MUL(3.0f, LGM_SQRT(fmaxf(l1, l2)))));
    const float px = ndc2pix(projx, W), py = ndc2pix(projy, H);
    int x0, y0, x1, y1;
    tile_rect(px, py, rad, gx, gy, x0, y0, x1, y1);
    const int area = (x1 - x0) * (y1 - y0);
    if (area == 0) return o;
    o.depth = pvz; o.radius = rad; o.px = px; o.py = py;
    o.cx = LGM_MUL(c, det_inv); o.cy = LGM_MUL(-b, det_inv); o.cz = LGM_MUL(a, det_inv);
    o.rx0 = x0; o.ry0 = y0; o.rx1 = x1; o.ry1 = y1; o.tiles = (uint32_t)area;
    return o;
}

// A.4 / A.5: exponent of the Gaussian at offset d = xy - pixel.
LGM_HD float pair_power(float cx, float cy, float cz, float dx, float dy)
{
    const float s = LGM_FMA(LGM_MUL(cx, dx), dx, LGM_MUL(LGM_MUL(cz, dy), dy));
    return LGM_FMA(s, -0.5f, -LGM_MUL(LGM_MUL(cy, dx), dy));
}

// ---------------------------------------------------------------------------------------------------------------
// A.6 preprocess backward for one (view, Gaussian) with radius > 0.  Tolerance-checked (1e-3 rel), not bit-pinned:
// natural expression form.  g2 = dL/dmean2D (NDC-scaled, 2), gc = dL/dconic (x, y, w slots), gd = dL/ddepth.
// Accumulates (+=) into dpos[3], dscale[3], drot[4].
// The compositing backward accumulates, per (view, Gaussian), the moments of q = G dL/dalpha about the Gaussian's
// centre over its pixels (d = centre - pixel):  Sx = sum q dx, Sy, Sxx = sum q dx dx, Sxy, Syy  (and S0 = sum q, which IS
// dL/dopacity).  Upstream's per-pixel terms (A.5) are these moments times per-Gaussian coefficients:
//   dL/dmean2D.x = -(W/2) o (cxx Sx + cxy Sy)     dL/dconic.xx = -o Sxx / 2
//   dL/dmean2D.y = -(H/2) o (cyy Sy + cxy Sx)     dL/dconic.xy = -o Sxy / 2      dL/dconic.yy = -o Syy / 2
// with (cxx, cxy, cyy, o) = the Gaussian's conic and opacity.
LGM_HD void moments_to_gradients(float W, float H, float cxx, float cxy, float cyy, float o, float Sx, float Sy, float Sxx,
                                 float Sxy, float Syy, float* g2x, float* g2y, float* gcx, float* gcy, float* gcz)
{
    const float h = -0.5f * o;
    *g2x = h * W * (cxx * Sx + cxy * Sy);
    *g2y = h * H * (cyy * Sy + cxy * Sx);
    *gcx = h * Sxx;
    *gcy = h * Sxy;
    *gcz = h * Syy;
}

// Preprocess backward, split so that a caller summing over the views of a scene (K7) pays the view-independent part
// once: `_view` is everything that depends on the camera — (i) conic -> cov2D -> cov3D gradient g6 and the mean through
// the EWA Jacobian, (ii) projection, (iii) depth — and ACCUMULATES dL/dpos and g6 = dL/dcov3D; `_finish` maps the summed
// g6 to scale / rotation ((v): linear in g6 for a fixed scale and rotation).
// moments = false: (g2x, g2y) = dL/dmean2D, (gcx, gcy, gcz) = dL/dconic as upstream passes them.
// moments = true : the same five arguments carry (Sx, Sy, Sxx, Sxy, Syy) and are converted here, with the conic
// re-derived from the 2D covariance this function computes anyway (W, H = image size, opacity = the Gaussian's).
LGM_HD void preprocess_point_bwd_view(const float* pos, const float* cov6, const float* mv, const float* mp, float tanx,
                                      float tany, float fx, float fy, float g2x, float g2y, float gcx, float gcy, float gcz,
                                      float gd, float* dpos, float* g6acc, bool moments, float W, float H, float opacity)
{
    const float x = pos[0], y = pos[1], z = pos[2];
    const float pvx = affine_row(mv, 0, x, y, z), pvy = affine_row(mv, 1, x, y, z), pvz = affine_row(mv, 2, x, y, z);
    float abc[3], Tm[6], t[3], txtz, tytz;
    cov2d_ewa(pvx, pvy, pvz, fx, fy, tanx, tany, cov6, mv, abc, Tm, t, &txtz, &tytz);
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float xg = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float yg = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float a = abc[0], b = abc[1], c = abc[2];
    const float denom = a * c - b * b;
    if (moments) {
        const float di = 1.0f / denom;  // the forward's conic = (c, -b, a) / det
        moments_to_gradients(W, H, c * di, -b * di, a * di, opacity, g2x, g2y, gcx, gcy, gcz, &g2x, &g2y, &gcx, &gcy, &gcz);
    }
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    const float denom2inv = 1.0f / ((denom * denom) + kWEps);
    float g6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* T0 = Tm;
    const float* T1 = Tm + 3;
    if (denom2inv != 0.f) {
        dL_da = denom2inv * (-c * c * gcx + 2.f * b * c * gcy + (denom - a * c) * gcz);
        dL_dc = denom2inv * (-a * a * gcz + 2.f * a * b * gcy + (denom - a * c) * gcx);
        dL_db = denom2inv * 2.f * (b * c * gcx - (denom + 2.f * b * b) * gcy + a * b * gcz);
        g6[0] = T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
        g6[3] = T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
        g6[5] = T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
        g6[1] = 2.f * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2.f * T1[0] * T1[1] * dL_dc;
        g6[2] = 2.f * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2.f * T1[0] * T1[2] * dL_dc;
        g6[4] = 2.f * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2.f * T1[1] * T1[2] * dL_dc;
    }
    const float V[9] = {cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]};
    float dT0[3], dT1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float tv0 = T0[0] * V[3 * k] + T0[1] * V[3 * k + 1] + T0[2] * V[3 * k + 2];
        const float tv1 = T1[0] * V[3 * k] + T1[1] * V[3 * k + 1] + T1[2] * V[3 * k + 2];
        dT0[k] = 2.f * tv0 * dL_da + tv1 * dL_db;
        dT1[k] = 2.f * tv1 * dL_dc + tv0 * dL_db;
    }
    const float dJ00 = mv[0] * dT0[0] + mv[4] * dT0[1] + mv[8] * dT0[2];
    const float dJ02 = mv[2] * dT0[0] + mv[6] * dT0[1] + mv[10] * dT0[2];
    const float dJ11 = mv[1] * dT1[0] + mv[5] * dT1[1] + mv[9] * dT1[2];
    const float dJ12 = mv[2] * dT1[0] + mv[6] * dT1[1] + mv[10] * dT1[2];
    const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = xg * -fx * tz2 * dJ02;
    const float dty = yg * -fy * tz2 * dJ12;
    const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * t[0]) * tz3 * dJ02 + (2.f * fy * t[1]) * tz3 * dJ12;
    float dm0 = mv[0] * dtx + mv[1] * dty + mv[2] * dtz;
    float dm1 = mv[4] * dtx + mv[5] * dty + mv[6] * dtz;
    float dm2 = mv[8] * dtx + mv[9] * dty + mv[10] * dtz;
    // (ii) projection
    const float hw = affine_row(mp, 3, x, y, z);
    const float m_w = 1.0f / (hw + kWEps);
    const float mul1 = (mp[0] * x + mp[4] * y + mp[8] * z + mp[12]) * m_w * m_w;
    const float mul2 = (mp[1] * x + mp[5] * y + mp[9] * z + mp[13]) * m_w * m_w;
    dm0 += (mp[0] * m_w - mp[3] * mul1) * g2x + (mp[1] * m_w - mp[3] * mul2) * g2y;
    dm1 += (mp[4] * m_w - mp[7] * mul1) * g2x + (mp[5] * m_w - mp[7] * mul2) * g2y;
    dm2 += (mp[8] * m_w - mp[11] * mul1) * g2x + (mp[9] * m_w - mp[11] * mul2) * g2y;
    // (iii) depth
    const float mul3 = mv[2] * x + mv[6] * y + mv[10] * z + mv[14];
    dm0 += (mv[2] - mv[3] * mul3) * gd;
    dm1 += (mv[6] - mv[7] * mul3) * gd;
    dm2 += (mv[10] - mv[11] * mul3) * gd;
    dpos[0] += dm0; dpos[1] += dm1; dpos[2] += dm2;
#pragma unroll
    for (int k = 0; k < 6; k++) g6acc[k] += g6[k];
}

LGM_HD void preprocess_point_bwd_finish(const float* scale, const float* rot, float mod, const float* g6, float* dscale,
                                        float* drot)
{
    float cov6[6], M[9];
    cov3d_from_scale_rot(scale[0], scale[1], scale[2], mod, rot[0], rot[1], rot[2], rot[3], cov6, M);
    // (v) cov3D -> scale / rotation.  dA = 2 * Gsym * (Rq S)
    const float Gs[9] = {g6[0], 0.5f * g6[1], 0.5f * g6[2], 0.5f * g6[1], g6[3], 0.5f * g6[4], 0.5f * g6[2], 0.5f * g6[4], g6[5]};
    float dA[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++)
            dA[3 * i + k] = 2.f * (Gs[3 * i] * M[k] + Gs[3 * i + 1] * M[3 + k] + Gs[3 * i + 2] * M[6 + k]);
    const float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
    const float r = rot[0], qx = rot[1], qy = rot[2], qz = rot[3];
    float Rq[9];
    Rq[0] = 1.f - 2.f * (qy * qy + qz * qz); Rq[1] = 2.f * (qx * qy - r * qz); Rq[2] = 2.f * (qx * qz + r * qy);
    Rq[3] = 2.f * (qx * qy + r * qz); Rq[4] = 1.f - 2.f * (qx * qx + qz * qz); Rq[5] = 2.f * (qy * qz - r * qx);
    Rq[6] = 2.f * (qx * qz - r * qy); Rq[7] = 2.f * (qy * qz + r * qx); Rq[8] = 1.f - 2.f * (qx * qx + qy * qy);
    // upstream quirk kept: no `mod` factor on dL_dscale (A.6 (v))
#pragma unroll
    for (int k = 0; k < 3; k++) dscale[k] += Rq[k] * dA[k] + Rq[3 + k] * dA[3 + k] + Rq[6 + k] * dA[6 + k];
    float dR[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) dR[3 * i + k] = dA[3 * i + k] * s[k];
#define LGM_DMT(c, rr) dR[3 * (rr) + (c)]
    drot[0] += 2.f * qz * (LGM_DMT(0, 1) - LGM_DMT(1, 0)) + 2.f * qy * (LGM_DMT(2, 0) - LGM_DMT(0, 2)) + 2.f * qx * (LGM_DMT(1, 2) - LGM_DMT(2, 1));
    drot[1] += 2.f * qy * (LGM_DMT(1, 0) + LGM_DMT(0, 1)) + 2.f * qz * (LGM_DMT(2, 0) + LGM_DMT(0, 2)) + 2.f * r * (LGM_DMT(1, 2) - LGM_DMT(2, 1)) - 4.f * qx * (LGM_DMT(2, 2) + LGM_DMT(1, 1));
    drot[2] += 2.f * qx * (LGM_DMT(1, 0) + LGM_DMT(0, 1)) + 2.f * r * (LGM_DMT(2, 0) - LGM_DMT(0, 2)) + 2.f * qz * (LGM_DMT(1, 2) + LGM_DMT(2, 1)) - 4.f * qy * (LGM_DMT(2, 2) + LGM_DMT(0, 0));
    drot[3] += 2.f * r * (LGM_DMT(0, 1) - LGM_DMT(1, 0)) + 2.f * qx * (LGM_DMT(2, 0) + LGM_DMT(0, 2)) + 2.f * qy * (LGM_DMT(1, 2) + LGM_DMT(2, 1)) - 4.f * qz * (LGM_DMT(1, 1) + LGM_DMT(0, 0));
#undef LGM_DMT
}

LGM_HD void preprocess_point_bwd_core(const float* pos, const float* scale, const float* rot, float mod, const float* mv,
                                      const float* mp, float tanx, float tany, float fx, float fy, float g2x, float g2y,
                                      float gcx, float gcy, float gcz, float gd, float* dpos, float* dscale, float* drot,
                                      bool moments, float W, float H, float opacity)
{
    float cov6[6], M[9], g6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    cov3d_from_scale_rot(scale[0], scale[1], scale[2], mod, rot[0], rot[1], rot[2], rot[3], cov6, M);
    preprocess_point_bwd_view(pos, cov6, mv, mp, tanx, tany, fx, fy, g2x, g2y, gcx, gcy, gcz, gd, dpos, g6, moments, W,
                              H, opacity);
    preprocess_point_bwd_finish(scale, rot, mod, g6, dscale, drot);
}

LGM_HD void preprocess_point_bwd(const float* pos, const float* scale, const float* rot, float mod, const float* mv,
                                 const float* mp, float tanx, float tany, float fx, float fy, float g2x, float g2y,
                                 float gcx, float gcy, float gcz, float gd, float* dpos, float* dscale, float* drot)
{
    preprocess_point_bwd_core(pos, scale, rot, mod, mv, mp, tanx, tany, fx, fy, g2x, g2y, gcx, gcy, gcz, gd, dpos, dscale, drot,
                              false, 0.f, 0.f, 0.f);
}

LGM_HD void preprocess_point_bwd_moments(const float* pos, const float* scale, const float* rot, float mod, const float* mv,
                                         const float* mp, float tanx, float tany, float fx, float fy, float W, float H,
                                         float opacity, float Sx, float Sy, float Sxx, float Sxy, float Syy, float gd,
                                         float* dpos, float* dscale, float* drot)
{
    preprocess_point_bwd_core(pos, scale, rot, mod, mv, mp, tanx, tany, fx, fy, Sx, Sy, Sxx, Sxy, Syy, gd, dpos, dscale, drot,
                              true, W, H, opacity);
}

}  // namespace lgm
