/*
 * splat_oracle.c — CPU restatement of the Gaussian-splat rasterizer path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this.
 * The product (lgm_b200/) never imports, links or calls anything in oracle/.
 *
 * PARITY UNPINNED.  The arithmetic of this path lives in the third-party package
 * ashawkey/diff-gaussian-rasterization (unpinned default branch, /root/reference/readme.md:13-15), whose
 * source is NOT under /root/reference and not installed; the reference holds no tests, golden vectors or
 * fixtures for it.  This file restates the published algorithm as specified in
 * SURVEY.md Appendix A (A.1 .. A.6) and is anchored on the reference's own call site
 * /root/reference/core/gs.py:45-49 (14-channel split), :58-71 (settings), :76-85 (call), :87 (clamp).
 *
 * Built twice by oracle/Makefile:
 *   liboracle_f32.so  real = float.  Every expression that feeds radii / tile rects / sort keys / the
 *                     alpha thresholds is written with an explicit fmaf()/mul/add sequence — the sequence
 *                     nvcc's default contraction (-fmad=true) produces for the upstream source form
 *                     (checked with nvcc 12.9 -> SASS on small probes, see DESIGN.md "arithmetic contract").
 *                     Compiled with -ffp-contract=off so gcc adds no fusion of its own.  The CUDA kernels
 *                     pin the same sequence with __fmaf_rn/__fmul_rn/__fadd_rn, so the integer / bit-level
 *                     outputs (radii, xy bits, depth bits, tiles_touched, keys, sorted order, ranges) compare
 *                     bit-exactly.  expf differs between glibc and the GPU by <= 1-2 ulp: compositing is
 *                     compared with a tolerance.
 *   liboracle_f64.so  real = double (-DORACLE_F64).  Same algorithm, same float-literal constants, no pinning;
 *                     the numerical arbiter for gradients (finite differences + 1e-3 relative checks).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_F64
typedef double real;
#define FMA(a, b, c) ((a) * (b) + (c))
#define SQRT(x) sqrt(x)
#define EXP(x) exp(x)
#define CEIL(x) ceil(x)
#define FMAX(a, b) fmax(a, b)
#define FMIN(a, b) fmin(a, b)
#else
typedef float real;
#define FMA(a, b, c) fmaf((a), (b), (c))
#define SQRT(x) sqrtf(x)
#define EXP(x) expf(x)
#define CEIL(x) ceilf(x)
#define FMAX(a, b) fmaxf(a, b)
#define FMIN(a, b) fminf(a, b)
#endif
#define R(x) ((real)(x))

#define TILE 16
#define NEAR_CULL R(0.2f)        /* A.1: cull if p_view.z <= 0.2 */
#define LOWPASS R(0.3f)          /* A.1: cov[0][0] += 0.3, cov[1][1] += 0.3 */
#define ALPHA_MAX R(0.99f)       /* A.4 */
#define ALPHA_MIN R(1.0f / 255.0f)
#define T_EPS R(0.0001f)
#define W_EPS R(0.0000001f)

/* CUDA float->int conversion (cvt.rzi.s32.f32): truncation, saturating, NaN -> 0. */
static int f2i_rz(real v)
{
    if (v != v) return 0;
    if (v >= R(2147483648.0)) return INT_MAX;
    if (v <= R(-2147483648.0)) return INT_MIN;
    return (int)v;
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* A.1 "row i of the transform": m[i]*x + m[i+4]*y + m[i+8]*z + m[i+12]
 * nvcc: t = m4*y ; t = fma(m0,x,t) ; t = fma(m8,z,t) ; t = t + m12                                   */
static real affine_row(const real *m, int i, real x, real y, real z)
{
    real t = m[i + 4] * y;
    t = FMA(m[i], x, t);
    t = FMA(m[i + 8], z, t);
    return t + m[i + 12];
}
/* a0*b0 + a1*b1 + a2*b2  ->  fma(a2,b2, fma(a0,b0, a1*b1))                                           */
static real dot3p(real a0, real b0, real a1, real b1, real a2, real b2)
{
    real t = a1 * b1;
    t = FMA(a0, b0, t);
    return FMA(a2, b2, t);
}

/* A.1 cov3D: rows of (R_q * S), Sigma = (R_q S)(R_q S)^T, 6 upper-triangular entries.
 * rot = (r,x,y,z) used as given (no normalisation).                                                  */
static void cov3d_from_scale_rot(const real *s3, real mod, const real *q, real *cov6, real Mrow[3][3])
{
    real s[3] = {mod * s3[0], mod * s3[1], mod * s3[2]};
    real r = q[0], x = q[1], y = q[2], z = q[3];
    real Rq[3][3];
    /* 1 - 2(y^2+z^2): fma(y,y,z*z) ; *2 ; 1 - .   |  2(xy - rz): fma(x,y,-(r*z)) ; *2                 */
    Rq[0][0] = R(1.0f) - R(2.0f) * FMA(y, y, z * z);
    Rq[0][1] = R(2.0f) * FMA(x, y, -(r * z));
    Rq[0][2] = R(2.0f) * FMA(x, z, r * y);
    Rq[1][0] = R(2.0f) * FMA(x, y, r * z);
    Rq[1][1] = R(1.0f) - R(2.0f) * FMA(x, x, z * z);
    Rq[1][2] = R(2.0f) * FMA(y, z, -(r * x));
    Rq[2][0] = R(2.0f) * FMA(x, z, -(r * y));
    Rq[2][1] = R(2.0f) * FMA(y, z, r * x);
    Rq[2][2] = R(1.0f) - R(2.0f) * FMA(x, x, y * y);
    for (int c = 0; c < 3; c++)
        for (int k = 0; k < 3; k++) Mrow[c][k] = s[k] * Rq[c][k];
    /* Sigma[i][j] = M_i0*M_j0 + M_i1*M_j1 + M_i2*M_j2 */
    cov6[0] = dot3p(Mrow[0][0], Mrow[0][0], Mrow[0][1], Mrow[0][1], Mrow[0][2], Mrow[0][2]);
    cov6[1] = dot3p(Mrow[1][0], Mrow[0][0], Mrow[1][1], Mrow[0][1], Mrow[1][2], Mrow[0][2]);
    cov6[2] = dot3p(Mrow[2][0], Mrow[0][0], Mrow[2][1], Mrow[0][1], Mrow[2][2], Mrow[0][2]);
    cov6[3] = dot3p(Mrow[1][0], Mrow[1][0], Mrow[1][1], Mrow[1][1], Mrow[1][2], Mrow[1][2]);
    cov6[4] = dot3p(Mrow[2][0], Mrow[1][0], Mrow[2][1], Mrow[1][1], Mrow[2][2], Mrow[1][2]);
    cov6[5] = dot3p(Mrow[2][0], Mrow[2][0], Mrow[2][1], Mrow[2][1], Mrow[2][2], Mrow[2][2]);
}

/* A.1 cov2D.  Returns (a,b,c) with the +0.3 low-pass applied; also the 2x3 matrix Tm = J * W_view
 * (Tm[0][i], Tm[1][i]) and the clamped t, for the backward.                                         */
static void cov2d_ewa(const real *pv, real fx, real fy, real tanx, real tany, const real *cov6, const real *mv,
                      real *abc, real Tm[2][3], real *tclamped, real *txtz_out, real *tytz_out)
{
    real tx = pv[0], ty = pv[1], tz = pv[2];
    real limx = R(1.3f) * tanx, limy = R(1.3f) * tany;
    real txtz = tx / tz, tytz = ty / tz;
    tx = FMIN(limx, FMAX(-limx, txtz)) * tz;
    ty = FMIN(limy, FMAX(-limy, tytz)) * tz;
    real J00 = fx / tz, J02 = -(fx * tx) / (tz * tz);
    real J11 = fy / tz, J12 = -(fy * ty) / (tz * tz);
    /* W_view[k][i] = mv[k + 4 i];  Tm[0][i] = fma(W[2][i], J02, W[0][i]*J00);  Tm[1][i] = fma(W[2][i], J12, W[1][i]*J11) */
    for (int i = 0; i < 3; i++) {
        Tm[0][i] = FMA(mv[2 + 4 * i], J02, mv[0 + 4 * i] * J00);
        Tm[1][i] = FMA(mv[2 + 4 * i], J12, mv[1 + 4 * i] * J11);
    }
    real V[3][3] = {{cov6[0], cov6[1], cov6[2]}, {cov6[1], cov6[3], cov6[4]}, {cov6[2], cov6[4], cov6[5]}};
    /* A[j][i] = T[i][0] V[0][j] + T[i][1] V[1][j] + T[i][2] V[2][j]   (i = 0,1) */
    real A[3][2];
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 2; i++) A[j][i] = dot3p(Tm[i][0], V[0][j], Tm[i][1], V[1][j], Tm[i][2], V[2][j]);
    /* cov[j][i] = A[0][i] T[j][0] + A[1][i] T[j][1] + A[2][i] T[j][2] */
    real c00 = dot3p(A[0][0], Tm[0][0], A[1][0], Tm[0][1], A[2][0], Tm[0][2]);
    real c01 = dot3p(A[0][1], Tm[0][0], A[1][1], Tm[0][1], A[2][1], Tm[0][2]);
    real c11 = dot3p(A[0][1], Tm[1][0], A[1][1], Tm[1][1], A[2][1], Tm[1][2]);
    abc[0] = c00 + LOWPASS;
    abc[1] = c01;
    abc[2] = c11 + LOWPASS;
    if (tclamped) { tclamped[0] = tx; tclamped[1] = ty; tclamped[2] = tz; }
    if (txtz_out) *txtz_out = txtz;
    if (tytz_out) *tytz_out = tytz;
}

/* upstream ndc2Pix uses double literals: ((v + 1.0) * S - 1.0) * 0.5 evaluated in double, then narrowed. */
static real ndc2pix(real v, int S) { return (real)((((double)v + 1.0) * (double)S - 1.0) * 0.5); }

static void tile_rect(real px, real py, int radius, int gx, int gy, int *rect)
{
    real r = (real)radius;
    rect[0] = imin(gx, imax(0, f2i_rz((px - r) / R(16.0f))));
    rect[1] = imin(gy, imax(0, f2i_rz((py - r) / R(16.0f))));
    rect[2] = imin(gx, imax(0, f2i_rz((((px + r) + R(16.0f)) - R(1.0f)) / R(16.0f))));
    rect[3] = imin(gy, imax(0, f2i_rz((((py + r) + R(16.0f)) - R(1.0f)) / R(16.0f))));
}

/* ------------------------------------------------------------------------------------------------ */
/* A.1 preprocess, one view.  Outputs are zero for culled Gaussians (radii = 0, tiles = 0).          */
void orc_preprocess(int P, const real *means, const real *scales, const real *rots, const real *opac, real mod,
                    const real *mv, const real *mp, int W, int H, real tanx, real tany, real *depth,
                    int32_t *radii, real *xy, real *conic_opacity, uint32_t *tiles, real *cov3d, int32_t *rects)
{
    const real fx = (real)W / (R(2.0f) * tanx), fy = (real)H / (R(2.0f) * tany);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    for (int idx = 0; idx < P; idx++) {
        radii[idx] = 0;
        tiles[idx] = 0;
        depth[idx] = 0;
        xy[2 * idx] = xy[2 * idx + 1] = 0;
        for (int k = 0; k < 4; k++) conic_opacity[4 * idx + k] = 0;
        if (rects) for (int k = 0; k < 4; k++) rects[4 * idx + k] = 0;
        real x = means[3 * idx], y = means[3 * idx + 1], z = means[3 * idx + 2];
        real pv[3] = {affine_row(mv, 0, x, y, z), affine_row(mv, 1, x, y, z), affine_row(mv, 2, x, y, z)};
        real Mrow[3][3], c6[6];
        cov3d_from_scale_rot(scales + 3 * idx, mod, rots + 4 * idx, c6, Mrow);
        if (cov3d) memcpy(cov3d + 6 * idx, c6, sizeof(c6));
        if (pv[2] <= NEAR_CULL) continue; /* A.1 near cull, written as upstream's "<=" (a NaN depth is NOT culled here) */
        real hx = affine_row(mp, 0, x, y, z), hy = affine_row(mp, 1, x, y, z), hw = affine_row(mp, 3, x, y, z);
        real pw = R(1.0f) / (hw + W_EPS);
        real projx = hx * pw, projy = hy * pw;
        real abc[3], Tm[2][3];
        cov2d_ewa(pv, fx, fy, tanx, tany, c6, mv, abc, Tm, NULL, NULL, NULL);
        real a = abc[0], b = abc[1], c = abc[2];
        real det = FMA(a, c, -(b * b));
        if (det == R(0.0f)) continue;
        real det_inv = R(1.0f) / det;
        real con[3] = {c * det_inv, -b * det_inv, a * det_inv};
        real mid = R(0.5f) * (a + c);
        real sq = SQRT(FMAX(R(0.1f), FMA(mid, mid, -det)));
        real l1 = mid + sq, l2 = mid - sq;
        real rad_f = CEIL(R(3.0f) * SQRT(FMAX(l1, l2)));
        int rad = f2i_rz(rad_f);
        real px = ndc2pix(projx, W), py = ndc2pix(projy, H);
        int rect[4];
        tile_rect(px, py, rad, gx, gy, rect);
        int area = (rect[2] - rect[0]) * (rect[3] - rect[1]);
        if (area == 0) continue;
        depth[idx] = pv[2];
        radii[idx] = rad;
        xy[2 * idx] = px;
        xy[2 * idx + 1] = py;
        conic_opacity[4 * idx + 0] = con[0];
        conic_opacity[4 * idx + 1] = con[1];
        conic_opacity[4 * idx + 2] = con[2];
        conic_opacity[4 * idx + 3] = opac[idx];
        tiles[idx] = (uint32_t)area;
        if (rects) memcpy(rects + 4 * idx, rect, sizeof(rect));
    }
}

/* markVisible: p_view.z > 0.2 (upstream checkFrustum / in_frustum) */
void orc_mark_visible(int P, const real *means, const real *mv, uint8_t *visible)
{
    for (int i = 0; i < P; i++)
        visible[i] = !(affine_row(mv, 2, means[3 * i], means[3 * i + 1], means[3 * i + 2]) <= NEAR_CULL);
}

/* ------------------------------------------------------------------------------------------------ */
/* A.2 emit + stable sort + A.3 ranges, one view.                                                    */
typedef struct { uint32_t tile; uint32_t idx; real depth; uint32_t dbits; } inst_t;

static int inst_cmp(const void *pa, const void *pb)
{
    const inst_t *a = (const inst_t *)pa, *b = (const inst_t *)pb;
    if (a->tile != b->tile) return a->tile < b->tile ? -1 : 1;
#ifdef ORACLE_F64
    if (a->depth != b->depth) return a->depth < b->depth ? -1 : 1;
#else
    if (a->dbits != b->dbits) return a->dbits < b->dbits ? -1 : 1; /* radix order on the float's bit pattern */
#endif
    if (a->idx != b->idx) return a->idx < b->idx ? -1 : 1; /* stable: emit order is ascending idx within a tile */
    return 0;
}

int64_t orc_count_instances(int P, const uint32_t *tiles)
{
    int64_t L = 0;
    for (int i = 0; i < P; i++) L += tiles[i];
    return L;
}

/* keys[L] (tile<<32 | float bits of depth), vals[L] (Gaussian idx), ranges[2*ntiles] (start,end).
 * unsorted_keys/unsorted_vals (optional) receive the emit order.                                    */
void orc_bin(int P, const int32_t *radii, const real *xy, const real *depth, int W, int H, int64_t L,
             uint64_t *keys, uint32_t *vals, uint32_t *ranges, uint64_t *unsorted_keys, uint32_t *unsorted_vals)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    inst_t *inst = (inst_t *)malloc(sizeof(inst_t) * (size_t)(L > 0 ? L : 1));
    int64_t n = 0;
    for (int idx = 0; idx < P; idx++) {
        if (radii[idx] <= 0) continue;
        int rect[4];
        tile_rect(xy[2 * idx], xy[2 * idx + 1], radii[idx], gx, gy, rect);
        float df = (float)depth[idx];
        uint32_t bits;
        memcpy(&bits, &df, 4);
        for (int y = rect[1]; y < rect[3]; y++)
            for (int x = rect[0]; x < rect[2]; x++) {
                inst[n].tile = (uint32_t)(y * gx + x);
                inst[n].idx = (uint32_t)idx;
                inst[n].depth = depth[idx];
                inst[n].dbits = bits;
                if (unsorted_keys) unsorted_keys[n] = ((uint64_t)inst[n].tile << 32) | bits;
                if (unsorted_vals) unsorted_vals[n] = (uint32_t)idx;
                n++;
            }
    }
    qsort(inst, (size_t)n, sizeof(inst_t), inst_cmp);
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)(gx * gy));
    for (int64_t i = 0; i < n; i++) {
        keys[i] = ((uint64_t)inst[i].tile << 32) | inst[i].dbits;
        vals[i] = inst[i].idx;
        uint32_t c = inst[i].tile;
        if (i == 0) ranges[2 * c] = 0;
        else if (c != inst[i - 1].tile) { ranges[2 * inst[i - 1].tile + 1] = (uint32_t)i; ranges[2 * c] = (uint32_t)i; }
        if (i == n - 1) ranges[2 * c + 1] = (uint32_t)n;
    }
    free(inst);
}

/* ------------------------------------------------------------------------------------------------ */
/* A.4 forward compositing, one view.                                                                */
void orc_composite_fwd(int W, int H, const uint32_t *ranges, const uint32_t *vals, const real *xy,
                       const real *conic_opacity, const real *colors, const real *depth, const real *bg,
                       real *image, real *alpha_img, real *depth_img, uint32_t *n_contrib)
{
    const int gx = (W + TILE - 1) / TILE;
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            int tile = (py / TILE) * gx + px / TILE;
            uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
            real pfx = (real)px, pfy = (real)py;
            real T = R(1.0f), C[3] = {0, 0, 0}, Wt = 0, D = 0;
            uint32_t contributor = 0, last = 0;
            for (uint32_t k = r0; k < r1; k++) {
                contributor++;
                uint32_t id = vals[k];
                real dx = xy[2 * id] - pfx, dy = xy[2 * id + 1] - pfy;
                const real *co = conic_opacity + 4 * id;
                /* power = -0.5f*(cx*dx*dx + cz*dy*dy) - cy*dx*dy
                 * nvcc: s = fma(cx*dx, dx, (cz*dy)*dy) ; power = fma(s, -0.5, -((cy*dx)*dy))          */
                real s = FMA(co[0] * dx, dx, (co[2] * dy) * dy);
                real power = FMA(s, R(-0.5f), -((co[1] * dx) * dy));
                if (power > R(0.0f)) continue;
                real alpha = FMIN(ALPHA_MAX, co[3] * EXP(power));
                if (alpha < ALPHA_MIN) continue;
                real test_T = T * (R(1.0f) - alpha);
                if (test_T < T_EPS) break; /* done: this Gaussian is NOT composited */
                for (int ch = 0; ch < 3; ch++) C[ch] = FMA(colors[3 * id + ch] * alpha, T, C[ch]);
                Wt = FMA(alpha, T, Wt);
                D = FMA(depth[id] * alpha, T, D);
                T = test_T;
                last = contributor;
            }
            int pix = py * W + px;
            n_contrib[pix] = last;
            for (int ch = 0; ch < 3; ch++) image[ch * H * W + pix] = FMA(T, bg[ch], C[ch]);
            alpha_img[pix] = Wt;
            depth_img[pix] = D;
        }
}

/* A.5 backward compositing, one view.  Outputs must be zero-initialised by the caller (accumulated). */
void orc_composite_bwd(int W, int H, const uint32_t *ranges, const uint32_t *vals, const real *xy,
                       const real *conic_opacity, const real *colors, const real *depth, const real *bg,
                       const real *alpha_img, const uint32_t *n_contrib, const real *dL_dimage,
                       const real *dL_dalpha_img, const real *dL_ddepth_img, real *dL_dmean2D, real *dL_dconic,
                       real *dL_dopacity, real *dL_dcolor, real *dL_ddepth)
{
    const int gx = (W + TILE - 1) / TILE;
    const real ddelx_dx = R(0.5f) * (real)W, ddely_dy = R(0.5f) * (real)H;
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            int tile = (py / TILE) * gx + px / TILE, pix = py * W + px;
            uint32_t r0 = ranges[2 * tile];
            real pfx = (real)px, pfy = (real)py;
            const real T_final = R(1.0f) - alpha_img[pix];
            real T = T_final;
            uint32_t last_contributor = n_contrib[pix];
            real acc[3] = {0, 0, 0}, accD = 0, accA = 0, last_alpha = 0, last_c[3] = {0, 0, 0}, last_d = 0;
            real dLdC[3] = {dL_dimage[pix], dL_dimage[H * W + pix], dL_dimage[2 * H * W + pix]};
            real dLdD = dL_ddepth_img[pix], dLdA = dL_dalpha_img[pix];
            real bg_dot = 0;
            for (int ch = 0; ch < 3; ch++) bg_dot += bg[ch] * dLdC[ch];
            for (int64_t pos = (int64_t)last_contributor - 1; pos >= 0; pos--) {
                uint32_t id = vals[r0 + pos];
                real dx = xy[2 * id] - pfx, dy = xy[2 * id + 1] - pfy;
                const real *co = conic_opacity + 4 * id;
                real s = FMA(co[0] * dx, dx, (co[2] * dy) * dy);
                real power = FMA(s, R(-0.5f), -((co[1] * dx) * dy));
                if (power > R(0.0f)) continue;
                real G = EXP(power);
                real alpha = FMIN(ALPHA_MAX, co[3] * G);
                if (alpha < ALPHA_MIN) continue;
                T = T / (R(1.0f) - alpha);
                real w = alpha * T;
                real dL_dalpha = 0;
                for (int ch = 0; ch < 3; ch++) {
                    real c = colors[3 * id + ch];
                    acc[ch] = last_alpha * last_c[ch] + (R(1.0f) - last_alpha) * acc[ch];
                    last_c[ch] = c;
                    dL_dalpha += (c - acc[ch]) * dLdC[ch];
                    dL_dcolor[3 * id + ch] += w * dLdC[ch];
                }
                real cd = depth[id];
                accD = last_alpha * last_d + (R(1.0f) - last_alpha) * accD;
                last_d = cd;
                dL_dalpha += (cd - accD) * dLdD;
                dL_ddepth[id] += w * dLdD;
                accA = last_alpha + (R(1.0f) - last_alpha) * accA;
                dL_dalpha += (R(1.0f) - accA) * dLdA;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (R(1.0f) - alpha)) * bg_dot;
                real dL_dG = co[3] * dL_dalpha;
                real gdx = G * dx, gdy = G * dy;
                real dG_ddelx = -gdx * co[0] - gdy * co[1];
                real dG_ddely = -gdy * co[2] - gdx * co[1];
                dL_dmean2D[2 * id] += dL_dG * dG_ddelx * ddelx_dx;
                dL_dmean2D[2 * id + 1] += dL_dG * dG_ddely * ddely_dy;
                dL_dconic[3 * id] += R(-0.5f) * gdx * dx * dL_dG;
                dL_dconic[3 * id + 1] += R(-0.5f) * gdx * dy * dL_dG;
                dL_dconic[3 * id + 2] += R(-0.5f) * gdy * dy * dL_dG;
                dL_dopacity[id] += G * dL_dalpha;
            }
        }
}

/* ------------------------------------------------------------------------------------------------ */
/* A.6 preprocess backward, one view.  dL_dmeans/scales/rots are ACCUMULATED (+=) so the caller can
 * sum the views of a scene; dL_dcov3d (optional, 6 per Gaussian) is overwritten.                     */
void orc_preprocess_bwd(int P, const real *means, const real *scales, const real *rots, real mod, const real *mv,
                        const real *mp, int W, int H, real tanx, real tany, const int32_t *radii,
                        const real *dL_dmean2D, const real *dL_dconic, const real *dL_ddepth, real *dL_dmeans,
                        real *dL_dscales, real *dL_drots, real *dL_dcov3d_out)
{
    const real fx = (real)W / (R(2.0f) * tanx), fy = (real)H / (R(2.0f) * tany);
    for (int idx = 0; idx < P; idx++) {
        if (dL_dcov3d_out) for (int k = 0; k < 6; k++) dL_dcov3d_out[6 * idx + k] = 0;
        if (!(radii[idx] > 0)) continue;
        real x = means[3 * idx], y = means[3 * idx + 1], z = means[3 * idx + 2];
        real pv[3] = {affine_row(mv, 0, x, y, z), affine_row(mv, 1, x, y, z), affine_row(mv, 2, x, y, z)};
        real Mrow[3][3], c6[6];
        cov3d_from_scale_rot(scales + 3 * idx, mod, rots + 4 * idx, c6, Mrow);
        real abc[3], Tm[2][3], t[3], txtz, tytz;
        cov2d_ewa(pv, fx, fy, tanx, tany, c6, mv, abc, Tm, t, &txtz, &tytz);
        real limx = R(1.3f) * tanx, limy = R(1.3f) * tany;
        real xg = (txtz < -limx || txtz > limx) ? R(0.0f) : R(1.0f);
        real yg = (tytz < -limy || tytz > limy) ? R(0.0f) : R(1.0f);
        real a = abc[0], b = abc[1], c = abc[2];
        real gx_ = dL_dconic[3 * idx], gy_ = dL_dconic[3 * idx + 1], gz_ = dL_dconic[3 * idx + 2];
        real denom = a * c - b * b;
        real dL_da = 0, dL_db = 0, dL_dc = 0;
        real denom2inv = R(1.0f) / ((denom * denom) + W_EPS);
        real g6[6] = {0, 0, 0, 0, 0, 0};
        if (denom2inv != R(0.0f)) {
            dL_da = denom2inv * (-c * c * gx_ + R(2.0f) * b * c * gy_ + (denom - a * c) * gz_);
            dL_dc = denom2inv * (-a * a * gz_ + R(2.0f) * a * b * gy_ + (denom - a * c) * gx_);
            dL_db = denom2inv * R(2.0f) * (b * c * gx_ - (denom + R(2.0f) * b * b) * gy_ + a * b * gz_);
            g6[0] = Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc;
            g6[3] = Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc;
            g6[5] = Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc;
            g6[1] = R(2.0f) * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db + R(2.0f) * Tm[1][0] * Tm[1][1] * dL_dc;
            g6[2] = R(2.0f) * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db + R(2.0f) * Tm[1][0] * Tm[1][2] * dL_dc;
            g6[4] = R(2.0f) * Tm[0][2] * Tm[0][1] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db + R(2.0f) * Tm[1][1] * Tm[1][2] * dL_dc;
        }
        if (dL_dcov3d_out) memcpy(dL_dcov3d_out + 6 * idx, g6, sizeof(g6));
        real V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
        real dT0[3], dT1[3];
        for (int k = 0; k < 3; k++) {
            real tv0 = Tm[0][0] * V[k][0] + Tm[0][1] * V[k][1] + Tm[0][2] * V[k][2];
            real tv1 = Tm[1][0] * V[k][0] + Tm[1][1] * V[k][1] + Tm[1][2] * V[k][2];
            dT0[k] = R(2.0f) * tv0 * dL_da + tv1 * dL_db;
            dT1[k] = R(2.0f) * tv1 * dL_dc + tv0 * dL_db;
        }
        /* Wv[k][i] = mv[k + 4 i]: dL_dJ0k' = sum_i Wv[k'][i] * dT0[i] */
        real dJ00 = mv[0] * dT0[0] + mv[4] * dT0[1] + mv[8] * dT0[2];
        real dJ02 = mv[2] * dT0[0] + mv[6] * dT0[1] + mv[10] * dT0[2];
        real dJ11 = mv[1] * dT1[0] + mv[5] * dT1[1] + mv[9] * dT1[2];
        real dJ12 = mv[2] * dT1[0] + mv[6] * dT1[1] + mv[10] * dT1[2];
        real tz = R(1.0f) / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        real dtx = xg * -fx * tz2 * dJ02;
        real dty = yg * -fy * tz2 * dJ12;
        real dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (R(2.0f) * fx * t[0]) * tz3 * dJ02 + (R(2.0f) * fy * t[1]) * tz3 * dJ12;
        real dm[3];
        dm[0] = mv[0] * dtx + mv[1] * dty + mv[2] * dtz;
        dm[1] = mv[4] * dtx + mv[5] * dty + mv[6] * dtz;
        dm[2] = mv[8] * dtx + mv[9] * dty + mv[10] * dtz;
        /* (ii) projection */
        real hw = affine_row(mp, 3, x, y, z);
        real m_w = R(1.0f) / (hw + W_EPS);
        real mul1 = (mp[0] * x + mp[4] * y + mp[8] * z + mp[12]) * m_w * m_w;
        real mul2 = (mp[1] * x + mp[5] * y + mp[9] * z + mp[13]) * m_w * m_w;
        real g2x = dL_dmean2D[2 * idx], g2y = dL_dmean2D[2 * idx + 1];
        dm[0] += (mp[0] * m_w - mp[3] * mul1) * g2x + (mp[1] * m_w - mp[3] * mul2) * g2y;
        dm[1] += (mp[4] * m_w - mp[7] * mul1) * g2x + (mp[5] * m_w - mp[7] * mul2) * g2y;
        dm[2] += (mp[8] * m_w - mp[11] * mul1) * g2x + (mp[9] * m_w - mp[11] * mul2) * g2y;
        /* (iii) depth */
        real mul3 = mv[2] * x + mv[6] * y + mv[10] * z + mv[14];
        real gd = dL_ddepth[idx];
        dm[0] += (mv[2] - mv[3] * mul3) * gd;
        dm[1] += (mv[6] - mv[7] * mul3) * gd;
        dm[2] += (mv[10] - mv[11] * mul3) * gd;
        for (int k = 0; k < 3; k++) dL_dmeans[3 * idx + k] += dm[k];
        /* (v) cov3D -> scale, rotation.  A = Rq*S (rows Mrow), Sigma = A A^T, dL_dA = 2 * Gsym * A        */
        real Gs[3][3] = {{g6[0], R(0.5f) * g6[1], R(0.5f) * g6[2]},
                         {R(0.5f) * g6[1], g6[3], R(0.5f) * g6[4]},
                         {R(0.5f) * g6[2], R(0.5f) * g6[4], g6[5]}};
        real dA[3][3];
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < 3; k++)
                dA[i][k] = R(2.0f) * (Gs[i][0] * Mrow[0][k] + Gs[i][1] * Mrow[1][k] + Gs[i][2] * Mrow[2][k]);
        real s[3] = {mod * scales[3 * idx], mod * scales[3 * idx + 1], mod * scales[3 * idx + 2]};
        real r = rots[4 * idx], qx = rots[4 * idx + 1], qy = rots[4 * idx + 2], qz = rots[4 * idx + 3];
        real Rq[3][3];
        Rq[0][0] = R(1.0f) - R(2.0f) * (qy * qy + qz * qz); Rq[0][1] = R(2.0f) * (qx * qy - r * qz); Rq[0][2] = R(2.0f) * (qx * qz + r * qy);
        Rq[1][0] = R(2.0f) * (qx * qy + r * qz); Rq[1][1] = R(1.0f) - R(2.0f) * (qx * qx + qz * qz); Rq[1][2] = R(2.0f) * (qy * qz - r * qx);
        Rq[2][0] = R(2.0f) * (qx * qz - r * qy); Rq[2][1] = R(2.0f) * (qy * qz + r * qx); Rq[2][2] = R(1.0f) - R(2.0f) * (qx * qx + qy * qy);
        /* dL_dscale[k] = sum_i Rq[i][k] * dA[i][k]   (upstream quirk: no `mod` factor, A.6 (v))           */
        for (int k = 0; k < 3; k++)
            dL_dscales[3 * idx + k] += Rq[0][k] * dA[0][k] + Rq[1][k] * dA[1][k] + Rq[2][k] * dA[2][k];
        /* dL_dRq[i][k] = dA[i][k] * s[k] ; upstream's dL_dMt[c][r] (glm col c,row r) == dL_dRq[r][c]      */
        real dR[3][3];
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < 3; k++) dR[i][k] = dA[i][k] * s[k];
#define DMT(c, rr) dR[rr][c]
        real dq0 = R(2.0f) * qz * (DMT(0, 1) - DMT(1, 0)) + R(2.0f) * qy * (DMT(2, 0) - DMT(0, 2)) + R(2.0f) * qx * (DMT(1, 2) - DMT(2, 1));
        real dq1 = R(2.0f) * qy * (DMT(1, 0) + DMT(0, 1)) + R(2.0f) * qz * (DMT(2, 0) + DMT(0, 2)) + R(2.0f) * r * (DMT(1, 2) - DMT(2, 1)) - R(4.0f) * qx * (DMT(2, 2) + DMT(1, 1));
        real dq2 = R(2.0f) * qx * (DMT(1, 0) + DMT(0, 1)) + R(2.0f) * r * (DMT(2, 0) - DMT(0, 2)) + R(2.0f) * qz * (DMT(1, 2) + DMT(2, 1)) - R(4.0f) * qy * (DMT(2, 2) + DMT(0, 0));
        real dq3 = R(2.0f) * r * (DMT(0, 1) - DMT(1, 0)) + R(2.0f) * qx * (DMT(2, 0) + DMT(0, 2)) + R(2.0f) * qy * (DMT(1, 2) + DMT(2, 1)) - R(4.0f) * qz * (DMT(1, 1) + DMT(0, 0));
#undef DMT
        dL_drots[4 * idx + 0] += dq0;
        dL_drots[4 * idx + 1] += dq1;
        dL_drots[4 * idx + 2] += dq2;
        dL_drots[4 * idx + 3] += dq3;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Whole step, the way /root/reference/core/gs.py:42-93 drives the rasterizer: for each scene b and
 * view v render, then (optionally) backward given upstream grads; per-Gaussian grads are summed over
 * the views of a scene into dgauss[B,N,14].  OpenMP over views (each view independent).
 * gaussians [B,N,14]: 0:3 pos, 3 opacity, 4:7 scale, 7:11 rot(wxyz), 11:14 rgb  (core/gs.py:45-49).
 * images: [B*V,3,H,W] (NOT clamped; caller clamps as core/gs.py:87), alphas/depths [B*V,1,H,W].
 * dimage/dalpha/ddepth may be NULL -> forward only.  Returns total instances rendered.              */
int64_t orc_render_step(int B, int N, int V, const real *gaussians, const real *view_mats, const real *proj_mats,
                        int W, int H, real tanx, real tany, real mod, const real *bg, real *images, real *alphas,
                        real *depths, int32_t *radii_out, const real *dimage, const real *dalpha,
                        const real *ddepth, real *dgauss)
{
    int64_t total = 0;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    if (dgauss) memset(dgauss, 0, sizeof(real) * (size_t)B * N * 14);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int bv = 0; bv < B * V; bv++) {
        int b = bv / V;
        const real *g = gaussians + (size_t)b * N * 14;
        real *means = malloc(sizeof(real) * 3 * N), *scales = malloc(sizeof(real) * 3 * N);
        real *rots = malloc(sizeof(real) * 4 * N), *opac = malloc(sizeof(real) * N), *cols = malloc(sizeof(real) * 3 * N);
        for (int i = 0; i < N; i++) {
            const real *r = g + (size_t)i * 14;
            memcpy(means + 3 * i, r, 3 * sizeof(real));
            opac[i] = r[3];
            memcpy(scales + 3 * i, r + 4, 3 * sizeof(real));
            memcpy(rots + 4 * i, r + 7, 4 * sizeof(real));
            memcpy(cols + 3 * i, r + 11, 3 * sizeof(real));
        }
        real *depth = malloc(sizeof(real) * N), *xy = malloc(sizeof(real) * 2 * N), *co = malloc(sizeof(real) * 4 * N);
        int32_t *radii = malloc(sizeof(int32_t) * N);
        uint32_t *tiles = malloc(sizeof(uint32_t) * N);
        const real *mv = view_mats + 16 * (size_t)bv, *mp = proj_mats + 16 * (size_t)bv;
        orc_preprocess(N, means, scales, rots, opac, mod, mv, mp, W, H, tanx, tany, depth, radii, xy, co, tiles, NULL, NULL);
        if (radii_out) memcpy(radii_out + (size_t)bv * N, radii, sizeof(int32_t) * N);
        int64_t L = orc_count_instances(N, tiles);
        total += L;
        uint64_t *keys = malloc(sizeof(uint64_t) * (size_t)(L > 0 ? L : 1));
        uint32_t *vals = malloc(sizeof(uint32_t) * (size_t)(L > 0 ? L : 1));
        uint32_t *ranges = malloc(sizeof(uint32_t) * 2 * gx * gy);
        orc_bin(N, radii, xy, depth, W, H, L, keys, vals, ranges, NULL, NULL);
        uint32_t *ncontrib = malloc(sizeof(uint32_t) * HW);
        real *img = images + (size_t)bv * 3 * HW, *al = alphas + (size_t)bv * HW, *dp = depths + (size_t)bv * HW;
        orc_composite_fwd(W, H, ranges, vals, xy, co, cols, depth, bg, img, al, dp, ncontrib);
        if (dimage && dgauss) {
            real *g2 = calloc(2 * N, sizeof(real)), *gc = calloc(3 * N, sizeof(real)), *go = calloc(N, sizeof(real));
            real *gcol = calloc(3 * N, sizeof(real)), *gd = calloc(N, sizeof(real));
            real *gm = calloc(3 * N, sizeof(real)), *gs = calloc(3 * N, sizeof(real)), *gr = calloc(4 * N, sizeof(real));
            /* upstream grads of the UNclamped image: caller passes them already masked by the clamp */
            orc_composite_bwd(W, H, ranges, vals, xy, co, cols, depth, bg, al, ncontrib, dimage + (size_t)bv * 3 * HW,
                              dalpha + (size_t)bv * HW, ddepth + (size_t)bv * HW, g2, gc, go, gcol, gd);
            orc_preprocess_bwd(N, means, scales, rots, mod, mv, mp, W, H, tanx, tany, radii, g2, gc, gd, gm, gs, gr, NULL);
            real *dg = dgauss + (size_t)b * N * 14;
#pragma omp critical
            for (int i = 0; i < N; i++) {
                real *d = dg + (size_t)i * 14;
                for (int k = 0; k < 3; k++) d[k] += gm[3 * i + k];
                d[3] += go[i];
                for (int k = 0; k < 3; k++) d[4 + k] += gs[3 * i + k];
                for (int k = 0; k < 4; k++) d[7 + k] += gr[4 * i + k];
                for (int k = 0; k < 3; k++) d[11 + k] += gcol[3 * i + k];
            }
            free(g2); free(gc); free(go); free(gcol); free(gd); free(gm); free(gs); free(gr);
        }
        free(means); free(scales); free(rots); free(opac); free(cols); free(depth); free(xy); free(co);
        free(radii); free(tiles); free(keys); free(vals); free(ranges); free(ncontrib);
    }
    return total;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int orc_real_bytes(void) { return (int)sizeof(real); }
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------ */
/* Spherical-harmonics colour (the `shs` input of GaussianRasterizer; upstream computeColorFromSH, not on LGM's
 * path).  shs [P, M, 3]; active bands 0..deg; colour = max(0.5 + sum_k basis_k(dir) sh_k, 0), dir = normalize(mean -
 * campos).  The backward is written from the basis table below (independent of the CUDA kernel's closed forms):
 * dL/dsh_k = basis_k * g, dL/ddir = sum_k grad(basis_k) sh_k g, then through the normalisation.                 */
static void sh_basis(int deg, real x, real y, real z, real *b, real (*db)[3])
{
    const real C0 = R(0.28209479177387814), C1 = R(0.4886025119029199);
    const real C2[5] = {R(1.0925484305920792), R(-1.0925484305920792), R(0.31539156525252005), R(-1.0925484305920792), R(0.5462742152960396)};
    const real C3[7] = {R(-0.5900435899266435), R(2.890611442640554), R(-0.4570457994644658), R(0.3731763325901154), R(-0.4570457994644658), R(1.445305721320277), R(-0.5900435899266435)};
    for (int k = 0; k < 16; k++) { b[k] = 0; db[k][0] = db[k][1] = db[k][2] = 0; }
    b[0] = C0;
    if (deg > 0) {
        b[1] = -C1 * y; db[1][1] = -C1;
        b[2] = C1 * z;  db[2][2] = C1;
        b[3] = -C1 * x; db[3][0] = -C1;
    }
    if (deg > 1) {
        b[4] = C2[0] * x * y; db[4][0] = C2[0] * y; db[4][1] = C2[0] * x;
        b[5] = C2[1] * y * z; db[5][1] = C2[1] * z; db[5][2] = C2[1] * y;
        b[6] = C2[2] * (2 * z * z - x * x - y * y); db[6][0] = -2 * C2[2] * x; db[6][1] = -2 * C2[2] * y; db[6][2] = 4 * C2[2] * z;
        b[7] = C2[3] * x * z; db[7][0] = C2[3] * z; db[7][2] = C2[3] * x;
        b[8] = C2[4] * (x * x - y * y); db[8][0] = 2 * C2[4] * x; db[8][1] = -2 * C2[4] * y;
    }
    if (deg > 2) {
        b[9] = C3[0] * y * (3 * x * x - y * y); db[9][0] = 6 * C3[0] * x * y; db[9][1] = C3[0] * (3 * x * x - 3 * y * y);
        b[10] = C3[1] * x * y * z; db[10][0] = C3[1] * y * z; db[10][1] = C3[1] * x * z; db[10][2] = C3[1] * x * y;
        b[11] = C3[2] * y * (4 * z * z - x * x - y * y); db[11][0] = -2 * C3[2] * x * y; db[11][1] = C3[2] * (4 * z * z - x * x - 3 * y * y); db[11][2] = 8 * C3[2] * y * z;
        b[12] = C3[3] * z * (2 * z * z - 3 * x * x - 3 * y * y); db[12][0] = -6 * C3[3] * x * z; db[12][1] = -6 * C3[3] * y * z; db[12][2] = C3[3] * (6 * z * z - 3 * x * x - 3 * y * y);
        b[13] = C3[4] * x * (4 * z * z - x * x - y * y); db[13][0] = C3[4] * (4 * z * z - 3 * x * x - y * y); db[13][1] = -2 * C3[4] * x * y; db[13][2] = 8 * C3[4] * x * z;
        b[14] = C3[5] * z * (x * x - y * y); db[14][0] = 2 * C3[5] * x * z; db[14][1] = -2 * C3[5] * y * z; db[14][2] = C3[5] * (x * x - y * y);
        b[15] = C3[6] * x * (x * x - 3 * y * y); db[15][0] = C3[6] * (3 * x * x - 3 * y * y); db[15][1] = -6 * C3[6] * x * y;
    }
}

void orc_sh_forward(int P, int deg, int M, const real *means, const real *campos, const real *shs, real *colors,
                    uint8_t *clamped)
{
    for (int i = 0; i < P; i++) {
        real d[3] = {means[3 * i] - campos[0], means[3 * i + 1] - campos[1], means[3 * i + 2] - campos[2]};
        real inv = R(1.0) / SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        real b[16], db[16][3];
        sh_basis(deg, d[0] * inv, d[1] * inv, d[2] * inv, b, db);
        int nk = (deg + 1) * (deg + 1);
        for (int c = 0; c < 3; c++) {
            real r = R(0.5);
            for (int k = 0; k < nk; k++) r += b[k] * shs[((size_t)i * M + k) * 3 + c];
            clamped[3 * i + c] = r < 0;
            colors[3 * i + c] = r < 0 ? 0 : r;
        }
    }
}

void orc_sh_backward(int P, int deg, int M, const real *means, const real *campos, const real *shs,
                     const uint8_t *clamped, const real *dL_dcolor, real *dL_dshs, real *dL_dmeans)
{
    for (int i = 0; i < P; i++) {
        real d[3] = {means[3 * i] - campos[0], means[3 * i + 1] - campos[1], means[3 * i + 2] - campos[2]};
        real inv = R(1.0) / SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        real u[3] = {d[0] * inv, d[1] * inv, d[2] * inv};
        real b[16], db[16][3];
        sh_basis(deg, u[0], u[1], u[2], b, db);
        int nk = (deg + 1) * (deg + 1);
        real gdir[3] = {0, 0, 0};
        for (int k = 0; k < M; k++)
            for (int c = 0; c < 3; c++) dL_dshs[((size_t)i * M + k) * 3 + c] = 0;
        for (int c = 0; c < 3; c++) {
            real g = clamped[3 * i + c] ? 0 : dL_dcolor[3 * i + c];
            for (int k = 0; k < nk; k++) {
                dL_dshs[((size_t)i * M + k) * 3 + c] = b[k] * g;
                for (int a = 0; a < 3; a++) gdir[a] += db[k][a] * shs[((size_t)i * M + k) * 3 + c] * g;
            }
        }
        real dot = u[0] * gdir[0] + u[1] * gdir[1] + u[2] * gdir[2];
        for (int a = 0; a < 3; a++) dL_dmeans[3 * i + a] = (gdir[a] - u[a] * dot) * inv;
    }
}
