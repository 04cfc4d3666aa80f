// TEST INFRASTRUCTURE ONLY: runs the product's __host__ __device__ arithmetic (lgm_b200/csrc/splat_math.cuh) on
// the CPU so that its bit-level agreement with the independent C oracle can be checked without a GPU.
// Built by tests/emul/build.py with g++ -ffp-contract=off -mfma.  Never loaded by lgm_b200/.
#include "../../lgm_b200/csrc/splat_math.cuh"
#include <string.h>
extern "C" {
void emul_preprocess(int P, const float* means, const float* scales, const float* rots, const float* opac, float mod,
                     const float* mv, const float* mp, int W, int H, float tanx, float tany, float* depth,
                     int32_t* radii, float* xy, float* conic_opacity, uint32_t* tiles, int32_t* rects)
{
    const float fx = (float)W / (2.0f * tanx), fy = (float)H / (2.0f * tany);
    const int gx = (W + 15) / 16, gy = (H + 15) / 16;
    for (int i = 0; i < P; i++) {
        lgm::Geom g = lgm::preprocess_point(means + 3 * i, scales + 3 * i, rots + 4 * i, mod, mv, mp, W, H, tanx, tany,
                                            fx, fy, gx, gy);
        depth[i] = g.depth; radii[i] = g.radius; xy[2 * i] = g.px; xy[2 * i + 1] = g.py;
        conic_opacity[4 * i] = g.cx; conic_opacity[4 * i + 1] = g.cy; conic_opacity[4 * i + 2] = g.cz;
        conic_opacity[4 * i + 3] = g.radius > 0 ? opac[i] : 0.f;
        tiles[i] = g.tiles;
        rects[4 * i] = g.rx0; rects[4 * i + 1] = g.ry0; rects[4 * i + 2] = g.rx1; rects[4 * i + 3] = g.ry1;
    }
}
void emul_pair_power(int n, const float* con, const float* d, float* out)
{
    for (int i = 0; i < n; i++) out[i] = lgm::pair_power(con[3 * i], con[3 * i + 1], con[3 * i + 2], d[2 * i], d[2 * i + 1]);
}
void emul_preprocess_bwd(int P, const float* means, const float* scales, const float* rots, float mod, const float* mv,
                         const float* mp, int W, int H, float tanx, float tany, const int32_t* radii, const float* g2,
                         const float* gc, const float* gd, float* dmeans, float* dscales, float* drots)
{
    const float fx = (float)W / (2.0f * tanx), fy = (float)H / (2.0f * tany);
    for (int i = 0; i < P; i++) {
        if (!(radii[i] > 0)) continue;
        lgm::preprocess_point_bwd(means + 3 * i, scales + 3 * i, rots + 4 * i, mod, mv, mp, tanx, tany, fx, fy, g2[2 * i],
                                  g2[2 * i + 1], gc[3 * i], gc[3 * i + 1], gc[3 * i + 2], gd[i], dmeans + 3 * i,
                                  dscales + 3 * i, drots + 4 * i);
    }
}
// K7's per-Gaussian arithmetic as the kernel runs it: the gradient rows arrive in MOMENT form (Sx, Sy, Sxx, Sxy, Syy),
// the per-view part accumulates dL/dpos and dL/dcov3D over n_views views (here: the same view n_views times), the map to
// scale / rotation runs once.  moments [P,5], opac [P].
void emul_preprocess_bwd_moments(int P, int n_views, const float* means, const float* scales, const float* rots,
                                 const float* opac, float mod, const float* mv, const float* mp, int W, int H, float tanx,
                                 float tany, const int32_t* radii, const float* moments, const float* gd, float* dmeans,
                                 float* dscales, float* drots)
{
    const float fx = (float)W / (2.0f * tanx), fy = (float)H / (2.0f * tany);
    for (int i = 0; i < P; i++) {
        if (!(radii[i] > 0)) continue;
        float cov6[6], M[9], g6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        lgm::cov3d_from_scale_rot(scales[3 * i], scales[3 * i + 1], scales[3 * i + 2], mod, rots[4 * i], rots[4 * i + 1],
                                  rots[4 * i + 2], rots[4 * i + 3], cov6, M);
        const float* m = moments + 5 * i;
        for (int v = 0; v < n_views; v++)
            lgm::preprocess_point_bwd_view(means + 3 * i, cov6, mv, mp, tanx, tany, fx, fy, m[0], m[1], m[2], m[3], m[4], gd[i],
                                           dmeans + 3 * i, g6, true, (float)W, (float)H, opac[i]);
        lgm::preprocess_point_bwd_finish(scales + 3 * i, rots + 4 * i, mod, g6, dscales + 3 * i, drots + 4 * i);
    }
}
}
