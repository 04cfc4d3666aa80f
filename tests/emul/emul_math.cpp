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
}
