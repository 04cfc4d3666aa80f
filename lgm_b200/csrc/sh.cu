// sh.cu — view-dependent colour from spherical harmonics (degrees 0..3) and its backward: the `shs` input of
// GaussianRasterizer (upstream computeColorFromSH fwd/bwd in forward.cu / backward.cu, SURVEY.md §8b level 1).
// Not on LGM's own path (core/gs.py:79-80 passes colors_precomp); kept out of the batched hot kernels: the colours
// are produced per view by these kernels and then enter the renderer exactly like colors_precomp.
//
// colour = clamp_min(0.5 + sum_l,m c_lm * Y_lm(dir) * sh_lm, 0),  dir = normalize(mean - campos); the clamp mask is
// kept for the backward; dL/dmean flows through the normalisation of dir.
#include "common.cuh"

namespace lgm {
namespace {

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                                   0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,  -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

__global__ void __launch_bounds__(kBlock)
sh_forward_kernel(int P, int deg, int max_coeffs, const float* __restrict__ means, const float* __restrict__ campos,
                  const float* __restrict__ shs, float* __restrict__ colors, uint8_t* __restrict__ clamped)
{
    const int idx = blockIdx.x * kBlock + threadIdx.x;
    if (idx >= P) return;
    const float3 pos = make_float3(means[3 * idx], means[3 * idx + 1], means[3 * idx + 2]);
    float3 dir = make_float3(pos.x - campos[0], pos.y - campos[1], pos.z - campos[2]);
    const float inv = 1.0f / sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x *= inv; dir.y *= inv; dir.z *= inv;
    const float* sh = shs + (size_t)idx * max_coeffs * 3;
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
#define SH(i) sh[(i) * 3 + c]
        float r = SH_C0 * SH(0);
        if (deg > 0) {
            const float x = dir.x, y = dir.y, z = dir.z;
            r = r - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) + SH_C2[2] * (2.0f * zz - xx - yy) * SH(6) +
                    SH_C2[3] * xz * SH(7) + SH_C2[4] * (xx - yy) * SH(8);
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (3.0f * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * SH(11) + SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SH(12) +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * SH(13) + SH_C3[5] * z * (xx - yy) * SH(14) +
                        SH_C3[6] * x * (xx - 3.0f * yy) * SH(15);
                }
            }
        }
#undef SH
        res[c] = r + 0.5f;
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        clamped[3 * idx + c] = res[c] < 0.0f;
        colors[3 * idx + c] = fmaxf(res[c], 0.0f);
    }
}

__global__ void __launch_bounds__(kBlock)
sh_backward_kernel(int P, int deg, int max_coeffs, const float* __restrict__ means, const float* __restrict__ campos,
                   const float* __restrict__ shs, const uint8_t* __restrict__ clamped, const float* __restrict__ dL_dcolor,
                   float* __restrict__ dL_dshs, float* __restrict__ dL_dmeans)
{
    const int idx = blockIdx.x * kBlock + threadIdx.x;
    if (idx >= P) return;
    const float3 pos = make_float3(means[3 * idx], means[3 * idx + 1], means[3 * idx + 2]);
    const float3 d0 = make_float3(pos.x - campos[0], pos.y - campos[1], pos.z - campos[2]);
    const float inv = 1.0f / sqrtf(d0.x * d0.x + d0.y * d0.y + d0.z * d0.z);
    const float x = d0.x * inv, y = d0.y * inv, z = d0.z * inv;
    const float* sh = shs + (size_t)idx * max_coeffs * 3;
    float* dsh = dL_dshs + (size_t)idx * max_coeffs * 3;
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;  // dL/d(dir)
    for (int c = 0; c < 3; c++) {
        const float g = clamped[3 * idx + c] ? 0.0f : dL_dcolor[3 * idx + c];
#define SH(i) sh[(i) * 3 + c]
#define DSH(i) dsh[(i) * 3 + c]
        DSH(0) = SH_C0 * g;
        if (deg > 0) {
            DSH(1) = -SH_C1 * y * g;
            DSH(2) = SH_C1 * z * g;
            DSH(3) = -SH_C1 * x * g;
            float dx = -SH_C1 * SH(3), dy = -SH_C1 * SH(1), dz = SH_C1 * SH(2);  // d(colour)/d(dir)
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                DSH(4) = SH_C2[0] * xy * g;
                DSH(5) = SH_C2[1] * yz * g;
                DSH(6) = SH_C2[2] * (2.0f * zz - xx - yy) * g;
                DSH(7) = SH_C2[3] * xz * g;
                DSH(8) = SH_C2[4] * (xx - yy) * g;
                dx += SH_C2[0] * y * SH(4) + SH_C2[2] * 2.0f * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * 2.0f * x * SH(8);
                dy += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * 2.0f * -y * SH(6) + SH_C2[4] * 2.0f * -y * SH(8);
                dz += SH_C2[1] * y * SH(5) + SH_C2[2] * 2.0f * 2.0f * z * SH(6) + SH_C2[3] * x * SH(7);
                if (deg > 2) {
                    DSH(9) = SH_C3[0] * y * (3.0f * xx - yy) * g;
                    DSH(10) = SH_C3[1] * xy * z * g;
                    DSH(11) = SH_C3[2] * y * (4.0f * zz - xx - yy) * g;
                    DSH(12) = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * g;
                    DSH(13) = SH_C3[4] * x * (4.0f * zz - xx - yy) * g;
                    DSH(14) = SH_C3[5] * z * (xx - yy) * g;
                    DSH(15) = SH_C3[6] * x * (xx - 3.0f * yy) * g;
                    dx += SH_C3[0] * SH(9) * 3.0f * 2.0f * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * -2.0f * xy +
                          SH_C3[3] * SH(12) * -3.0f * 2.0f * xz + SH_C3[4] * SH(13) * (-3.0f * xx + 4.0f * zz - yy) +
                          SH_C3[5] * SH(14) * 2.0f * xz + SH_C3[6] * SH(15) * 3.0f * (xx - yy);
                    dy += SH_C3[0] * SH(9) * 3.0f * (xx - yy) + SH_C3[1] * SH(10) * xz + SH_C3[2] * SH(11) * (-3.0f * yy + 4.0f * zz - xx) +
                          SH_C3[3] * SH(12) * -3.0f * 2.0f * yz + SH_C3[4] * SH(13) * -2.0f * xy + SH_C3[5] * SH(14) * -2.0f * yz +
                          SH_C3[6] * SH(15) * -3.0f * 2.0f * xy;
                    dz += SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * 4.0f * 2.0f * yz + SH_C3[3] * SH(12) * 3.0f * (2.0f * zz - xx - yy) +
                          SH_C3[4] * SH(13) * 4.0f * 2.0f * xz + SH_C3[5] * SH(14) * (xx - yy);
                }
            }
            ddx += dx * g; ddy += dy * g; ddz += dz * g;
        }
        for (int i = (deg + 1) * (deg + 1); i < max_coeffs; i++) DSH(i) = 0.0f;  // inactive bands get no gradient
#undef SH
#undef DSH
    }
    // through dir = d0 / |d0|:  dL/dd0 = (I - dir dir^T) dL/ddir / |d0|
    const float dot = x * ddx + y * ddy + z * ddz;
    dL_dmeans[3 * idx] = (ddx - x * dot) * inv;
    dL_dmeans[3 * idx + 1] = (ddy - y * dot) * inv;
    dL_dmeans[3 * idx + 2] = (ddz - z * dot) * inv;
}

}  // namespace

cudaError_t launch_sh_forward(cudaStream_t stream, int P, int deg, int max_coeffs, const float* means, const float* campos,
                              const float* shs, float* colors, uint8_t* clamped)
{
    if (P == 0) return cudaSuccess;
    sh_forward_kernel<<<(P + kBlock - 1) / kBlock, kBlock, 0, stream>>>(P, deg, max_coeffs, means, campos, shs, colors, clamped);
    return cudaGetLastError();
}

cudaError_t launch_sh_backward(cudaStream_t stream, int P, int deg, int max_coeffs, const float* means, const float* campos,
                               const float* shs, const uint8_t* clamped, const float* dL_dcolor, float* dL_dshs,
                               float* dL_dmeans)
{
    if (P == 0) return cudaSuccess;
    sh_backward_kernel<<<(P + kBlock - 1) / kBlock, kBlock, 0, stream>>>(P, deg, max_coeffs, means, campos, shs, clamped, dL_dcolor,
                                                                       dL_dshs, dL_dmeans);
    return cudaGetLastError();
}

}  // namespace lgm
