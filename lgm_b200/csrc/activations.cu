// activations.cu — the step immediately before the path: /root/reference/core/models.py:40-44,107-115
// turns the network's raw splatter image x [B,N,14] into Gaussians with five elementwise activations, five slices and
// a cat (and as many autograd nodes on the way back).  Here: one launch forward, one backward, thread per Gaussian.
//   pos      = clamp(x[0:3], -1, 1)            gradient passes where -1 <= x <= 1 (torch.clamp)
//   opacity  = sigmoid(x[3])
//   scale    = 0.1 softplus(x[4:7])            F.softplus: beta 1, threshold 20 (identity above)
//   rotation = x[7:11] / max(|x[7:11]|, 1e-12) F.normalize
//   rgb      = 0.5 tanh(x[11:14]) + 0.5
#include "common.cuh"

namespace lgm {
namespace {

__device__ __forceinline__ void load_row14(const float* __restrict__ p, float (&r)[14])
{
    const float2* q = reinterpret_cast<const float2*>(p);  // rows are 56 B: 8-byte aligned
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const float2 v = q[k];
        r[2 * k] = v.x;
        r[2 * k + 1] = v.y;
    }
}
__device__ __forceinline__ void store_row14(float* __restrict__ p, const float (&r)[14])
{
    float2* q = reinterpret_cast<float2*>(p);
#pragma unroll
    for (int k = 0; k < 7; k++) q[k] = make_float2(r[2 * k], r[2 * k + 1]);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kBlock)
activate_fwd_kernel(size_t n_rows, const float* __restrict__ x, float* __restrict__ g)
{
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n_rows) return;
    float r[14], o[14];
    load_row14(x + i * 14, r);
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = fminf(fmaxf(r[k], -1.0f), 1.0f);
    o[3] = sigmoidf_(r[3]);
#pragma unroll
    for (int k = 4; k < 7; k++) o[k] = 0.1f * (r[k] > 20.0f ? r[k] : log1pf(expf(r[k])));
    const float nrm = sqrtf(r[7] * r[7] + r[8] * r[8] + r[9] * r[9] + r[10] * r[10]);
    const float inv = 1.0f / fmaxf(nrm, 1e-12f);
#pragma unroll
    for (int k = 7; k < 11; k++) o[k] = r[k] * inv;
#pragma unroll
    for (int k = 11; k < 14; k++) o[k] = 0.5f * tanhf(r[k]) + 0.5f;
    store_row14(g + i * 14, o);
}

__global__ void __launch_bounds__(kBlock)
activate_bwd_kernel(size_t n_rows, const float* __restrict__ x, const float* __restrict__ dg, float* __restrict__ dx)
{
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n_rows) return;
    float r[14], d[14], o[14];
    load_row14(x + i * 14, r);
    load_row14(dg + i * 14, d);
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = (r[k] >= -1.0f && r[k] <= 1.0f) ? d[k] : 0.0f;
    const float s = sigmoidf_(r[3]);
    o[3] = d[3] * s * (1.0f - s);
#pragma unroll
    for (int k = 4; k < 7; k++) o[k] = 0.1f * d[k] * (r[k] > 20.0f ? 1.0f : sigmoidf_(r[k]));
    const float nrm = sqrtf(r[7] * r[7] + r[8] * r[8] + r[9] * r[9] + r[10] * r[10]);
    if (nrm > 1e-12f) {
        const float inv = 1.0f / nrm;
        const float n0 = r[7] * inv, n1 = r[8] * inv, n2 = r[9] * inv, n3 = r[10] * inv;
        const float dot = n0 * d[7] + n1 * d[8] + n2 * d[9] + n3 * d[10];
        o[7] = (d[7] - n0 * dot) * inv;
        o[8] = (d[8] - n1 * dot) * inv;
        o[9] = (d[9] - n2 * dot) * inv;
        o[10] = (d[10] - n3 * dot) * inv;
    } else {  // below the eps clamp the map is x / 1e-12
#pragma unroll
        for (int k = 7; k < 11; k++) o[k] = d[k] * 1e12f;
    }
#pragma unroll
    for (int k = 11; k < 14; k++) {
        const float t = tanhf(r[k]);
        o[k] = 0.5f * d[k] * (1.0f - t * t);
    }
    store_row14(dx + i * 14, o);
}

}  // namespace

cudaError_t launch_activate_fwd(cudaStream_t stream, size_t n_rows, const float* x, float* g)
{
    if (n_rows == 0) return cudaSuccess;
    activate_fwd_kernel<<<(unsigned)((n_rows + kBlock - 1) / kBlock), kBlock, 0, stream>>>(n_rows, x, g);
    return cudaGetLastError();
}

cudaError_t launch_activate_bwd(cudaStream_t stream, size_t n_rows, const float* x, const float* dg, float* dx)
{
    if (n_rows == 0) return cudaSuccess;
    activate_bwd_kernel<<<(unsigned)((n_rows + kBlock - 1) / kBlock), kBlock, 0, stream>>>(n_rows, x, dg, dx);
    return cudaGetLastError();
}

}  // namespace lgm
