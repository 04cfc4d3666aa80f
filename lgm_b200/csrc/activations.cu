// activations.cu — the step immediately before the path: /root/reference/core/models.py:40-44,107-115
// turns the network's raw splatter image x [B,N,14] into Gaussians with five elementwise activations, five slices and
// a cat (and as many autograd nodes on the way back).  Here: one launch forward, one backward, thread per Gaussian.
//   pos      = clamp(x[0:3], -1, 1)            gradient passes where -1 <= x <= 1 (torch.clamp)
//   opacity  = sigmoid(x[3])
//   scale    = 0.1 softplus(x[4:7])            F.softplus: beta 1, threshold 20 (identity above)
//   rotation = F.normalize(x[..., 7:11])        as the reference CALLS it (models.py:43,112): no dim argument, so the default
//              dim = 1 applies to the [B,N,4] slice — each of the four components is divided by its L2 norm over the N
//              Gaussians of the scene, max(norm, 1e-12).  Not a per-quaternion normalisation; reference checkpoints
//              were trained under it, so it is the default here (rot_axis 0).  rot_axis 1 = dim -1, unit quaternions.
//              The column norms (and, backward, the column sums of g x) come from one small reduction launch.
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

// Column sums over the Gaussians of one scene (blockIdx.y): cols[scene][0..3] += sum x_c^2 and, when dg is given,
// cols[scene][4..7] += sum dg_c x_c, for the four rotation channels c.  Doubles: the float result is then independent
// of the order of the atomics except in the last bit's rare ties.
constexpr int kColRows = 8;  // rows per thread
__global__ void __launch_bounds__(kBlock)
rot_column_sums_kernel(size_t n_per_scene, const float* __restrict__ x, const float* __restrict__ dg, double* __restrict__ cols)
{
    __shared__ double s_part[kBlock / 32][8];
    const size_t base = (size_t)blockIdx.y * n_per_scene;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < kColRows; k++) {
        const size_t i = ((size_t)blockIdx.x * kColRows + k) * kBlock + threadIdx.x;
        if (i >= n_per_scene) break;
        const float* r = x + (base + i) * 14 + 7;
#pragma unroll
        for (int c = 0; c < 4; c++) a[c] = fmaf(r[c], r[c], a[c]);
        if (dg) {
            const float* d = dg + (base + i) * 14 + 7;
#pragma unroll
            for (int c = 0; c < 4; c++) a[4 + c] = fmaf(d[c], r[c], a[4 + c]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        double v = (double)a[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < (dg ? 8 : 4)) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; w++) v += s_part[w][threadIdx.x];
        atomicAdd(cols + (size_t)blockIdx.y * 8 + threadIdx.x, v);
    }
}

__global__ void __launch_bounds__(kBlock)
activate_fwd_kernel(size_t n_rows, size_t n_per_scene, const float* __restrict__ x, float* __restrict__ g,
                    const double* __restrict__ cols)
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
    if (cols) {  // reference axis: per-component norm over the scene's Gaussians
        const double* cs = cols + (i / n_per_scene) * 8;
#pragma unroll
        for (int k = 0; k < 4; k++) o[7 + k] = r[7 + k] / fmaxf(sqrtf((float)cs[k]), 1e-12f);
    } else {
        const float nrm = sqrtf(r[7] * r[7] + r[8] * r[8] + r[9] * r[9] + r[10] * r[10]);
        const float inv = 1.0f / fmaxf(nrm, 1e-12f);
#pragma unroll
        for (int k = 7; k < 11; k++) o[k] = r[k] * inv;
    }
#pragma unroll
    for (int k = 11; k < 14; k++) o[k] = 0.5f * tanhf(r[k]) + 0.5f;
    store_row14(g + i * 14, o);
}

__global__ void __launch_bounds__(kBlock)
activate_bwd_kernel(size_t n_rows, size_t n_per_scene, const float* __restrict__ x, const float* __restrict__ dg,
                    float* __restrict__ dx, const double* __restrict__ cols)
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
    if (cols) {
        // y_i = x_i / n with n = max(|x_c| over the scene, eps):  dx_i = g_i / n - x_i (sum_j g_j x_j) / n^3
        const double* cs = cols + (i / n_per_scene) * 8;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float nrm = sqrtf((float)cs[k]);
            if (nrm > 1e-12f) {
                const float inv = 1.0f / nrm;
                o[7 + k] = (d[7 + k] - r[7 + k] * (float)cs[4 + k] * inv * inv) * inv;
            } else {
                o[7 + k] = d[7 + k] * 1e12f;
            }
        }
    } else if (const float nrm = sqrtf(r[7] * r[7] + r[8] * r[8] + r[9] * r[9] + r[10] * r[10]); nrm > 1e-12f) {
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

// cols: [n_scenes, 8] doubles of scratch (reference axis), or null (per-quaternion normalisation)
cudaError_t launch_activate_fwd(cudaStream_t stream, size_t n_scenes, size_t n_per_scene, const float* x, float* g, double* cols)
{
    const size_t n_rows = n_scenes * n_per_scene;
    if (n_rows == 0) return cudaSuccess;
    if (cols) {
        cudaError_t e = cudaMemsetAsync(cols, 0, n_scenes * 8 * sizeof(double), stream);
        if (e != cudaSuccess) return e;
        dim3 grid((unsigned)((n_per_scene + kBlock * kColRows - 1) / (kBlock * kColRows)), (unsigned)n_scenes);
        rot_column_sums_kernel<<<grid, kBlock, 0, stream>>>(n_per_scene, x, nullptr, cols);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    activate_fwd_kernel<<<(unsigned)((n_rows + kBlock - 1) / kBlock), kBlock, 0, stream>>>(n_rows, n_per_scene, x, g, cols);
    return cudaGetLastError();
}

cudaError_t launch_activate_bwd(cudaStream_t stream, size_t n_scenes, size_t n_per_scene, const float* x, const float* dg,
                                float* dx, double* cols)
{
    const size_t n_rows = n_scenes * n_per_scene;
    if (n_rows == 0) return cudaSuccess;
    if (cols) {
        cudaError_t e = cudaMemsetAsync(cols, 0, n_scenes * 8 * sizeof(double), stream);
        if (e != cudaSuccess) return e;
        dim3 grid((unsigned)((n_per_scene + kBlock * kColRows - 1) / (kBlock * kColRows)), (unsigned)n_scenes);
        rot_column_sums_kernel<<<grid, kBlock, 0, stream>>>(n_per_scene, x, dg, cols);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    activate_bwd_kernel<<<(unsigned)((n_rows + kBlock - 1) / kBlock), kBlock, 0, stream>>>(n_rows, n_per_scene, x, dg, dx, cols);
    return cudaGetLastError();
}

}  // namespace lgm
