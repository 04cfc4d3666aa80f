// resize.cu — the LPIPS input preparation of /root/reference/core/models.py:155-163:
//   F.interpolate(images.view(-1, 3, S, S) * 2 - 1, (256, 256), mode='bilinear', align_corners=False)
// as one launch forward (scale, shift and resample fused; the weights of a bilinear tap sum to 1, so
// resize(2x - 1) = 2 resize(x) - 1) and one backward.  PyTorch's sampling rule: src = (dst + 0.5) * in / out - 0.5,
// clamped at 0; taps floor(src) and min(floor(src) + 1, in - 1) with weights (1 - l, l).
#include "common.cuh"

namespace lgm {
namespace {

struct Tap {
    int i0, i1;
    float w0, w1;
};
__device__ __forceinline__ Tap tap_of(int dst, float scale, int n_in)
{
    float src = ((float)dst + 0.5f) * scale - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    Tap t;
    t.i0 = min((int)src, n_in - 1);
    t.i1 = min(t.i0 + 1, n_in - 1);
    t.w1 = src - (float)t.i0;
    t.w0 = 1.0f - t.w1;
    return t;
}

// one thread per output pixel of one plane; planes in blockIdx.y
__global__ void __launch_bounds__(kBlock)
resize_bilinear_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int h_in, int w_in, int h_out, int w_out,
                           float mul, float add)
{
    const int o = blockIdx.x * kBlock + threadIdx.x;
    if (o >= h_out * w_out) return;
    const int oy = o / w_out, ox = o - oy * w_out;
    const Tap ty = tap_of(oy, (float)h_in / (float)h_out, h_in), tx = tap_of(ox, (float)w_in / (float)w_out, w_in);
    const float* p = x + (size_t)blockIdx.y * h_in * w_in;
    const float v = ty.w0 * (tx.w0 * p[ty.i0 * w_in + tx.i0] + tx.w1 * p[ty.i0 * w_in + tx.i1]) +
                    ty.w1 * (tx.w0 * p[ty.i1 * w_in + tx.i0] + tx.w1 * p[ty.i1 * w_in + tx.i1]);
    y[(size_t)blockIdx.y * h_out * w_out + o] = mul * v + add;
}

// dx must be zero on entry: every output pixel adds its four weighted taps (fp32 RED)
__global__ void __launch_bounds__(kBlock)
resize_bilinear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int h_in, int w_in, int h_out, int w_out,
                           float mul)
{
    const int o = blockIdx.x * kBlock + threadIdx.x;
    if (o >= h_out * w_out) return;
    const int oy = o / w_out, ox = o - oy * w_out;
    const Tap ty = tap_of(oy, (float)h_in / (float)h_out, h_in), tx = tap_of(ox, (float)w_in / (float)w_out, w_in);
    float* p = dx + (size_t)blockIdx.y * h_in * w_in;
    const float g = mul * dy[(size_t)blockIdx.y * h_out * w_out + o];
    atomicAdd(p + ty.i0 * w_in + tx.i0, g * ty.w0 * tx.w0);
    atomicAdd(p + ty.i0 * w_in + tx.i1, g * ty.w0 * tx.w1);
    atomicAdd(p + ty.i1 * w_in + tx.i0, g * ty.w1 * tx.w0);
    atomicAdd(p + ty.i1 * w_in + tx.i1, g * ty.w1 * tx.w1);
}

}  // namespace

cudaError_t launch_resize_bilinear_fwd(cudaStream_t stream, const float* x, float* y, int n_planes, int h_in, int w_in, int h_out,
                                       int w_out, float mul, float add)
{
    if (n_planes == 0 || h_out * w_out == 0) return cudaSuccess;
    dim3 grid((h_out * w_out + kBlock - 1) / kBlock, n_planes);
    resize_bilinear_fwd_kernel<<<grid, kBlock, 0, stream>>>(x, y, h_in, w_in, h_out, w_out, mul, add);
    return cudaGetLastError();
}

cudaError_t launch_resize_bilinear_bwd(cudaStream_t stream, const float* dy, float* dx, int n_planes, int h_in, int w_in, int h_out,
                                       int w_out, float mul)
{
    if (n_planes == 0) return cudaSuccess;
    cudaError_t err = cudaMemsetAsync(dx, 0, (size_t)n_planes * h_in * w_in * sizeof(float), stream);
    if (err != cudaSuccess || h_out * w_out == 0) return err;
    dim3 grid((h_out * w_out + kBlock - 1) / kBlock, n_planes);
    resize_bilinear_bwd_kernel<<<grid, kBlock, 0, stream>>>(dy, dx, h_in, w_in, h_out, w_out, mul);
    return cudaGetLastError();
}

}  // namespace lgm
