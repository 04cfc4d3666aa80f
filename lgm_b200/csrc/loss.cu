// loss.cu — the step immediately after the path: the MSE supervision of
// /root/reference/core/models.py:153,  loss = mse(pred_images, gt_images) + mse(pred_alphas, gt_masks),
// evaluated together with its gradient in ONE pass over the rendered views, so that the compositing backward's inputs
// (dL/dimage, dL/dalpha) are produced directly instead of by four elementwise / reduction launches of autograd.
//   d_image = 2 w_i (image - gt_image),  d_alpha = 2 w_a (alpha - gt_alpha),  loss = w_i sum (.)^2 + w_a sum (.)^2
// with w_i = 1 / numel(image), w_a = 1 / numel(alpha) for the reference's mean reduction.  The loss is accumulated in
// double (one atomic per CTA), so its value does not depend on the grid to more than 1 ulp of the float result.
// Either output may be omitted (loss == nullptr / d_* == nullptr): autograd calls it once for the loss in the forward
// and once for the gradients in the backward, where `grad_scale` (a device float, the incoming dL/dloss) multiplies them
// — no host synchronisation to find out whether that factor is 1.
#include "common.cuh"

namespace lgm {
namespace {

constexpr int kLossBlock = 256;

__device__ __forceinline__ float4 sq_grad(const float4 x, const float4 y, float w2, float& acc)
{
    const float4 d = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    return make_float4(w2 * d.x, w2 * d.y, w2 * d.z, w2 * d.w);
}

// ground truth as float, or as 8-bit values scaled by 1/255 (what an image file holds: a quarter of the PCIe bytes)
__device__ __forceinline__ float4 load_gt4(const float* y, size_t i) { return __ldg(reinterpret_cast<const float4*>(y) + i); }
__device__ __forceinline__ float load_gt1(const float* y, size_t i) { return y[i]; }
__device__ __forceinline__ float4 load_gt4(const uint8_t* y, size_t i)
{
    const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(y) + i);
    constexpr float k = 1.0f / 255.0f;
    return make_float4(k * (float)v.x, k * (float)v.y, k * (float)v.z, k * (float)v.w);
}
__device__ __forceinline__ float load_gt1(const uint8_t* y, size_t i) { return (1.0f / 255.0f) * (float)y[i]; }

// part 0: image (n_img floats, weight w_img), part 1: alpha; both arrays are walked as float4 with a scalar tail
template <typename GT>
__global__ void __launch_bounds__(kLossBlock)
mse_loss_grad_kernel(const float* __restrict__ image, const GT* __restrict__ gt_image, float* __restrict__ d_image,
                     size_t n_img, float w_img, const float* __restrict__ alpha, const GT* __restrict__ gt_alpha,
                     float* __restrict__ d_alpha, size_t n_alpha, float w_alpha, double* __restrict__ loss,
                     const float* __restrict__ grad_scale)
{
    __shared__ double s_part[kLossBlock / 32];
    const size_t tid = (size_t)blockIdx.x * kLossBlock + threadIdx.x, stride = (size_t)gridDim.x * kLossBlock;
    const float gs = grad_scale ? __ldg(grad_scale) : 1.0f;
    double total = 0.0;
#pragma unroll
    for (int part = 0; part < 2; part++) {
        const float* x = part ? alpha : image;
        const GT* y = part ? gt_alpha : gt_image;
        float* d = part ? d_alpha : d_image;
        const size_t n = part ? n_alpha : n_img;
        const float w = part ? w_alpha : w_img;
        const float w2 = 2.0f * w * gs;
        float acc = 0.f;
        const size_t n4 = n / 4;
        for (size_t i = tid; i < n4; i += stride) {
            const float4 xv = reinterpret_cast<const float4*>(x)[i], yv = load_gt4(y, i);
            const float4 g = sq_grad(xv, yv, w2, acc);
            if (d) reinterpret_cast<float4*>(d)[i] = g;
        }
        for (size_t i = n4 * 4 + tid; i < n; i += stride) {
            const float df = x[i] - load_gt1(y, i);
            acc += df * df;
            if (d) d[i] = w2 * df;
        }
        total += (double)w * (double)acc;
    }
    if (!loss) return;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kLossBlock / 32; w++) t += s_part[w];
        atomicAdd(loss, t);
    }
}

}  // namespace

// Pointers must be 16-byte aligned (4-byte for 8-bit ground truth); d_image / d_alpha / loss / grad_scale may be null.
template <typename GT>
static cudaError_t launch_mse(cudaStream_t stream, const float* image, const GT* gt_image, float* d_image, size_t n_img,
                              float w_img, const float* alpha, const GT* gt_alpha, float* d_alpha, size_t n_alpha, float w_alpha,
                              double* loss, const float* grad_scale)
{
    if (loss) {
        cudaError_t err = cudaMemsetAsync(loss, 0, sizeof(double), stream);
        if (err != cudaSuccess) return err;
    }
    if (n_img + n_alpha == 0) return cudaSuccess;
    const int n_sm = device_sm_count();
    const size_t want = (n_img / 4 + n_alpha / 4 + kLossBlock - 1) / kLossBlock + 1;
    const unsigned grid = (unsigned)(want < (size_t)n_sm * 16 ? want : (size_t)n_sm * 16);
    mse_loss_grad_kernel<GT><<<grid, kLossBlock, 0, stream>>>(image, gt_image, d_image, n_img, w_img, alpha, gt_alpha, d_alpha,
                                                              n_alpha, w_alpha, loss, grad_scale);
    return cudaGetLastError();
}

cudaError_t launch_mse_loss_grad(cudaStream_t stream, const float* image, const float* gt_image, float* d_image, size_t n_img,
                                 float w_img, const float* alpha, const float* gt_alpha, float* d_alpha, size_t n_alpha,
                                 float w_alpha, double* loss, const float* grad_scale)
{
    return launch_mse<float>(stream, image, gt_image, d_image, n_img, w_img, alpha, gt_alpha, d_alpha, n_alpha, w_alpha, loss, grad_scale);
}

cudaError_t launch_mse_loss_grad_u8(cudaStream_t stream, const float* image, const uint8_t* gt_image, float* d_image, size_t n_img,
                                    float w_img, const float* alpha, const uint8_t* gt_alpha, float* d_alpha, size_t n_alpha,
                                    float w_alpha, double* loss, const float* grad_scale)
{
    return launch_mse<uint8_t>(stream, image, gt_image, d_image, n_img, w_img, alpha, gt_alpha, d_alpha, n_alpha, w_alpha, loss, grad_scale);
}

}  // namespace lgm
