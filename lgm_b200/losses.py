"""The supervision immediately after the render path: /root/reference/core/models.py:153

    loss_mse = F.mse_loss(pred_images, gt_images) + F.mse_loss(pred_alphas, gt_masks)

as one launch for the loss and one for the gradients w.r.t. the rendered image / alpha, written straight into the
tensors the compositing backward reads (lgm_mse_loss_grad, lgm_b200/csrc/loss.cu) — instead of autograd's chain of
elementwise and reduction launches over the images.  No CPU path.
"""
import torch

from . import _lib, ops


class _MSEImageAlpha(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_images, pred_alphas, gt_images, gt_masks, w_image, w_alpha):
        for name, t in (("pred_images", pred_images), ("pred_alphas", pred_alphas), ("gt_images", gt_images), ("gt_masks", gt_masks)):
            if not t.is_cuda:
                raise _lib.LgmError(f"{name} must be a CUDA tensor (lgm_b200 has no CPU path)")
        if pred_images.shape != gt_images.shape or pred_alphas.shape != gt_masks.shape:
            raise _lib.LgmError("prediction and ground-truth shapes differ")
        x, a = pred_images.contiguous().float(), pred_alphas.contiguous().float()
        # 8-bit ground truth (both tensors uint8, value / 255 — what an image file holds) stays 8-bit: a quarter of the
        # bytes to copy from the host; anything else is read as float32
        ctx.u8 = gt_images.dtype == torch.uint8 and gt_masks.dtype == torch.uint8
        if ctx.u8:
            gx, ga = gt_images.contiguous(), gt_masks.contiguous()
        else:
            gx, ga = gt_images.contiguous().float(), gt_masks.contiguous().float()
        loss = torch.empty(1, dtype=torch.float64, device=x.device)
        ctx.wi = 1.0 / max(x.numel(), 1) if w_image is None else float(w_image)
        ctx.wa = 1.0 / max(a.numel(), 1) if w_alpha is None else float(w_alpha)
        fn = _lib.lib().lgm_mse_loss_grad_u8 if ctx.u8 else _lib.lib().lgm_mse_loss_grad
        _lib.check(fn(ops._stream(), _lib.ptr(x), _lib.ptr(gx), None, x.numel(), ctx.wi, _lib.ptr(a), _lib.ptr(ga), None,
                      a.numel(), ctx.wa, _lib.ptr(loss), None), "lgm_mse_loss_grad")
        ops.launch_counter["kernels"] += 1
        ctx.save_for_backward(x, a, gx, ga)
        return loss[0].float()

    @staticmethod
    def backward(ctx, grad_loss):
        x, a, gx, ga = ctx.saved_tensors
        if grad_loss is None:
            return None, None, None, None, None, None
        d_x, d_a = torch.empty_like(x), torch.empty_like(a)
        scale = grad_loss.reshape(1).float().contiguous()  # stays on the device: no host sync to look at its value
        fn = _lib.lib().lgm_mse_loss_grad_u8 if ctx.u8 else _lib.lib().lgm_mse_loss_grad
        _lib.check(fn(ops._stream(), _lib.ptr(x), _lib.ptr(gx), _lib.ptr(d_x), x.numel(), ctx.wi, _lib.ptr(a), _lib.ptr(ga),
                      _lib.ptr(d_a), a.numel(), ctx.wa, None, _lib.ptr(scale)), "lgm_mse_loss_grad")
        ops.launch_counter["kernels"] += 1
        return d_x, d_a, None, None, None, None


def mse_image_alpha_loss(pred_images, pred_alphas, gt_images, gt_masks, w_image=None, w_alpha=None):
    """mse_loss(pred_images, gt_images) + mse_loss(pred_alphas, gt_masks) (mean reduction by default; w_image / w_alpha
    override the per-element weights, e.g. 1 / global element count when the views are sharded over ranks).
    gt_images and gt_masks may both be uint8 (0..255 = 0..1): they are then read as they are, without a float copy."""
    return _MSEImageAlpha.apply(pred_images, pred_alphas, gt_images, gt_masks, w_image, w_alpha)


class _ResizeBilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, mul, add):
        if not x.is_cuda:
            raise _lib.LgmError("images must be a CUDA tensor (lgm_b200 has no CPU path)")
        if x.dim() < 2:
            raise _lib.LgmError("images must end in (H, W)")
        xc = x.contiguous().float()
        h_in, w_in = xc.shape[-2], xc.shape[-1]
        h_out, w_out = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
        planes = xc.numel() // max(h_in * w_in, 1)
        y = torch.empty(*xc.shape[:-2], h_out, w_out, dtype=torch.float32, device=xc.device)
        L = _lib.lib()
        for p0 in range(0, planes, 65535):  # grid.y limit
            n = min(65535, planes - p0)
            _lib.check(L.lgm_resize_bilinear_forward(ops._stream(), xc.data_ptr() + 4 * p0 * h_in * w_in,
                                                     y.data_ptr() + 4 * p0 * h_out * w_out, n, h_in, w_in, h_out, w_out,
                                                     float(mul), float(add)), "lgm_resize_bilinear_forward")
            ops.launch_counter["kernels"] += 1
        ctx.geom = (planes, h_in, w_in, h_out, w_out, float(mul), tuple(xc.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        planes, h_in, w_in, h_out, w_out, mul, shape = ctx.geom
        dyc = dy.contiguous().float()
        dx = torch.empty(shape, dtype=torch.float32, device=dyc.device)
        L = _lib.lib()
        for p0 in range(0, planes, 65535):
            n = min(65535, planes - p0)
            _lib.check(L.lgm_resize_bilinear_backward(ops._stream(), dyc.data_ptr() + 4 * p0 * h_out * w_out,
                                                      dx.data_ptr() + 4 * p0 * h_in * w_in, n, h_in, w_in, h_out, w_out, mul),
                       "lgm_resize_bilinear_backward")
            ops.launch_counter["kernels"] += 1
        return dx, None, None, None


def lpips_input(images, size=256):
    """F.interpolate(images * 2 - 1, (size, size), mode='bilinear', align_corners=False) for images [..., H, W] — the
    LPIPS input preparation of /root/reference/core/models.py:155-163 (the LPIPS network itself stays the caller's)."""
    return _ResizeBilinear.apply(images, size, 2.0, -1.0)
