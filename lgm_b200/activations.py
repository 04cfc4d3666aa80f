"""The step immediately before the render path: /root/reference/core/models.py:40-44,107-115

    pos = x[..., 0:3].clamp(-1, 1); opacity = sigmoid(x[..., 3:4]); scale = 0.1 * softplus(x[..., 4:7])
    rotation = F.normalize(x[..., 7:11]); rgbs = 0.5 * tanh(x[..., 11:]) + 0.5
    gaussians = torch.cat([pos, opacity, scale, rotation, rgbs], dim=-1)

as one launch forward and one backward (lgm_b200/csrc/activations.cu) instead of five activations, five slices and a
cat with their autograd nodes.  No CPU path.
"""
import torch

from . import _lib, ops


class _ActivateGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if not x.is_cuda:
            raise _lib.LgmError("x must be a CUDA tensor (lgm_b200 has no CPU path)")
        if x.shape[-1] != 14:
            raise _lib.LgmError(f"x must have 14 channels last, got {tuple(x.shape)}")
        xc = x.contiguous().float()
        g = torch.empty_like(xc)
        _lib.check(_lib.lib().lgm_activate_forward(ops._stream(), xc.numel() // 14, _lib.ptr(xc), _lib.ptr(g)),
                   "lgm_activate_forward")
        ops.launch_counter["kernels"] += 1 if xc.numel() else 0
        ctx.save_for_backward(xc)
        return g

    @staticmethod
    def backward(ctx, dg):
        (xc,) = ctx.saved_tensors
        dgc = dg.contiguous().float()
        dx = torch.empty_like(xc)
        _lib.check(_lib.lib().lgm_activate_backward(ops._stream(), xc.numel() // 14, _lib.ptr(xc), _lib.ptr(dgc), _lib.ptr(dx)),
                   "lgm_activate_backward")
        ops.launch_counter["kernels"] += 1 if xc.numel() else 0
        return dx


def activate_gaussians(x):
    """x [..., 14] raw network output (channels last, as after `x.permute(0, 1, 3, 4, 2).reshape(B, -1, 14)`) ->
    Gaussians [..., 14] = (pos 3, opacity 1, scale 3, rotation 4, rgb 3), the input of GaussianRenderer.render."""
    return _ActivateGaussians.apply(x)
