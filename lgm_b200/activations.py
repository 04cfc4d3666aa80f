"""The step immediately before the render path: /root/reference/core/models.py:40-44,107-115

    pos = x[..., 0:3].clamp(-1, 1); opacity = sigmoid(x[..., 3:4]); scale = 0.1 * softplus(x[..., 4:7])
    rotation = F.normalize(x[..., 7:11]); rgbs = 0.5 * tanh(x[..., 11:]) + 0.5
    gaussians = torch.cat([pos, opacity, scale, rotation, rgbs], dim=-1)

as one launch forward and one backward (lgm_b200/csrc/activations.cu) instead of five activations, five slices and a
cat with their autograd nodes.  No CPU path.
"""
import torch

from . import _lib, ops


ROT_AXES = {"reference": 0, "quaternion": 1}


class _ActivateGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rot_axis):
        if not x.is_cuda:
            raise _lib.LgmError("x must be a CUDA tensor (lgm_b200 has no CPU path)")
        if x.shape[-1] != 14:
            raise _lib.LgmError(f"x must have 14 channels last, got {tuple(x.shape)}")
        xc = x.contiguous().float()
        # F.normalize's default dim = 1: the norm runs over the SECOND dimension of the slice the reference passes,
        # [B, N, 4] -> per (scene, component) over the N Gaussians.  A 2-D input [N, 14] has dim 1 = the channel axis.
        if rot_axis == 0 and xc.dim() >= 3:
            n_scenes, n_per = xc.shape[0], xc.numel() // 14 // max(xc.shape[0], 1)
            if xc.dim() > 3:
                raise _lib.LgmError("rot_axis='reference' expects x as [B, N, 14] (the reference's reshape), got "
                                    f"{tuple(xc.shape)}")
            axis = 0
        else:
            n_scenes, n_per, axis = 1, xc.numel() // 14, 1
        cols = torch.empty(max(n_scenes, 1) * 8, dtype=torch.float64, device=xc.device) if axis == 0 else None
        g = torch.empty_like(xc)
        _lib.check(_lib.lib().lgm_activate_forward(ops._stream(), n_scenes, n_per, _lib.ptr(xc), _lib.ptr(g), axis,
                                                   _lib.ptr(cols)), "lgm_activate_forward")
        ops.launch_counter["kernels"] += (2 if axis == 0 else 1) if xc.numel() else 0
        ctx.save_for_backward(xc)
        ctx.shape3 = (n_scenes, n_per, axis)
        return g

    @staticmethod
    def backward(ctx, dg):
        (xc,) = ctx.saved_tensors
        n_scenes, n_per, axis = ctx.shape3
        dgc = dg.contiguous().float()
        dx = torch.empty_like(xc)
        cols = torch.empty(max(n_scenes, 1) * 8, dtype=torch.float64, device=xc.device) if axis == 0 else None
        _lib.check(_lib.lib().lgm_activate_backward(ops._stream(), n_scenes, n_per, _lib.ptr(xc), _lib.ptr(dgc), _lib.ptr(dx),
                                                    axis, _lib.ptr(cols)), "lgm_activate_backward")
        ops.launch_counter["kernels"] += (2 if axis == 0 else 1) if xc.numel() else 0
        return dx, None


def activate_gaussians(x, rot_axis="reference"):
    """x [B, N, 14] raw network output (channels last, as after `x.permute(0, 1, 3, 4, 2).reshape(B, -1, 14)`) ->
    Gaussians [B, N, 14] = (pos 3, opacity 1, scale 3, rotation 4, rgb 3), the input of GaussianRenderer.render.

    rot_axis="reference" (default) reproduces the reference exactly: `self.rot_act = F.normalize` is called WITHOUT a
    dim (/root/reference/core/models.py:43,112), so torch's default dim=1 normalises every quaternion component over
    the N Gaussians of its scene — not each quaternion to unit length.  Reference checkpoints were trained under that
    behaviour.  rot_axis="quaternion" gives the per-quaternion normalisation (dim=-1) instead."""
    if rot_axis not in ROT_AXES:
        raise _lib.LgmError(f"rot_axis must be 'reference' or 'quaternion', got {rot_axis!r}")
    return _ActivateGaussians.apply(x, ROT_AXES[rot_axis])
