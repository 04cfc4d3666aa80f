"""View-sharded rendering over the GPUs of one box: every rank holds all Gaussians, renders a
contiguous block of the step's B*V views with the local CUDA path, and the per-Gaussian gradients are combined with
ONE all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).  Images stay on the rank that
rendered them.  The reference itself never shards views (each DDP rank renders its own batch,
/root/reference/main.py:82,102); this is the north_star's addition.
"""
import torch
import torch.distributed as dist


def partition_views(n_items, world_size, rank):
    """Contiguous block [begin, end) of ceil(n/world) items for `rank` (last ranks may be short or empty)."""
    per = (n_items + world_size - 1) // world_size
    b = min(n_items, rank * per)
    return b, min(n_items, b + per)


class _ReplicatedInput(torch.autograd.Function):
    """Identity in forward (optionally a broadcast from `src`); all-reduce(sum) of the gradient in backward, so the
    collective sits in the autograd graph of `gaussians` exactly once per step."""

    @staticmethod
    def forward(ctx, x, group, src):
        ctx.group = group
        if src is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
            x = x.contiguous().clone()
            dist.broadcast(x, src=src, group=group)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        if dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            g = g.clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None, None


def replicate_for_view_sharding(gaussians, group=None, broadcast_src=None):
    return _ReplicatedInput.apply(gaussians, group, broadcast_src)


def shard_views(cam_view, cam_view_proj, cam_pos, rank=None, world_size=None):
    """Flatten [B,V,...] cameras to B*V jobs and return this rank's block plus its scene indices.

    Returns (view_mats [n,16], proj_mats [n,16], cam_pos [n,3], view_scene_cpu [n] int32, (begin, end))."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    B, V = cam_view.shape[:2]
    b, e = partition_views(B * V, world_size, rank)
    vm = cam_view.reshape(B * V, 16)[b:e].contiguous().float()
    pm = cam_view_proj.reshape(B * V, 16)[b:e].contiguous().float()
    cp = cam_pos.reshape(B * V, 3)[b:e].contiguous().float()
    scene = (torch.arange(b, e, dtype=torch.int64) // V).int()
    return vm, pm, cp, scene, (b, e)


class ShardedGaussianRenderer:
    """GaussianRenderer.render semantics with the step's views partitioned over the ranks of `group`.

    render() returns this rank's views only: image [n_local,3,H,W], alpha, depth [n_local,1,H,W] and the (begin,end)
    block of the flattened B*V index space.  Back-propagating any loss on them yields, on EVERY rank, the gradient
    of the sum of all ranks' losses w.r.t. `gaussians` (one all-reduce)."""

    def __init__(self, opt, device="cuda", group=None):
        from .renderer import GaussianRenderer
        self.inner = GaussianRenderer(opt, device=device)
        self.group = group

    def render(self, gaussians, cam_view, cam_view_proj, cam_pos, bg_color=None, scale_modifier=1, broadcast_src=None):
        from . import ops
        g = replicate_for_view_sharding(gaussians.contiguous().float(), self.group, broadcast_src)
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        vm, pm, _cp, scene, (b, e) = shard_views(cam_view, cam_view_proj, cam_pos, rank, world)
        S = int(self.inner.opt.output_size)
        bg = (self.inner.bg_color if bg_color is None else bg_color).to(g.device).float().reshape(3).contiguous()
        cfg = ops.ViewConfig(S, S, float(self.inner.tan_half_fov), float(self.inner.tan_half_fov), float(scale_modifier),
                             clamp_image=True)  # core/gs.py:87, fused
        image, alpha, depth, _ = ops.render_views(g, vm.to(g.device), pm.to(g.device), scene, bg, cfg)
        return {"image": image, "alpha": alpha, "depth": depth, "views": (b, e)}
