"""View-sharded rendering over the GPUs of one box: every rank renders a contiguous block of the step's
B*V views with the local CUDA path; the per-Gaussian gradients are combined once per step, inside the autograd graph,
by the cheapest exchange that fits (all-reduce; all-gather of per-scene blocks; scatter + gather when one rank produces
the Gaussians; or, fused, the preprocess-backward kernel writing straight into the producer's peer-mapped buffer) —
NCCL over NVLink on GPUs, gloo in the CPU tests of the host logic.  Images stay on the rank that rendered them.  The
reference itself never shards views (each DDP rank renders its own batch, /root/reference/main.py:82,102); this is the
north_star's addition.
"""
import torch
import torch.distributed as dist


def partition_views(n_items, world_size, rank):
    """Contiguous block [begin, end) of ceil(n/world) items for `rank` (last ranks may be short or empty)."""
    per = (n_items + world_size - 1) // world_size
    b = min(n_items, rank * per)
    return b, min(n_items, b + per)


def scene_blocks(n_scenes, views_per_scene, world_size):
    """If the view partition gives every rank the same number of WHOLE scenes, return that number, else None.
    (Then rank r's gradient is non-zero only in scenes [r k, (r+1) k): an all-gather of the owned blocks equals the
    all-reduce of the zero-padded tensors at a fraction of the traffic.)"""
    n = n_scenes * views_per_scene
    per = (n + world_size - 1) // world_size
    if per == 0 or per % views_per_scene or n % world_size or n_scenes % world_size:
        return None
    return per // views_per_scene


_symm_cache = {}


def _producer_buffer(shape, device, group, src):
    """A [B,N,14] float32 buffer in SYMMETRIC memory (torch.distributed._symmetric_memory: every rank's copy is mapped
    into every process of the node over NVLink) and this process's view of rank `src`'s copy.  Cached per shape.
    Returns (handle, local buffer, view of src's buffer) or None when symmetric memory is not available."""
    key = (tuple(shape), str(device), id(group), src)
    hit = _symm_cache.get(key)
    if hit is None:
        try:
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            import warnings
            with warnings.catch_warnings():  # needed by some releases, deprecated (a no-op) in others
                warnings.simplefilter("ignore")
                try:
                    symm_mem.enable_symm_mem_for_group(grp.group_name)
                except Exception:
                    pass
            local = symm_mem.empty(*shape, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(local, grp)
            peer = hdl.get_buffer(src, tuple(shape), torch.float32)
            hit = (hdl, local, peer)
        except Exception as e:  # no symmetric memory on this system / backend: the caller uses the NCCL gather
            hit = (None, None, repr(e))
        _symm_cache[key] = hit
    return None if hit[0] is None else hit


class _ReplicatedInput(torch.autograd.Function):
    """Identity in forward (optionally a broadcast from `src`); sum of the ranks' gradients in backward, so the
    collective sits in the autograd graph of `gaussians` exactly once per step.  The sum is an all-reduce in general;
    when every rank owns whole scenes (`block` scenes each) it is an all-gather of the owned blocks.

    producer_only (needs `src` and whole-scene blocks): the Gaussians exist on `src` alone and only `src` wants their
    gradient.  Then every rank is SENT just the scenes it renders (scatter) — the function returns that block,
    [block,N,14] — and returns just its block of the gradient to `src`: with a gather, or, when a peer-memory sink is
    in use (ShardedGaussianRenderer(peer_gradients=True)), not at all — the preprocess-backward kernel has already
    written the block into `src`'s buffer over NVLink, and the backward here is a device-side barrier.  On the other
    ranks the input receives no gradient."""

    @staticmethod
    def forward(ctx, x, group, src, block, producer_only, peer):
        ctx.group, ctx.block, ctx.src, ctx.peer = group, block, src, peer
        multi = dist.is_initialized() and dist.get_world_size(group) > 1
        ctx.producer_only = bool(producer_only and multi and src is not None and block is not None and
                                 x.shape[0] == block * dist.get_world_size(group))
        ctx.full_shape = tuple(x.shape)
        if src is not None and multi:
            rank = dist.get_rank(group)
            if ctx.producer_only:
                x_in = x.detach().contiguous()
                x = torch.empty((block,) + tuple(x_in.shape[1:]), dtype=x_in.dtype, device=x_in.device)
                dist.scatter(x, scatter_list=list(x_in.split(block)) if rank == src else None, src=src, group=group)
                return x
            x = x.contiguous().clone()
            dist.broadcast(x, src=src, group=group)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            world, rank = dist.get_world_size(ctx.group), dist.get_rank(ctx.group)
            if ctx.producer_only and ctx.peer is not None:
                hdl, local, peer = ctx.peer
                sink = peer[rank * ctx.block:(rank + 1) * ctx.block]
                # Normally g IS this rank's block of src's buffer, written by K7 through the peer mapping: nothing to
                # send.  (A render that had to split its views into chunks sums the chunk gradients locally: copy.)
                if g.data_ptr() != sink.data_ptr():
                    sink.copy_(g)
                hdl.barrier(channel=0)  # all blocks have landed before src reads them
                return (local if rank == ctx.src else None), None, None, None, None, None
            g = g.contiguous()
            if ctx.producer_only:
                out = torch.empty(ctx.full_shape, dtype=g.dtype, device=g.device) if rank == ctx.src else None
                dist.gather(g, gather_list=list(out.split(ctx.block)) if rank == ctx.src else None, dst=ctx.src,
                            group=ctx.group)
                return (out if rank == ctx.src else None), None, None, None, None, None
            if ctx.block is not None and g.shape[0] == ctx.block * world:
                out = torch.empty_like(g)
                dist.all_gather_into_tensor(out, g[rank * ctx.block:(rank + 1) * ctx.block].contiguous(), group=ctx.group)
                g = out
            else:
                g = g.clone()
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None, None, None, None, None


def replicate_for_view_sharding(gaussians, group=None, broadcast_src=None, scenes_per_rank=None, producer_only=False,
                                peer=None):
    return _ReplicatedInput.apply(gaussians, group, broadcast_src, scenes_per_rank, producer_only, peer)


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
    of the sum of all ranks' losses w.r.t. `gaussians` (one collective: all-reduce, or all-gather when every rank
    owns whole scenes).  With `broadcast_src` and `producer_only=True` the Gaussians are taken from rank `broadcast_src`
    alone and the summed gradient is delivered to that rank alone (scatter + gather instead of broadcast + all-gather
    when every rank owns whole scenes).

    peer_gradients=True (with producer_only): the gather is fused into the preprocess-backward kernel — every rank's
    K7 writes its block of dL/dgaussians directly into the producer's buffer through symmetric (peer-mapped) memory, so
    the transfer over NVLink overlaps the kernel, and the step ends with a device-side barrier instead of a collective.
    The gradient the producer receives is then a view of a communication buffer that the next step with the same shape
    overwrites — consume it (optimizer step, copy) before rendering again."""

    def __init__(self, opt, device="cuda", group=None, peer_gradients=False):
        from .renderer import GaussianRenderer
        self.inner = GaussianRenderer(opt, device=device)
        self.group = group
        self.peer_gradients = peer_gradients

    def render(self, gaussians, cam_view, cam_view_proj, cam_pos, bg_color=None, scale_modifier=1, broadcast_src=None,
               producer_only=False, return_depth=True, max_views_per_call=None):
        from . import ops
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        block = scene_blocks(cam_view.shape[0], cam_view.shape[1], world)
        blockwise = bool(producer_only and world > 1 and broadcast_src is not None and block is not None)
        peer, sink = None, None
        if blockwise and self.peer_gradients and gaussians.is_cuda and torch.is_grad_enabled():
            peer = _producer_buffer(tuple(gaussians.shape), gaussians.device, self.group, broadcast_src)
            if peer is not None:
                sink = peer[2][rank * block:(rank + 1) * block]
        # how the gradient travels back (for reports)
        self.exchange = ("none (single rank)" if world == 1 else
                         "scatter + K7 writes into the producer's peer-mapped buffer, device barrier" if sink is not None else
                         "scatter + gather to the producer" if blockwise else
                         "all-gather of per-scene blocks" if block is not None else "all-reduce")
        g = replicate_for_view_sharding(gaussians.contiguous().float(), self.group, broadcast_src, block, producer_only, peer)
        vm, pm, _cp, scene, (b, e) = shard_views(cam_view, cam_view_proj, cam_pos, rank, world)
        if blockwise:
            scene = scene - rank * block  # g holds this rank's scenes only
        S = int(self.inner.opt.output_size)
        bg = (self.inner.bg_color if bg_color is None else bg_color).to(g.device).float().reshape(3).contiguous()
        cfg = ops.ViewConfig(S, S, float(self.inner.tan_half_fov), float(self.inner.tan_half_fov), float(scale_modifier),
                             clamp_image=True, want_depth=bool(return_depth))  # clamp: core/gs.py:87, fused
        image, alpha, depth, _ = ops.render_views(g, vm.to(g.device), pm.to(g.device), scene, bg, cfg,
                                                  max_views_per_call=max_views_per_call, grad_sink=sink)
        return {"image": image, "alpha": alpha, "depth": depth, "views": (b, e)}
