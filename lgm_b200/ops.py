"""Host-side driver of the CUDA splat-render path: tensor allocation, the single per-step readback of the instance
count, and the autograd bridge.  PyTorch here is plumbing (device memory, streams, autograd); all arithmetic runs in
lgm_b200/csrc through the C-ABI of include/lgm_b200.h.  No CPU / PyTorch fallback exists.

Replaces, batched over every view of a step, what `_RasterizeGaussians.forward/backward` and
`RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA` do per view in the external rasterizer
(SURVEY.md §3 call stack A, §8a rows a4-a13), as driven from /root/reference/core/gs.py:42-93.
"""
import dataclasses
from typing import Optional

import torch

from . import _lib

MAX_INSTANCES = (1 << 30) - 1       # look-back counters of the sort are 30 bit
MAX_PAIRS_PER_CALL = 1 << 28        # (view, Gaussian) pairs per launch group: ~22 GB of state at 84 B / pair
MAX_VIEWS_PER_CALL = 65535          # the per-(view, Gaussian) kernels put the view in gridDim.y

# launch accounting for bench.py's "gpu_launches" (kernels this library enqueues; memsets not counted)
launch_counter = {"kernels": 0}
BIN_MODES = {0: "none", 1: "onesweep", 2: "hybrid", 3: "direct"}
last_bin_mode = {"mode": "none", "coarse": False}

# Optional per-stage CUDA-event timing (bench.py's stage breakdown; off in normal use).  When enabled, every stage
# call is bracketed by events on the launching stream; read with stage_times_ms() after a synchronize.
_stage_events = None


def enable_stage_timing(on=True):
    global _stage_events
    _stage_events = {} if on else None


def _timed(name, fn):
    if _stage_events is None:
        return fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    _stage_events.setdefault(name, []).append((a, b))
    return r


def stage_times_ms(reset=True):
    """{stage: [ms per call]} for the calls recorded since the last reset (synchronises)."""
    torch.cuda.synchronize()
    out = {k: [a.elapsed_time(b) for a, b in v] for k, v in (_stage_events or {}).items()}
    if reset and _stage_events is not None:
        _stage_events.clear()
    return out


@dataclasses.dataclass(frozen=True)
class ViewConfig:
    """Per-call constants (GaussianRasterizationSettings minus the per-view matrices)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    scale_modifier: float = 1.0
    keep_binning: bool = False  # keep sorted keys / tiles_touched (parity tests)
    clamp_image: bool = False   # fuse the renderer's clamp(0,1) (core/gs.py:87) and its gradient mask into the kernels
    want_depth: bool = True     # False: no depth image (returned as None); LGM computes it and drops it (core/gs.py:76)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_cuda_f32(t, name, shape_tail=None):
    if not t.is_cuda:
        raise _lib.LgmError(f"{name} must be a CUDA tensor (lgm_b200 has no CPU path)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise _lib.LgmError(f"{name} must be contiguous float32")
    if shape_tail is not None and tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise _lib.LgmError(f"{name} must have trailing shape {shape_tail}, got {tuple(t.shape)}")


class ForwardState:
    """Everything the backward needs (the analogue of upstream's geomBuffer / binningBuffer / imgBuffer)."""
    __slots__ = ("cfg", "n_scenes", "P", "n_views", "view_scene", "scene_view_offsets", "depth", "radii", "xy",
                 "conic_opacity", "vals", "ranges", "n_contrib", "num_rendered", "keys", "tiles_touched", "grad_rows")


def forward_views(gaussians, view_mats, proj_mats, view_scene, scene_view_offsets, bg, cfg: ViewConfig,
                  prepare_backward=False, cov3d=None):
    """All views in one set of launches.  Returns (image [VW,3,H,W], alpha [VW,1,H,W], depth [VW,1,H,W], state).
    prepare_backward: also allocate the backward's gradient rows now and let K1 zero them (lgm_forward_geom_rows) — no
    separate fill of 48 B per (view, Gaussian) pair at the head of the backward.
    cov3d [B,P,6]: upstream's cov3D_precomp — replaces the covariance built from the scale / rotation columns."""
    L = _lib.lib()
    _check_cuda_f32(gaussians, "gaussians", (14,))
    _check_cuda_f32(view_mats, "view_mats", (16,))
    _check_cuda_f32(proj_mats, "proj_mats", (16,))
    _check_cuda_f32(bg, "bg", (3,))
    dev = gaussians.device
    B, P = gaussians.shape[0], gaussians.shape[1]
    VW = view_mats.shape[0]
    H, W = cfg.image_height, cfg.image_width
    prm = _lib.make_params(B, P, VW, H, W, cfg.tanfovx, cfg.tanfovy, cfg.scale_modifier)
    n_tiles = L.lgm_tiles_per_view(H, W)
    bin_mode = _lib.apply_env_tuning()
    st = ForwardState()
    st.cfg, st.n_scenes, st.P, st.n_views = cfg, B, P, VW
    st.view_scene, st.scene_view_offsets = view_scene, scene_view_offsets
    npair = VW * P
    st.depth = torch.empty(npair, dtype=torch.float32, device=dev)
    st.radii = torch.empty(npair, dtype=torch.int32, device=dev)
    st.xy = torch.empty(npair, 2, dtype=torch.float32, device=dev)
    st.conic_opacity = torch.empty(npair, 4, dtype=torch.float32, device=dev)
    st.tiles_touched = torch.empty(npair, dtype=torch.int32, device=dev) if cfg.keep_binning else None
    nsum = int(L.lgm_num_block_sums(P, VW))
    block_sums = torch.empty(max(nsum, 1), dtype=torch.int32, device=dev)
    block_offsets = torch.empty(max(nsum, 1), dtype=torch.int32, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev)  # lgm_step_counts: total instances | longest tile
    s = _stream()
    if cov3d is not None:
        _check_cuda_f32(cov3d, "cov3d", (6,))
        if cov3d.numel() != B * P * 6:
            raise _lib.LgmError(f"cov3d must hold n_scenes * P * 6 values, got {tuple(cov3d.shape)}")
    # the backward's gradient rows: K1 zeroes the row of every (view, Gaussian) pair it processes (no separate fill)
    st.grad_rows = torch.empty(npair, _lib.GRAD_ROW, dtype=torch.float32, device=dev) if (prepare_backward and npair) else None
    _timed("geom", lambda: _lib.check(L.lgm_forward_geom_rows(
        s, prm, _lib.ptr(gaussians), _lib.ptr(view_mats), _lib.ptr(proj_mats), _lib.ptr(view_scene), _lib.ptr(st.depth),
        _lib.ptr(st.radii), _lib.ptr(st.xy), _lib.ptr(st.conic_opacity), _lib.ptr(st.tiles_touched), _lib.ptr(block_sums),
        _lib.ptr(block_offsets), _lib.ptr(counts), _lib.ptr(cov3d), _lib.ptr(st.grad_rows)), "lgm_forward_geom"))
    launch_counter["kernels"] += 2 if npair else 0
    # first half of the binning (per-tile counts and their scan = the final ranges of the direct path): it does not
    # need the instance count, so it runs before the step's readback and leaves the longest tile next to the count
    st.ranges = torch.empty(max(VW * n_tiles, 1), 2, dtype=torch.int32, device=dev)
    count_ws = None
    if bin_mode in (0, 3) and npair:
        nb = _lib._sz(0)
        _lib.check(L.lgm_count_workspace_bytes(prm, nb), "lgm_count_workspace_bytes")
        count_ws = torch.empty(int(nb.value), dtype=torch.uint8, device=dev)
        _timed("bin_count", lambda: _lib.check(L.lgm_forward_count(
            s, prm, _lib.ptr(st.radii), _lib.ptr(st.xy), _lib.ptr(st.ranges), _lib.ptr(count_ws), count_ws.numel(),
            _lib.ptr(counts)), "lgm_forward_count"))
        launch_counter["kernels"] += 2
    # work that does not depend on the instance count is queued before the host waits for it
    image = torch.empty(VW, 3, H, W, dtype=torch.float32, device=dev)
    alpha = torch.empty(VW, 1, H, W, dtype=torch.float32, device=dev)
    depth_img = torch.empty(VW, 1, H, W, dtype=torch.float32, device=dev) if cfg.want_depth else None
    st.n_contrib = torch.empty(VW, H, W, dtype=torch.int32, device=dev)
    # THE host<->device synchronisation of the step (upstream: one per view): 16 bytes — the instance count sizes the
    # instance buffers, the longest tile selects the binning path.  Nothing inside the library synchronises.
    n_inst, word = (int(v) for v in counts.tolist())
    longest = (word & 0xFFFFFFFF) if count_ws is not None else -1
    coarse_entries = ((word >> 32) & 0xFFFFFFFF) if count_ws is not None else 0
    st.num_rendered = n_inst
    if n_inst > MAX_INSTANCES:
        raise TooManyInstances(n_inst)
    ws_bytes = _lib._sz(0)
    _lib.check(L.lgm_bin_workspace_bytes(prm, n_inst, coarse_entries, ws_bytes), "lgm_bin_workspace_bytes")
    # the direct path (every tile fits its shared-memory sort) orders the instances without a key buffer
    direct = count_ws is not None and 0 <= longest <= int(L.lgm_direct_bin_tile_cap())
    keys = None if (direct and not cfg.keep_binning) else torch.empty(max(n_inst, 1), dtype=torch.int64, device=dev)
    st.vals = torch.empty(max(n_inst, 1), dtype=torch.int32, device=dev)
    workspace = torch.empty(max(int(ws_bytes.value), 1), dtype=torch.uint8, device=dev)
    # (lgm_forward_bin_render is these two calls back to back; split here so that stages can be timed)
    _timed("bin", lambda: _lib.check(L.lgm_forward_bin(
        s, prm, _lib.ptr(st.radii), _lib.ptr(st.xy), _lib.ptr(st.depth), _lib.ptr(block_offsets), n_inst, longest, coarse_entries, bin_mode,
        _lib.ptr(keys), _lib.ptr(st.vals), _lib.ptr(st.ranges), _lib.ptr(workspace), workspace.numel(), _lib.ptr(count_ws),
        1 if cfg.keep_binning else 0), "lgm_forward_bin"))
    _timed("composite_fwd", lambda: _lib.check(L.lgm_forward_composite(
        s, prm, _lib.ptr(gaussians), _lib.ptr(view_scene), _lib.ptr(st.xy), _lib.ptr(st.conic_opacity), _lib.ptr(st.depth),
        _lib.ptr(st.vals), _lib.ptr(st.ranges), _lib.ptr(bg), 1 if cfg.clamp_image else 0, _lib.ptr(image), _lib.ptr(alpha),
        _lib.ptr(depth_img), _lib.ptr(st.n_contrib)), "lgm_forward_composite"))
    if n_inst > 0:
        mode = int(L.lgm_last_bin_mode())
        last_bin_mode["mode"] = BIN_MODES.get(mode, "none")
        if mode == 3:    # scatter (plain, or coarse grouping + fine scatter), tile sort (one to four size classes)
            last_bin_mode["coarse"] = bool(L.lgm_last_bin_coarse())
            launch_counter["kernels"] += (2 if last_bin_mode["coarse"] else 1) + 1 + (longest > 2048) + (longest > 5632) + (longest > 11776)
        elif mode == 2:  # emit, histogram, tile-bit passes, ranges, short + long tile sort
            launch_counter["kernels"] += 5 + tile_bit_passes(VW * n_tiles)
        else:            # emit, histogram, ranges + onesweep passes
            launch_counter["kernels"] += 3 + sort_passes(VW * n_tiles)
    launch_counter["kernels"] += 1 if VW else 0                     # compositing
    st.keys = keys if cfg.keep_binning else None
    return image, alpha, depth_img, st


def sort_passes(n_global_tiles):
    return (sort_end_bit(n_global_tiles) + 7) // 8


def tile_bit_passes(n_global_tiles):
    """onesweep passes of the hybrid sort: the (view|tile) bits only (api.cu lgm_forward_bin)."""
    return (sort_end_bit(n_global_tiles) - 31 + 7) // 8


def sort_end_bit(n_global_tiles):
    """Sort bits of the renderer's compressed key: 31 depth bits + global-tile bits (api.cu key_end_bit)."""
    return 31 + max(1, (max(n_global_tiles, 1) - 1).bit_length())


class TooManyInstances(_lib.LgmError):
    def __init__(self, n):
        super().__init__(f"{n} Gaussian instances in one call (limit {MAX_INSTANCES}): split the views into chunks")
        self.n_instances = n


def backward_views(gaussians, view_mats, proj_mats, bg, st: ForwardState, alpha, d_image, d_alpha, d_depth, cov3d=None,
                   d_gauss_out=None):
    """Returns (dL_dgaussians [B,P,14], grad_rows [VW*P,12]).  grad_rows is in MOMENT form (include/lgm_b200.h,
    lgm_backward); screen_gradients() converts it to upstream's dL/dmean2D, dL/dconic, ... .
    With cov3d (the forward's cov3D_precomp) returns (dL_dgaussians, grad_rows, dL_dcov3d [B,P,6]).
    d_gauss_out: where K7 writes dL_dgaussians instead of a fresh tensor — e.g. a peer GPU's memory (lgm_b200.dist)."""
    L = _lib.lib()
    dev = gaussians.device
    cfg = st.cfg
    prm = _lib.make_params(st.n_scenes, st.P, st.n_views, cfg.image_height, cfg.image_width, cfg.tanfovx, cfg.tanfovy,
                           cfg.scale_modifier)
    grad_rows = getattr(st, "grad_rows", None)  # zeroed during the forward when the caller announced a backward
    st.grad_rows = None                          # used once: a second backward over the same state gets fresh rows
    if grad_rows is None:
        grad_rows = torch.zeros(max(st.n_views * st.P, 1), _lib.GRAD_ROW, dtype=torch.float32, device=dev)
    if d_gauss_out is not None and (d_gauss_out.shape != gaussians.shape or d_gauss_out.dtype != torch.float32 or
                                    not d_gauss_out.is_contiguous()):
        raise _lib.LgmError("d_gauss_out must be a contiguous float32 tensor of the shape of gaussians")
    d_gauss = torch.empty_like(gaussians) if d_gauss_out is None else d_gauss_out
    # (lgm_backward is these two calls back to back)
    _timed("composite_bwd", lambda: _lib.check(L.lgm_backward_composite(
        _stream(), prm, _lib.ptr(gaussians), _lib.ptr(st.view_scene), _lib.ptr(st.xy), _lib.ptr(st.conic_opacity),
        _lib.ptr(st.depth), _lib.ptr(st.vals), _lib.ptr(st.ranges), _lib.ptr(bg), _lib.ptr(alpha), _lib.ptr(st.n_contrib),
        _lib.ptr(d_image), _lib.ptr(d_alpha), _lib.ptr(d_depth), _lib.ptr(grad_rows)), "lgm_backward_composite"))
    d_cov = None if cov3d is None else torch.zeros(gaussians.shape[0], gaussians.shape[1], 6, dtype=torch.float32, device=dev)
    _timed("geom_bwd", lambda: _lib.check(L.lgm_backward_geom_cov3d(
        _stream(), prm, _lib.ptr(gaussians), _lib.ptr(view_mats), _lib.ptr(proj_mats), _lib.ptr(st.scene_view_offsets),
        _lib.ptr(st.radii), _lib.ptr(st.conic_opacity), _lib.ptr(grad_rows), _lib.ptr(d_gauss), 0, _lib.ptr(cov3d),
        _lib.ptr(d_cov)), "lgm_backward_geom"))
    launch_counter["kernels"] += 2 if st.n_views * st.P else 0
    if st.n_views * st.P == 0:
        d_gauss.zero_()
    if cov3d is not None:
        return d_gauss, grad_rows, d_cov
    return d_gauss, grad_rows


def screen_gradients(st: ForwardState, grad_rows):
    """Moment rows -> upstream's per-(view, Gaussian) screen-space gradients [VW*P,12]: [0:2] dL/dmean2D (what
    means2D.grad receives upstream), [2:5] dL/dconic, [5] dL/dopacity, [6:9] dL/dcolour, [9] dL/ddepth."""
    L = _lib.lib()
    cfg = st.cfg
    prm = _lib.make_params(st.n_scenes, st.P, st.n_views, cfg.image_height, cfg.image_width, cfg.tanfovx, cfg.tanfovy,
                           cfg.scale_modifier)
    out = torch.zeros_like(grad_rows)
    _lib.check(L.lgm_screen_gradients(_stream(), prm, _lib.ptr(st.conic_opacity), _lib.ptr(grad_rows), _lib.ptr(out)),
               "lgm_screen_gradients")
    launch_counter["kernels"] += 1 if st.n_views * st.P else 0
    return out


def _grad_or_zeros(g, like):
    if g is None:
        return torch.zeros_like(like)
    return g.contiguous().float()


class _RenderViews(torch.autograd.Function):
    """image, alpha, depth, radii = f(gaussians [B,P,14]); gradients flow to gaussians only."""

    @staticmethod
    def forward(ctx, gaussians, view_mats, proj_mats, view_scene, scene_view_offsets, bg, cfg, grad_sink=None):
        image, alpha, depth_img, st = forward_views(gaussians, view_mats, proj_mats, view_scene, scene_view_offsets, bg,
                                                    cfg, prepare_backward=ctx.needs_input_grad[0])
        ctx.st = st
        ctx.grad_sink = grad_sink  # optional destination of dL/dgaussians (not an autograd input: no gradient)
        ctx.set_materialize_grads(False)  # an unused output (depth, in LGM) arrives as None, not as a zero image
        ctx.save_for_backward(gaussians, view_mats, proj_mats, bg, alpha)
        radii = st.radii.view(st.n_views, st.P)
        ctx.mark_non_differentiable(radii)
        return image, alpha, depth_img, radii

    @staticmethod
    def backward(ctx, d_image, d_alpha, d_depth, _d_radii):
        gaussians, view_mats, proj_mats, bg, alpha = ctx.saved_tensors
        st = ctx.st
        d_image = _grad_or_zeros(d_image, alpha.expand(-1, 3, -1, -1))
        d_alpha = _grad_or_zeros(d_alpha, alpha)
        d_depth = None if d_depth is None else d_depth.contiguous().float()  # None -> NULL: no depth gradient
        d_gauss, _ = backward_views(gaussians, view_mats, proj_mats, bg, st, alpha, d_image, d_alpha, d_depth,
                                    d_gauss_out=ctx.grad_sink)
        return d_gauss, None, None, None, None, None, None, None


def _scene_offsets(view_scene_cpu, n_scenes):
    counts = torch.bincount(view_scene_cpu.long(), minlength=n_scenes)
    off = torch.zeros(n_scenes + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(counts, 0).int()
    return off


_view_map_cache = {}


def _device_view_maps(local_scene, n_scenes, device):
    """Device copies of (view_scene, scene_view_offsets), cached: the same maps recur every step."""
    key = (local_scene.numpy().tobytes(), n_scenes, str(device))
    hit = _view_map_cache.get(key)
    if hit is None:
        if len(_view_map_cache) > 256:
            _view_map_cache.clear()
        hit = (local_scene.to(device), _scene_offsets(local_scene, n_scenes).to(device))
        _view_map_cache[key] = hit
    return hit


def render_views(gaussians, view_mats, proj_mats, view_scene_cpu, bg, cfg: ViewConfig,
                 max_views_per_call: Optional[int] = None, grad_sink=None):
    """Differentiable rendering of VW views of B scenes.

    gaussians [B,P,14] cuda fp32; view_mats / proj_mats [VW,16]; view_scene_cpu: CPU int tensor [VW], non-decreasing
    (views of a scene contiguous).  Splits the views into chunks when one call would exceed the library limits
    (pairs, instances); autograd sums the chunk gradients.  Returns image, alpha, depth, radii [VW,P].
    grad_sink [B,P,14]: where the backward writes dL/dgaussians (used when the whole call is one chunk; see lgm_b200.dist).
    """
    VW = view_mats.shape[0]
    B, P = gaussians.shape[0], gaussians.shape[1]
    if VW and bool((view_scene_cpu[1:] < view_scene_cpu[:-1]).any()):
        raise _lib.LgmError("view_scene must be non-decreasing (views of one scene contiguous)")
    chunk = VW if max_views_per_call is None else max_views_per_call
    if P > 0:
        chunk = min(chunk, max(1, MAX_PAIRS_PER_CALL // P))
    chunk = max(min(chunk, MAX_VIEWS_PER_CALL), 1)
    outs, v0 = [], 0
    while v0 < VW or (VW == 0 and not outs):
        v1 = min(VW, v0 + chunk)
        vs = view_scene_cpu[v0:v1]
        b0 = int(vs[0]) if v1 > v0 else 0
        b1 = int(vs[-1]) + 1 if v1 > v0 else B
        local_scene = (vs - b0).int()
        try:
            scene_dev, offsets_dev = _device_view_maps(local_scene, b1 - b0, gaussians.device)
            whole = (b0, b1) == (0, B) and (v0, v1) == (0, VW)
            o = _RenderViews.apply(
                gaussians[b0:b1] if (b0, b1) != (0, B) else gaussians, view_mats[v0:v1], proj_mats[v0:v1],
                scene_dev, offsets_dev, bg, cfg, grad_sink if whole else None)
        except TooManyInstances as e:
            if v1 - v0 <= 1:
                raise
            chunk = max(1, int((v1 - v0) * 0.8 * MAX_INSTANCES / e.n_instances))
            continue
        outs.append(o)
        v0 = v1
        if VW == 0:
            break
    if len(outs) == 1:
        return outs[0]
    return tuple(None if outs[0][i] is None else torch.cat([o[i] for o in outs], 0) for i in range(4))
