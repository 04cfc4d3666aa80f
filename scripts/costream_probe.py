#!/usr/bin/env python
"""Developer probe: can the latency-bound front half of a launch group (preprocess + count + binning) run CONCURRENTLY
with the issue-bound compositing of another launch group?

Splits the headline step into two groups of scenes, prepares both, then times (CUDA events, 20 repetitions):
  front(B) alone          K1 + count + scatter + tile sort of group B
  fwd(A) alone            compositing forward of group A
  bwd(A) alone            compositing backward of group A
  front(B) || fwd(A)      on two streams (optionally with a high-priority stream for one of them)
  front(B) || bwd(A)
A concurrent time close to max(alone) means the two halves overlap; close to the sum means they serialise.
    python scripts/costream_probe.py
"""
import ctypes
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from lgm_b200 import _lib, ops  # noqa: E402
from lgm_b200.synthetic import make_bg, make_cameras, make_gaussians, make_upstream_grads  # noqa: E402


def main():
    B, V, N, S, fovy, _ = WORKLOADS["zero123g"]
    dev = torch.device("cuda:0")
    L = _lib.lib()
    t = math.tan(0.5 * math.radians(fovy))
    cfg = ops.ViewConfig(S, S, t, t, 1.0, clamp_image=True)
    g = make_gaussians(B, N, "trained").to(dev)
    cv, cvp, _ = make_cameras(B, V, fovy=fovy)
    bg = make_bg().to(dev)
    d_img, d_alpha, _ = make_upstream_grads(B, V, S, S)
    half = B // 2
    groups = []
    for b0, b1 in ((0, half), (half, B)):
        nv = (b1 - b0) * V
        vm = cv[b0:b1].reshape(nv, 16).contiguous().to(dev)
        pm = cvp[b0:b1].reshape(nv, 16).contiguous().to(dev)
        scene = torch.arange(b1 - b0, dtype=torch.int32).repeat_interleave(V).to(dev)
        off = (torch.arange(b1 - b0 + 1, dtype=torch.int32) * V).to(dev)
        gg = g[b0:b1].contiguous()
        img, al, dp, st = ops.forward_views(gg, vm, pm, scene, off, bg, cfg, prepare_backward=True)
        groups.append(dict(g=gg, vm=vm, pm=pm, scene=scene, off=off, img=img, al=al, dp=dp, st=st, nv=nv,
                           d_img=d_img[b0:b1].reshape(nv, 3, S, S).contiguous().to(dev),
                           d_alpha=d_alpha[b0:b1].reshape(nv, 1, S, S).contiguous().to(dev)))
    torch.cuda.synchronize()
    A, Bg = groups
    P = N

    # ---- front(B): K1 + scan + count + ranges scan + scatter + tile sorts, sizes known from the preparation run ----
    stB = Bg["st"]
    prmB = _lib.make_params(half, P, Bg["nv"], S, S, t, t, 1.0)
    nsum = int(L.lgm_num_block_sums(P, Bg["nv"]))
    block_sums = torch.empty(nsum, dtype=torch.int32, device=dev)
    block_offsets = torch.empty(nsum, dtype=torch.int32, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev)
    nb = _lib._sz(0)
    _lib.check(L.lgm_count_workspace_bytes(prmB, nb), "cws")
    count_ws = torch.empty(int(nb.value), dtype=torch.uint8, device=dev)
    # one run to learn longest / coarse
    s0 = torch.cuda.current_stream().cuda_stream
    _lib.check(L.lgm_forward_geom_rows(s0, prmB, _lib.ptr(Bg["g"]), _lib.ptr(Bg["vm"]), _lib.ptr(Bg["pm"]), _lib.ptr(Bg["scene"]),
                                       _lib.ptr(stB.depth), _lib.ptr(stB.radii), _lib.ptr(stB.xy), _lib.ptr(stB.conic_opacity), None,
                                       _lib.ptr(block_sums), _lib.ptr(block_offsets), _lib.ptr(counts), None, _lib.ptr(stB.grad_rows)), "geom")
    _lib.check(L.lgm_forward_count(s0, prmB, _lib.ptr(stB.radii), _lib.ptr(stB.xy), _lib.ptr(stB.ranges), _lib.ptr(count_ws),
                                   count_ws.numel(), _lib.ptr(counts)), "count")
    n_inst, word = (int(v) for v in counts.tolist())
    longest, coarse = word & 0xFFFFFFFF, (word >> 32) & 0xFFFFFFFF
    wsb = _lib._sz(0)
    _lib.check(L.lgm_bin_workspace_bytes(prmB, n_inst, coarse, wsb), "ws")
    workspace = torch.empty(int(wsb.value), dtype=torch.uint8, device=dev)
    vals = torch.empty(n_inst, dtype=torch.int32, device=dev)

    def front_B(stream):
        s = stream.cuda_stream
        _lib.check(L.lgm_forward_geom_rows(s, prmB, _lib.ptr(Bg["g"]), _lib.ptr(Bg["vm"]), _lib.ptr(Bg["pm"]), _lib.ptr(Bg["scene"]),
                                           _lib.ptr(stB.depth), _lib.ptr(stB.radii), _lib.ptr(stB.xy), _lib.ptr(stB.conic_opacity),
                                           None, _lib.ptr(block_sums), _lib.ptr(block_offsets), _lib.ptr(counts), None,
                                           _lib.ptr(stB.grad_rows)), "geom")
        _lib.check(L.lgm_forward_count(s, prmB, _lib.ptr(stB.radii), _lib.ptr(stB.xy), _lib.ptr(stB.ranges), _lib.ptr(count_ws),
                                       count_ws.numel(), _lib.ptr(counts)), "count")
        _lib.check(L.lgm_forward_bin(s, prmB, _lib.ptr(stB.radii), _lib.ptr(stB.xy), _lib.ptr(stB.depth), _lib.ptr(block_offsets),
                                     n_inst, longest, coarse, 0, None, _lib.ptr(vals), _lib.ptr(stB.ranges), _lib.ptr(workspace),
                                     workspace.numel(), _lib.ptr(count_ws), 0), "bin")

    stA = A["st"]
    prmA = _lib.make_params(half, P, A["nv"], S, S, t, t, 1.0)

    def fwd_A(stream):
        _lib.check(L.lgm_forward_composite(stream.cuda_stream, prmA, _lib.ptr(A["g"]), _lib.ptr(A["scene"]), _lib.ptr(stA.xy),
                                           _lib.ptr(stA.conic_opacity), _lib.ptr(stA.depth), _lib.ptr(stA.vals), _lib.ptr(stA.ranges),
                                           _lib.ptr(bg), 1, _lib.ptr(A["img"]), _lib.ptr(A["al"]), _lib.ptr(A["dp"]),
                                           _lib.ptr(stA.n_contrib)), "fwd")

    rowsA = torch.zeros(A["nv"] * P, _lib.GRAD_ROW, dtype=torch.float32, device=dev)

    def bwd_A(stream):
        _lib.check(L.lgm_backward_composite(stream.cuda_stream, prmA, _lib.ptr(A["g"]), _lib.ptr(A["scene"]), _lib.ptr(stA.xy),
                                            _lib.ptr(stA.conic_opacity), _lib.ptr(stA.depth), _lib.ptr(stA.vals), _lib.ptr(stA.ranges),
                                            _lib.ptr(bg), _lib.ptr(A["al"]), _lib.ptr(stA.n_contrib), _lib.ptr(A["d_img"]),
                                            _lib.ptr(A["d_alpha"]), None, _lib.ptr(rowsA)), "bwd")

    lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
    s_norm1, s_norm2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    s_high = torch.cuda.Stream(device=dev, priority=-1)

    def timed(tasks, reps=20):
        """tasks: list of (fn, stream); all launched back to back, time from the first launch to the last completion."""
        main = torch.cuda.current_stream()
        out = []
        for it in range(reps + 3):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(main)
            for _, st_ in tasks:
                st_.wait_event(a)
            for fn, st_ in tasks:
                fn(st_)
            for _, st_ in tasks:
                e = torch.cuda.Event()
                e.record(st_)
                main.wait_event(e)
            b.record(main)
            torch.cuda.synchronize()
            if it >= 3:
                out.append(a.elapsed_time(b))
        out.sort()
        return out[len(out) // 2]

    res = {
        "front(B) alone": timed([(front_B, s_norm1)]),
        "fwd(A) alone": timed([(fwd_A, s_norm1)]),
        "bwd(A) alone": timed([(bwd_A, s_norm1)]),
        "fwd(A) || front(B)  [equal priority, fwd launched first]": timed([(fwd_A, s_norm1), (front_B, s_norm2)]),
        "front(B) || fwd(A)  [equal priority, front launched first]": timed([(front_B, s_norm2), (fwd_A, s_norm1)]),
        "fwd(A) || front(B)  [front on a high-priority stream]": timed([(fwd_A, s_norm1), (front_B, s_high)]),
        "fwd(A) high-priority || front(B)": timed([(fwd_A, s_high), (front_B, s_norm2)]),
        "bwd(A) || front(B)  [equal priority]": timed([(bwd_A, s_norm1), (front_B, s_norm2)]),
        "bwd(A) || front(B)  [front on a high-priority stream]": timed([(bwd_A, s_norm1), (front_B, s_high)]),
    }
    for k, v in res.items():
        print(f"{v:8.3f} ms  {k}")


if __name__ == "__main__":
    main()
