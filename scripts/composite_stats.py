#!/usr/bin/env python
"""Developer tool: per-hit statistics of the compositing kernels on one step of a bench workload.

Builds a second copy of the library with -DLGM_STATS (lgm_b200/liblgm_b200_stats.so; the counters cost time, so the
shipped library never carries them), renders one step forward + backward and prints, per kernel, how many candidate
(warp, Gaussian) evaluations there were, how many of them had at least one contributing pixel, how many pixels those
were, and for the backward the distribution of the number of lanes with a valid pixel.

    python scripts/composite_stats.py [--workload zero123g] [--kind trained] [--out gpurun_out/composite_stats.json]
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_stats_lib():
    from lgm_b200 import build as b
    objdir = os.path.join(b.HERE, "build_stats")
    os.makedirs(objdir, exist_ok=True)
    out = os.path.join(b.HERE, "liblgm_b200_stats.so")
    newest = max(os.path.getmtime(f) for f in b._deps() + [os.path.join(b.CSRC, s) for s in b.SOURCES])
    if os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    objs = []
    for src in b.SOURCES:
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        r = subprocess.run([b._nvcc()] + b.NVCC_FLAGS + ["-DLGM_STATS", "-c", os.path.join(b.CSRC, src), "-o", o],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        objs.append(o)
    r = subprocess.run([b._nvcc(), "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="zero123g")
    ap.add_argument("--kind", default="trained")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "composite_stats.json"))
    ap.add_argument("--build-only", action="store_true")
    args = ap.parse_args()
    path = build_stats_lib()
    if args.build_only:
        print(path)
        return
    import torch
    from lgm_b200 import _lib
    _lib.LIB_PATH = path
    from bench import WORKLOADS
    from lgm_b200 import GaussianRenderer, default_options
    from lgm_b200.synthetic import make_bg, make_cameras, make_gaussians, make_upstream_grads

    B, V, N, S, fovy, _ = WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    opt = default_options(output_size=S, fovy=fovy)
    r = GaussianRenderer(opt, device=dev)
    g = make_gaussians(B, N, args.kind).to(dev).requires_grad_(True)
    cv, cvp, cp = (t.to(dev) for t in make_cameras(B, V, fovy=fovy))
    bg = make_bg().to(dev)
    d_img, d_alpha, _ = (t.to(dev) for t in make_upstream_grads(B, V, S, S))
    lib = _lib.lib()
    lib.lgm_debug_stats.restype = ctypes.c_int
    lib.lgm_debug_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
    buf = (ctypes.c_ulonglong * 32)()
    lib.lgm_debug_stats(None, 1)
    out = r.render(g, cv, cvp, cp, bg_color=bg)
    ((out["image"] * d_img).sum() + (out["alpha"] * d_alpha).sum()).backward()
    torch.cuda.synchronize()
    assert lib.lgm_debug_stats(buf, 0) == 0
    s = list(buf)
    res = {
        "workload": args.workload, "kind": args.kind, "views": B * V, "pixels": B * V * S * S,
        "fwd": {"staged_instances": s[3], "candidates": s[0], "hits": s[1], "composited_pixels": s[2],
                "pixels_per_hit": s[2] / max(s[1], 1), "hit_rate": s[1] / max(s[0], 1)},
        "bwd": {"staged_instances": s[11], "candidates": s[8], "hits": s[9], "valid_pixels": s[10],
                "pixels_per_hit": s[10] / max(s[9], 1), "hit_rate": s[9] / max(s[8], 1),
                "hits_by_valid_lanes": dict(zip(["1", "2", "3-4", "5-8", "9-16", "17-32"], s[16:22]))},
    }
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
