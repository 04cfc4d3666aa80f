#!/usr/bin/env python
"""Developer probe: does the HBM / latency-bound half of a step (preprocess + binning) overlap with the issue-bound half
(compositing) when two halves of the batch run as independent pipelines on two CUDA streams?

Runs the headline step (a) as one batch on one stream and (b) as two half batches driven by two host threads, each on
its own stream (the halves drift out of phase, so one half's binning meets the other half's compositing), and prints
the views/s of both.  python scripts/overlap_probe.py [--steps 20] [--groups 2]
"""
import argparse
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from lgm_b200 import GaussianRenderer, default_options  # noqa: E402
from lgm_b200.synthetic import make_bg, make_cameras, make_gaussians, make_upstream_grads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--groups", type=int, default=2)
    ap.add_argument("--workload", default="zero123g")
    args = ap.parse_args()
    B, V, N, S, fovy, _ = WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    r = GaussianRenderer(default_options(output_size=S, fovy=fovy), device=dev)
    g = make_gaussians(B, N, "trained").to(dev)
    cv, cvp, cp = (t.to(dev) for t in make_cameras(B, V, fovy=fovy))
    bg = make_bg().to(dev)
    d_img, d_alpha, _ = (t.to(dev) for t in make_upstream_grads(B, V, S, S))

    def step(b0, b1):
        gd = g[b0:b1].detach().requires_grad_(True)
        out = r.render(gd, cv[b0:b1], cvp[b0:b1], cp[b0:b1], bg_color=bg)
        torch.autograd.backward([out["image"], out["alpha"]], [d_img[b0:b1], d_alpha[b0:b1]])
        return gd.grad

    def run(groups, steps):
        bounds = [(i * B // groups, (i + 1) * B // groups) for i in range(groups)]
        streams = [torch.cuda.Stream(device=dev) for _ in bounds]

        def worker(i):
            with torch.cuda.stream(streams[i]):
                for _ in range(steps):
                    step(*bounds[i])
                streams[i].synchronize()

        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if groups == 1:
            worker(0)
        else:
            th = [threading.Thread(target=worker, args=(i,)) for i in range(groups)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return B * V * steps / dt, dt / steps * 1e3

    run(1, 3)
    run(args.groups, 3)
    for groups in (1, args.groups, 1, args.groups):
        v, ms = run(groups, args.steps)
        print(f"groups {groups}: {v:9.0f} views/s  {ms:7.3f} ms per {B * V}-view step")


if __name__ == "__main__":
    main()
