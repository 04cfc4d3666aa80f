#!/usr/bin/env python
"""Multi-GPU check of the view-sharded renderer (run under torchrun on >= 2 GPUs; not a pytest: the GPU test tier has one
GPU).  On rank 0 the gradient of a seeded loss w.r.t. the Gaussians must agree between
  (a) the single-GPU GaussianRenderer over all views,
  (b) ShardedGaussianRenderer with broadcast + all-gather,
  (c) producer_only (scatter + gather),
  (d) producer_only with peer_gradients (K7 writes into rank 0's symmetric buffer, device barrier),
and, for the north_star's partition proper — ONE scene whose views straddle the ranks (B = 1, V = 2 x world; also
B = 3 with V chosen so that rank boundaries fall inside scenes): Gaussians broadcast from rank 0, views rendered locally,
per-Gaussian gradients combined by the NCCL all-reduce (lgm_b200/dist.py, _ReplicatedInput.backward) — every rank must
end with the single-GPU gradient.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lgm_b200 import GaussianRenderer, default_options  # noqa: E402
from lgm_b200.dist import ShardedGaussianRenderer  # noqa: E402
from lgm_b200.synthetic import make_cameras, make_gaussians  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    B, V, N, S = 2 * world, 3, 6000, 96
    opt = default_options(output_size=S)
    g0 = make_gaussians(B, N, "trained", seed=7)
    g0[:, :, 4:7] *= 4.0
    g0 = g0.to(dev)
    cv, cvp, cp = [t.to(dev) for t in make_cameras(B, V, seed=7)]
    w_all = torch.randn(B * V, 3, S, S, generator=torch.Generator().manual_seed(3)).to(dev)

    def run(renderer, **kw):
        g = (g0 if rank == 0 else torch.full_like(g0, float("nan"))).clone().requires_grad_(True)
        out = renderer.render(g, cv, cvp, cp, **kw)
        b, e = out.get("views", (0, B * V))
        (out["image"].reshape(-1, 3, S, S) * w_all[b:e]).sum().backward()
        torch.cuda.synchronize()
        return None if g.grad is None else g.grad.clone()

    ref = run(GaussianRenderer(opt, device=dev)) if rank == 0 else None
    res = {
        "broadcast + all-gather": run(ShardedGaussianRenderer(opt, device=dev), broadcast_src=0),
        "scatter + gather": run(ShardedGaussianRenderer(opt, device=dev), broadcast_src=0, producer_only=True),
    }
    peer = ShardedGaussianRenderer(opt, device=dev, peer_gradients=True)
    res["peer-memory"] = run(peer, broadcast_src=0, producer_only=True)
    res["peer-memory, second step"] = run(peer, broadcast_src=0, producer_only=True)
    ok = True
    if rank == 0:
        scale = ref.abs().amax(dim=(0, 1), keepdim=True).clamp_min(1e-20)
        for name, g in res.items():
            err = ((g - ref).abs() / scale).max().item()
            print(f"{name:28s} max error / column scale = {err:.2e}   ({peer.exchange if name.startswith('peer') else ''})")
            ok = ok and err <= 1e-4
    else:
        assert res["scatter + gather"] is None and res["peer-memory"] is None  # only the producer gets the gradient

    # ---- the all-reduce branch: scenes whose views straddle ranks ----
    for (Bs, Vs, chunk) in ((1, 2 * world, None), (3, 2 * world + 1, None), (1, 4 * world, 2)):
        gs = make_gaussians(Bs, N, "trained", seed=11)
        gs[:, :, 4:7] *= 4.0
        gs = gs.to(dev)
        cvs, cvps, cps = [t.to(dev) for t in make_cameras(Bs, Vs, seed=11)]
        ws = torch.randn(Bs * Vs, 3, S, S, generator=torch.Generator().manual_seed(5)).to(dev)
        wa = torch.randn(Bs * Vs, 1, S, S, generator=torch.Generator().manual_seed(6)).to(dev)

        def run_s(renderer, **kw):
            # every rank but 0 starts from garbage: the broadcast must deliver rank 0's Gaussians
            g = (gs if rank == 0 or not kw else torch.full_like(gs, float("nan"))).clone().requires_grad_(True)
            out = renderer.render(g, cvs, cvps, cps, **kw)
            b, e = out.get("views", (0, Bs * Vs))
            ((out["image"].reshape(-1, 3, S, S) * ws[b:e]).sum() + (out["alpha"].reshape(-1, 1, S, S) * wa[b:e]).sum()).backward()
            torch.cuda.synchronize()
            return g.grad.clone()

        ref_s = run_s(GaussianRenderer(opt, device=dev))   # every rank computes the single-GPU gradient itself
        sh = ShardedGaussianRenderer(opt, device=dev)
        got = run_s(sh, broadcast_src=0, max_views_per_call=chunk)
        scale = ref_s.abs().amax(dim=(0, 1), keepdim=True).clamp_min(1e-20)
        err = torch.tensor([((got - ref_s).abs() / scale).max().item()], device=dev)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"B={Bs} V={Vs} chunk={chunk}: {sh.exchange:12s} max error / column scale over all ranks = {err.item():.2e}")
            assert sh.exchange == "all-reduce", sh.exchange
        ok = ok and err.item() <= 1e-4
    if rank == 0:
        print("DIST CHECK", "PASSED" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
