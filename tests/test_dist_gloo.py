"""world_size = 2 on CPU (gloo): the autograd collective of the view-sharded renderer — identity forward, one
all-reduce(sum) of the Gaussian gradient in backward — with a stand-in differentiable 'render' (the CUDA path needs
a GPU; its sharded run is covered by bench.py --gpus N and tests marked gpu)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lgm_b200.dist import partition_views, replicate_for_view_sharding, scene_blocks, shard_views
        torch.manual_seed(0)
        B, N, V = 2, 16, 5
        g_full = torch.randn(B, N, 14)
        cams = torch.randn(B, V, 4, 4)
        # rank 1 starts from garbage and receives the Gaussians by broadcast from rank 0
        g_in = g_full.clone() if rank == 0 else torch.zeros_like(g_full)
        g_in.requires_grad_(True)
        g = replicate_for_view_sharding(g_in, None, broadcast_src=0)
        assert torch.equal(g.detach(), g_full)
        vm, _, _, scene, (b, e) = shard_views(cams, cams, torch.zeros(B, V, 3), rank, world)
        assert (b, e) == partition_views(B * V, world, rank)
        # stand-in per-view "render": view j of scene s -> sum(g[s] * w_j)
        w = vm.sum(-1)
        loss = sum((g[int(scene[j])] * w[j]).sum() for j in range(e - b))
        loss.backward()
        # expected: gradient of the sum over ALL views
        w_all = cams.reshape(B * V, 16).sum(-1)
        exp = torch.zeros_like(g_full)
        for j in range(B * V):
            exp[j // V] += w_all[j]
        ok = torch.allclose(g_in.grad, exp, rtol=1e-5, atol=1e-5)
        # whole scenes per rank (B = 2, world = 2): the all-gather path must give the same sum
        assert scene_blocks(B, V, world) == 1 and scene_blocks(3, V, world) is None and scene_blocks(4, 3, 2) == 2
        g2 = g_full.clone().requires_grad_(True)
        gg = replicate_for_view_sharding(g2, None, None, scene_blocks(B, V, world))
        sum((gg[int(scene[j])] * w[j]).sum() for j in range(e - b)).backward()
        ok = ok and torch.allclose(g2.grad, exp, rtol=1e-5, atol=1e-5)
        # producer-only: rank 0 holds the Gaussians and is the only one that wants the gradient -> scatter + gather
        g3 = (g_full.clone() if rank == 0 else torch.full_like(g_full, float("nan"))).requires_grad_(True)
        g3r = replicate_for_view_sharding(g3, None, 0, scene_blocks(B, V, world), producer_only=True)
        assert torch.equal(g3r.detach(), g_full[rank:rank + 1])  # exactly the scenes this rank renders arrived
        sum((g3r[int(scene[j]) - rank] * w[j]).sum() for j in range(e - b)).backward()
        if rank == 0:
            ok = ok and torch.allclose(g3.grad, exp, rtol=1e-5, atol=1e-5)
        else:
            ok = ok and g3.grad is None  # only the producer receives the gradient
        # ONE scene whose views straddle the ranks (the north_star's partition, BASELINE.json configs[4]): no whole-scene
        # blocks exist, so the exchange must be the all-reduce; every rank ends with the gradient of ALL views
        cams1 = torch.randn(1, 4, 4, 4)
        g4 = (g_full[:1].clone() if rank == 0 else torch.full_like(g_full[:1], float("nan"))).requires_grad_(True)
        assert scene_blocks(1, 4, world) is None
        g4r = replicate_for_view_sharding(g4, None, 0, scene_blocks(1, 4, world))
        vm1, _, _, scene1, (b1, e1) = shard_views(cams1, cams1, torch.zeros(1, 4, 3), rank, world)
        assert (b1, e1) == (2 * rank, 2 * rank + 2) and int(scene1.max()) == 0
        w1 = vm1.sum(-1)
        sum((g4r[0] * w1[j]).sum() for j in range(e1 - b1)).backward()
        exp1 = torch.zeros_like(g_full[:1]) + cams1.reshape(4, 16).sum()
        ok = ok and torch.allclose(g4.grad, exp1, rtol=1e-5, atol=1e-5)
        q.put((rank, bool(ok), (b, e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_view_sharded_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    blocks = sorted(b for _, _, b in res)
    assert blocks == [(0, 5), (5, 10)]
