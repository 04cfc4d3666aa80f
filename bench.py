#!/usr/bin/env python
"""bench.py — rendered views/sec, forward + backward, of the splat-render path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--kind trained|init] [--workload zero123g|lgm_big]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the path's CPU implementation (oracle port) on the host cores

A step = one pass of the hot path over one batch of synthetic input: GaussianRenderer.render of all B x V views
(forward), then backward from given upstream gradients to dL/d[B,N,14] (clamp mask included), as
/root/reference/core/models.py:141 + main.py:102 drive it.  Workload at N = 1 ("zero123g", BASELINE.json configs[2],
the configuration the north_star target is quoted on): 8 scenes x 26 views, 98,304 Gaussians per scene, 320^2,
fovy 60.  At N GPUs the step has 8N scenes (weak scaling): the B*V views are partitioned across ranks; rank 0 produces the
Gaussians and wants their gradient, so every rank is sent the scenes it renders and returns its block of
the gradient (lgm_b200.dist, producer_only: NCCL scatter + gather; all-reduce / all-gather when every rank needs it).

Prints ONE JSON line (rank 0).  `value` = whole-job views/s with inputs resident in HBM; `e2e` = the same metric
with, every step, the step's inputs (Gaussians, cameras, ground-truth images and masks) copied from pinned host
memory — every rank copies what IT renders — and the loss read back; `roofline` = the HBM-bound group K1 + binning:
the bytes the implemented algorithm must move / CUDA-event time against MEASURED_PEAKS.json (SURVEY.md 8d's figure,
which counts a radix sort this path does not run, and the onesweep sort alone are reported beside it); `issue_roofline`
= the compositing kernels against the warp-instruction issue ceiling that bounds them; `cpu_baseline` = the CPU oracle
port on a bounded sample; `gpu_baseline` = the reference-shaped CUDA rasterizer (baseline/) driven per view as
core/gs.py:42-93 drives the external package; `scale_sweep` = BASELINE.json configs[4] (ONE scene of 1 M Gaussians, 256
views at 1024^2) STRONG-scaled over the N ranks as the north_star partitions it: Gaussians broadcast from rank 0, views
rendered locally, gradients combined by one NCCL all-reduce.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scenes per GPU, views per scene, Gaussians per scene, image size, fovy, BASELINE.json config)
    "zero123g": (8, 26, 98304, 320, 60.0, "configs[2] Zero-1-to-G: 6x128^2 splatter = 98,304 Gaussians, 26 views at 320^2, batch 8"),
    "lgm_big": (1, 8, 65536, 512, 49.1, "configs[1] LGM default: 65,536 Gaussians, 8 views at 512^2, batch 1"),
    "tiny": (1, 1, 16384, 256, 49.1, "configs[0] tiny: 16,384 Gaussians, 1 view at 256^2"),
    # the two multi-GPU configurations, expressed per GPU (weak scaling: x N scenes / views at N GPUs)
    "sharded_step": (4, 20, 98304, 320, 60.0, "configs[3] view-sharded step: 32 scenes x 20 views at 320^2 over 8 GPUs = 4 scenes x 20 views per GPU"),
    "scale_sweep": (1, 32, 1000000, 1024, 49.1, "configs[4] scale sweep: 1M Gaussians, 256 views at 1024^2 over 8 GPUs = 32 views per GPU"),
    # the inference orbit of /root/reference/infer.py:113-145 (180 azimuths of one object) as ONE batched call
    "orbit": (1, 180, 65536, 512, 49.1, "inference orbit (infer.py:113-145): 65,536 Gaussians, 180 views at 512^2, forward only"),
}
METRIC = "rendered views/sec fwd+bwd"
UNIT = "views/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="zero123g", choices=sorted(WORKLOADS))
    ap.add_argument("--kind", default="trained", choices=["trained", "init"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stages", action="store_true")
    ap.add_argument("--forward-only", action="store_true", help="time render() under torch.no_grad() (implied by --workload orbit)")
    ap.add_argument("--no-peer-gradients", action="store_true",
                    help="N > 1: return the gradient blocks with an NCCL gather instead of K7 writing into the producer's peer-mapped buffer")
    ap.add_argument("--torch-loss", action="store_true", help="e2e: autograd's mse_loss instead of lgm_b200.mse_image_alpha_loss")
    ap.add_argument("--sort-sweep", action="store_true", help="also time every onesweep launch shape (LGM_SORT_VARIANT)")
    ap.add_argument("--loop-baseline", action="store_true",
                    help="also time the reference's per-view driver loop (core/gs.py:42-93) on the same kernels")
    ap.add_argument("--no-scale-sweep", action="store_true", help="skip the configs[4] strong-scaling leg")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-shaped CUDA baseline (baseline/)")
    ap.add_argument("--sweep-views", type=int, default=256, help="views of the configs[4] leg (256 = BASELINE.json)")
    ap.add_argument("--sweep-gaussians", type=int, default=1000000)
    ap.add_argument("--sweep-chunk", type=int, default=32, help="views per launch group in the configs[4] leg (all N alike)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(workload, kind, steps, warmup, budget_s=20.0):
    """The oracle port (oracle/splat_oracle.c, OpenMP over views) on a bounded sample of the workload."""
    import numpy as np
    from oracle.oracle import Oracle, build
    from lgm_b200.synthetic import make_bg, make_cameras, make_gaussians, make_upstream_grads
    build()
    o = Oracle("f32")
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    o.set_num_threads(ncpu)  # all the host threads it can use (torchrun pins OMP_NUM_THREADS=1 otherwise)
    cores = o.num_threads()
    B, V, N, S, fovy, _ = WORKLOADS[workload]
    nv = max(1, min(V, cores))  # one scene, up to one view per thread
    g = make_gaussians(1, N, kind, seed=1234).numpy()
    cv, cvp, _ = make_cameras(1, V, fovy=fovy, seed=1234)
    cv, cvp = cv[:, :nv].numpy(), cvp[:, :nv].numpy()
    d_img, d_alpha, d_depth = [x[:, :nv].numpy() for x in make_upstream_grads(1, V, S, S, seed=1234)]
    bg = make_bg().numpy()
    t = math.tan(0.5 * math.radians(fovy))

    def step():
        o.render_step(g, cv, cvp, bg, S, S, t, t, 1.0, d_img, d_alpha, d_depth)

    t0 = time.time()
    step()  # first call also sizes the sample
    one = time.time() - t0
    steps = max(1, min(steps, int(budget_s / max(one, 1e-3))))
    for _ in range(max(0, min(warmup, 1))):
        step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = (time.time() - t0) / steps
    return dict(value=nv / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"1 scene x {nv} of {V} views ({workload}, {kind}-like), fwd+bwd, {steps} timed steps, "
                       f"oracle/splat_oracle.c with OpenMP over views"), dt, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, dt, steps = cpu_reference_run(args.workload, args.kind, args.steps, args.warmup, budget_s=60.0)
    B, V, N, S, fovy, cfgname = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cfgname}; {args.kind}-like Gaussians; bounded sample: {cb['sample']}",
                   "note": "the reference's rasterizer (ashawkey/diff-gaussian-rasterization) is CUDA-only and not "
                           "available offline; this arm times the CPU oracle port of the same path on the host cores"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_counters():
    """ncu-measured per-launch counters of the shipped build on the headline workload (profiles/*_counters.json: DRAM bytes
    and executed warp instructions per kernel, from one `ncu --set full` capture; see profiles/*_summary.md)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if f.endswith("_counters.json"):
            best = os.path.join(pdir, f)
    if best is None:
        return None, None
    try:
        return json.load(open(best)), os.path.relpath(best, ROOT)
    except Exception:
        return None, None


def run_native(args):
    import torch
    import torch.distributed as dist
    from lgm_b200 import GaussianRenderer, default_options, mse_image_alpha_loss, ops, _lib
    from lgm_b200.dist import ShardedGaussianRenderer, partition_views
    from lgm_b200.synthetic import make_bg, make_cameras, make_gaussians

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    _lib.lib()  # fail loudly now if the extension is missing

    Bg, V, N, S, fovy, cfgname = WORKLOADS[args.workload]
    B = Bg * world
    opt = default_options(output_size=S, fovy=fovy)
    renderer = ShardedGaussianRenderer(opt, device=dev, peer_gradients=not args.no_peer_gradients)
    local_renderer = GaussianRenderer(opt, device=dev)
    n_views_total = B * V
    b0, b1 = partition_views(n_views_total, world, rank)
    n_local = b1 - b0
    # Device-resident inputs of the HBM-resident `value`: rank 0 produces the Gaussians of the whole step (the other
    # ranks' copies are only shapes: the step scatters rank 0's); every rank holds the cameras.
    g_dev = make_gaussians(B, N, args.kind, seed=1234).to(dev) if rank == 0 else torch.zeros(B, N, 14, device=dev)
    cv, cvp, cp = make_cameras(B, V, fovy=fovy, seed=1234)
    cv_dev, cvp_dev, cp_dev = cv.to(dev), cvp.to(dev), cp.to(dev)
    bg = make_bg().to(dev)
    # Host-side (pinned) inputs of `e2e`: what THIS rank renders — its own Bg scenes, their cameras and ground truth, as
    # the reference's data-parallel loader hands every rank its own batch (/root/reference/main.py:82-102).
    g_host = make_gaussians(Bg, N, args.kind, seed=1234 + rank * Bg).pin_memory()
    sl = slice(rank * Bg, (rank + 1) * Bg)
    cv_host, cvp_host, cp_host = cv[sl].contiguous().pin_memory(), cvp[sl].contiguous().pin_memory(), cp[sl].contiguous().pin_memory()
    gen = torch.Generator().manual_seed(4321 + rank)
    gt_img_host = torch.rand(Bg, V, 3, S, S, generator=gen).pin_memory()
    gt_mask_host = (torch.rand(Bg, V, 1, S, S, generator=gen) > 0.5).float().pin_memory()
    # upstream gradients of the loss shape 2 (x - gt) / numel (core/models.py:153), resident
    d_img = (2.0 * (torch.rand(n_local, 3, S, S, generator=gen) - torch.rand(n_local, 3, S, S, generator=gen)) / (n_views_total * 3 * S * S)).to(dev)
    d_alpha = (2.0 * (torch.rand(n_local, 1, S, S, generator=gen) - torch.rand(n_local, 1, S, S, generator=gen)) / (n_views_total * S * S)).to(dev)
    src = 0 if world > 1 else None

    forward_only = args.forward_only or args.workload == "orbit"
    if forward_only:
        args.no_e2e, args.no_stages, args.no_cpu_baseline, args.no_scale_sweep, args.no_gpu_baseline = True, True, True, True, True

    def step_resident():
        if forward_only:
            with torch.no_grad():
                return renderer.render(g_dev, cv_dev, cvp_dev, cp_dev, bg_color=bg, broadcast_src=src, producer_only=True)["image"]
        g = g_dev.detach().requires_grad_(True)
        out = renderer.render(g, cv_dev, cvp_dev, cp_dev, bg_color=bg, broadcast_src=src, producer_only=True)
        torch.autograd.backward([out["image"], out["alpha"]], [d_img, d_alpha])
        return g.grad

    copy_stream = torch.cuda.Stream(device=dev)
    w_i, w_a = 1.0 / (n_views_total * 3 * S * S), 1.0 / (n_views_total * S * S)   # the loss is normalised over the whole job
    gt8 = {}

    def host_set(u8):
        if not u8:
            return (g_host, cv_host, cvp_host, cp_host, gt_img_host, gt_mask_host)
        if not gt8:  # the ground truth as an image file holds it: 8 bit
            gt8["img"] = (gt_img_host * 255).round().to(torch.uint8).pin_memory()
            gt8["mask"] = (gt_mask_host * 255).to(torch.uint8).pin_memory()
        return (g_host, cv_host, cvp_host, cp_host, gt8["img"], gt8["mask"])

    def render_loss_backward(gd, cvd, cvpd, cpd, gt_i, gt_m):
        g = gd.requires_grad_(True)
        out = local_renderer.render(g, cvd, cvpd, cpd, bg_color=bg, return_depth=False)
        if args.torch_loss:
            mse = torch.nn.functional.mse_loss
            loss = mse(out["image"], gt_i, reduction="sum") * w_i + mse(out["alpha"], gt_m, reduction="sum") * w_a
        else:
            loss = mse_image_alpha_loss(out["image"], out["alpha"], gt_i, gt_m, w_image=w_i, w_alpha=w_a)
        loss.backward()
        return float(loss.item())  # D2H read of the step's result

    def step_e2e_serial():
        """Everything of ONE step in sequence: its copies start when the step starts (the ground truth on a side stream
        while the same step's forward renders), nothing overlaps the previous step."""
        main = torch.cuda.current_stream()
        gd, cvd, cvpd, cpd = [x.to(dev, non_blocking=True) for x in (g_host, cv_host, cvp_host, cp_host)]
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            gt_i, gt_m = gt_img_host.to(dev, non_blocking=True), gt_mask_host.to(dev, non_blocking=True)
        main.wait_stream(copy_stream)  # (the loss needs it; the renderer's launches queued before this point do not wait)
        gt_i.record_stream(main)
        gt_m.record_stream(main)
        return render_loss_backward(gd, cvd, cvpd, cpd, gt_i, gt_m)

    def issue_copies(u8=False):
        """All inputs of one step, host -> device on the copy stream; returns the device tensors and an event."""
        with torch.cuda.stream(copy_stream):
            t = [x.to(dev, non_blocking=True) for x in host_set(u8)]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return t, ev

    pending = {False: None, True: None}

    def step_e2e_pipelined(u8=False):
        """The data-loader overlap: every step issues the host->device copy of the NEXT step's inputs (double
        buffered, copy engine) and consumes the set copied while the previous step computed.  Every step still copies one
        full input set from pinned host memory and reads its loss back inside the timed region."""
        main = torch.cuda.current_stream()
        if pending[u8] is None:
            pending[u8] = issue_copies(u8)
        (gd, cvd, cvpd, cpd, gt_i, gt_m), ev = pending[u8]
        main.wait_event(ev)
        copy_stream.wait_stream(main)  # the next set's buffers may reuse memory the main stream has just released
        pending[u8] = issue_copies(u8)
        for t in (gd, cvd, cvpd, cpd, gt_i, gt_m):
            t.record_stream(main)
        return render_loss_backward(gd, cvd, cvpd, cpd, gt_i, gt_m)

    def step_copies_only():
        t, ev = issue_copies()
        torch.cuda.current_stream().wait_event(ev)
        for x in t:
            x.record_stream(torch.cuda.current_stream())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # Timing rule: inputs larger than L2, or L2 flushed between timed iterations.  The per-step working set (geometry
    # rows, gradient rows, images; instances come on top) is far above the 126 MB L2 for the headline workload; small
    # workloads are timed step by step with a 256 MB write between the steps, outside the timed events.
    working_set = n_local * N * (32 + (0 if forward_only else 48)) + n_local * S * S * 4 * (6 if forward_only else 11)
    flush_l2 = working_set < 4 * 126e6
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush_l2 else None

    def timed(fn, steps, warmup, flush=None, reduce="max"):
        flush = flush_l2 if flush is None else flush
        for _ in range(warmup):
            fn()
        barrier()
        k0 = ops.launch_counter["kernels"]
        if flush:
            total = 0.0
            for _ in range(steps):
                flush_buf.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                total += e0.elapsed_time(e1)
            barrier()
            ms = total / steps
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1) / steps
        launches = ops.launch_counter["kernels"] - k0
        if world > 1 and reduce is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX if reduce == "max" else dist.ReduceOp.MIN)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_resident, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    value = n_views_total / (ms * 1e-3)

    e2e = None
    if not args.no_e2e:
        h2d = sum(t.numel() * t.element_size() for t in host_set(False))
        h2d_8 = sum(t.numel() * t.element_size() for t in host_set(True))
        k = max(3, args.steps)
        ms_p, _ = timed(step_e2e_pipelined, k, 3)
        ms_s, _ = timed(step_e2e_serial, k, 3)
        ms_c, _ = timed(step_copies_only, 5, 3)
        ms_8, _ = timed(lambda: step_e2e_pipelined(True), k, 3)
        pending[False] = pending[True] = None
        e2e = {"value": n_views_total / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4 + 16,
               "what": "per rank and step: ITS scenes' Gaussians, cameras and float32 ground-truth images / masks copied from "
                       "pinned host memory (copy engine, double buffered: the copy of step k+1's inputs is issued at the head of "
                       f"step k) -> GaussianRenderer.render -> MSE loss ({'torch autograd' if args.torch_loss else 'lgm_b200.mse_image_alpha_loss'}) "
                       "-> backward -> loss.item(); whole-job views / max-over-ranks step time",
               "per_rank_h2d_gbs_in_step": h2d / (ms_p * 1e-3) / 1e9,
               "serial": {"value": n_views_total / (ms_s * 1e-3), "ms_per_step": ms_s,
                          "what": "the same step with its own copies issued at its head (ground truth on a side stream during "
                                  "the forward): no overlap with the previous step — PCIe-bound"},
               "h2d_alone_ms": ms_c, "h2d_alone_gbs_per_rank": h2d / (ms_c * 1e-3) / 1e9,
               "u8_ground_truth": {"value": n_views_total / (ms_8 * 1e-3), "ms_per_step": ms_8, "h2d_bytes_per_step": int(h2d_8),
                                   "what": "pipelined, with the ground-truth images / masks kept 8-bit on the host (what the image "
                                           "files hold) and read as such by the loss kernel"}}

    # ---- stage breakdown + rooflines, measured live with CUDA events (rank 0) ----
    stages, roofline, issue_roofline, extra = None, None, None, {}
    # (N = 1 only: at N > 1 a step contains collectives, which rank 0 must not enter alone)
    if rank == 0 and world == 1 and not args.no_stages:
        ops.enable_stage_timing(True)
        for _ in range(2):
            step_resident()
        ops.stage_times_ms()
        nrep = 5
        per_rep = []
        for _ in range(nrep):
            step_resident()
            per_rep.append({k: sum(v) for k, v in ops.stage_times_ms().items()})
        ops.enable_stage_timing(False)
        # median over the repetitions: one disturbed repetition (a straggling copy of the e2e legs, an allocator refill)
        # must not move a stage
        stages = {k: sorted(r.get(k, 0.0) for r in per_rep)[nrep // 2] for k in per_rep[0]}
        # instance count and tile statistics of this rank's block
        from lgm_b200.dist import shard_views
        vm, pm, _, scene, _ = shard_views(cv_dev, cvp_dev, cp_dev, rank, world)
        bsc = int(scene[0]); esc = int(scene[-1]) + 1
        cfg = ops.ViewConfig(S, S, float(renderer.inner.tan_half_fov), float(renderer.inner.tan_half_fov), 1.0, keep_binning=True)
        sc = (scene - bsc).int()
        off = torch.zeros(esc - bsc + 1, dtype=torch.int32)
        off[1:] = torch.cumsum(torch.bincount(sc.long(), minlength=esc - bsc), 0).int()
        with torch.no_grad():
            _, _, _, st = ops.forward_views(g_dev[bsc:esc], vm, pm, sc.to(dev), off.to(dev), bg, cfg)
        Lr = st.num_rendered
        n_tiles = ((S + 15) // 16) ** 2
        npass = ops.sort_passes(n_local * n_tiles)
        end_bit = ops.sort_end_bit(n_local * n_tiles)
        # the sort alone, on the step's real keys in EMIT order (ascending view*P+idx, tiles row-major within a
        # Gaussian = a stable sort of the sorted pairs by value), compressed-key mode as in the renderer
        Lb = _lib.lib()
        perm = torch.sort(st.vals[:Lr].long() & 0xFFFFFFFF, stable=True).indices
        keys_u = st.keys[:Lr][perm].contiguous()
        vals_u = st.vals[:Lr][perm].contiguous()
        del perm
        k_other, v_other = torch.empty_like(keys_u), torch.empty_like(vals_u)
        in_tmp = bool(Lb.lgm_sort_input_is_tmp(end_bit))
        import ctypes

        def time_sort():
            nb = ctypes.c_size_t(0)
            _lib.check(Lb.lgm_sort_workspace_bytes(Lr, end_bit, nb), "ws")
            ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
            out = []
            for it in range(5):
                kin, vin = keys_u.clone(), vals_u.clone()
                ko, vo, kt, vt = (k_other, v_other, kin, vin) if in_tmp else (kin, vin, k_other, v_other)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _lib.check(Lb.lgm_sort_pairs(torch.cuda.current_stream().cuda_stream, _lib.ptr(ko), _lib.ptr(vo), _lib.ptr(kt),
                                             _lib.ptr(vt), Lr, end_bit, 1, _lib.ptr(ws), nb.value), "sort")
                b.record()
                torch.cuda.synchronize()
                if it >= 2:
                    out.append(a.elapsed_time(b))
            assert bool((ko[1:] >= ko[:-1]).all()), "standalone sort produced unsorted keys"
            return sum(out) / len(out)

        t_sort = time_sort()
        sort_variants = None
        if args.sort_sweep:
            sort_variants = {}
            keep = os.environ.get("LGM_SORT_VARIANT")
            for vi in range(12):
                os.environ["LGM_SORT_VARIANT"] = str(vi)
                _lib.apply_env_tuning()
                sort_variants[str(vi)] = time_sort()
            if keep is None:
                del os.environ["LGM_SORT_VARIANT"]
            else:
                os.environ["LGM_SORT_VARIANT"] = keep
            _lib.apply_env_tuning()
        peak, peak_src = load_peaks()
        counters, counters_src = load_counters()
        counters = counters or {}
        same_workload = counters.get("workload") == f"{args.workload}/{args.kind}"
        sort_bytes = (npass * 24 + 8) * Lr
        ach = sort_bytes / (t_sort * 1e-3) / 1e9
        onesweep_sort = {"kernel": f"onesweep radix sort alone (histogram + {npass} passes, u64 key + u32 value) — not on the default path",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": int((npass * 23.6 + 8.0) * Lr), "traffic_source": "profiles/r01_summary.md: 23.6 B per pair per pass + 8.0 for the histogram (ncu dram bytes)",
                         "algorithmic_bytes": int(sort_bytes), "ms": t_sort,
                         "per_pass": {"bytes": int(24 * Lr), "ms": t_sort / (npass + 8.0 / 24.0)}}
        # ---- roofline of the HBM-bound group: K1 preprocess + binning, as IMPLEMENTED ----
        # bytes the implemented algorithm must move: K1 44 + 36 per (view, Gaussian) + the 48-byte gradient row
        # it zeroes for the backward; direct binning:
        # count 12 and scatter 16 per (view, Gaussian) + 8 per instance written, per-tile sort 8 read + 4 written per
        # instance, 20 per tile; onesweep path: emit 20 per pair + 12 per instance, sort (24 n_pass + 8) per instance,
        # ranges 8 per instance + 8 per tile.
        P_ = N
        bin_mode = ops.last_bin_mode["mode"]
        if bin_mode == "direct":
            impl_bytes = n_local * (80 + 48 + 12 + 16) * P_ + 20 * Lr + 20 * n_local * n_tiles
        else:
            impl_bytes = n_local * (80 + 48 + 20) * P_ + 12 * Lr + sort_bytes + 8 * Lr + 8 * n_local * n_tiles
        t_grp = stages.get("geom", 0.0) + stages.get("bin_count", 0.0) + stages.get("bin", 0.0)
        ach_grp = impl_bytes / (t_grp * 1e-3) / 1e9 if t_grp > 0 else None
        survey_bytes = n_local * (80 * P_ + 20 * P_) + 12 * Lr + (6 * 24 + 8) * Lr + 8 * Lr + 8 * n_local * n_tiles
        grp_traffic = counters.get("hbm_group_dram_bytes") if same_workload else None
        roofline = {"bound": "hbm", "kernel": f"K1 preprocess + binning ({bin_mode} path): the HBM-bound group of the step",
                    "achieved": ach_grp, "peak": peak, "unit": "GB/s", "frac": ach_grp / peak if ach_grp else None,
                    "traffic": grp_traffic, "traffic_source": counters_src if grp_traffic else None,
                    "peak_source": peak_src, "algorithmic_bytes": int(impl_bytes), "ms": t_grp,
                    "definition": "bytes the implemented algorithm must move (156 B per (view, Gaussian), 48 of them the gradient row K1 zeroes, + 20 B per instance + 20 B "
                                  "per tile on the direct path) / CUDA-event time of the geom + bin_count + bin stages",
                    "survey_8d": {"algorithmic_bytes": int(survey_bytes), "frac": survey_bytes / (t_grp * 1e-3) / 1e9 / peak if t_grp > 0 else None,
                                  "note": "SURVEY.md 8d's figure (100 B per (view, Gaussian) + 172 B per instance) counts a 6-pass LSD radix "
                                          "sort of the instance list that the direct path does not execute; it is an avoided-traffic "
                                          "ratio, not an achieved bandwidth"}}
        lens = (st.ranges[:, 1] - st.ranges[:, 0]).long()
        pair_evals = int(lens.sum()) * 256
        # ---- what bounds compositing: warp-instruction issue (4 schedulers x 148 SMs x clock) ----
        sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
        issue_peak = 148 * 4 * sm_clock * 1e6 / 1e9  # G warp-instructions / s
        ir = {}
        for name, stage in (("composite_fwd", "composite_fwd"), ("composite_bwd", "composite_bwd")):
            winst = (counters.get("warp_instructions") or {}).get(name) if same_workload else None
            t = stages.get(stage)
            if winst and t:
                ir[name] = {"warp_instructions": winst, "ms": t, "achieved": winst / (t * 1e-3) / 1e9, "frac": winst / (t * 1e-3) / 1e9 / issue_peak}
        issue_roofline = {"bound": "issue", "peak": issue_peak, "unit": "G warp-instr/s", "kernels": ir,
                          "peak_source": f"148 SMs x 4 schedulers x {sm_clock:.0f} MHz (sampled under load)",
                          "counters_source": counters_src if ir else None}
        extra = {
            "instances_per_step_rank0": Lr, "instances_per_view": Lr / max(n_local, 1), "sort_variants_ms": sort_variants,
            "bin_mode": bin_mode, "onesweep_sort": onesweep_sort,
            "composite": {"pair_evals_upper_bound": pair_evals,
                          "fwd_gpairs_per_s": pair_evals / (stages.get("composite_fwd", float("nan")) * 1e-3) / 1e9,
                          "bwd_gpairs_per_s": pair_evals / (stages.get("composite_bwd", float("nan")) * 1e-3) / 1e9,
                          "max_tile_len": int(lens.max()), "mean_tile_len": float(lens.float().mean())},
        }
        del st, keys_u, vals_u, k_other, v_other

    # ---- the reference's DRIVER shape on the same kernels: core/gs.py:42-93's Python loop over B and V with one
    # GaussianRasterizer call (and one host readback) per view, clamp, stack, one autograd backward ----
    if rank == 0 and world == 1 and not args.no_stages and args.loop_baseline:
        from lgm_b200 import GaussianRasterizationSettings, GaussianRasterizer

        def step_loop():
            g = g_dev.detach().requires_grad_(True)
            images, alphas = [], []
            tanh = float(renderer.inner.tan_half_fov)
            for b in range(B):
                means3D, opacity = g[b, :, 0:3].contiguous().float(), g[b, :, 3:4].contiguous().float()
                scales, rotations, rgbs = g[b, :, 4:7].contiguous().float(), g[b, :, 7:11].contiguous().float(), g[b, :, 11:].contiguous().float()
                for v in range(V):
                    rs = GaussianRasterizationSettings(
                        image_height=S, image_width=S, tanfovx=tanh, tanfovy=tanh, bg=bg, scale_modifier=1.0,
                        viewmatrix=cv_dev[b, v], projmatrix=cvp_dev[b, v], sh_degree=0, campos=cp_dev[b, v],
                        prefiltered=False, debug=False)
                    img, _radii, _depth, al = GaussianRasterizer(raster_settings=rs)(
                        means3D=means3D, means2D=torch.zeros_like(means3D), shs=None, colors_precomp=rgbs, opacities=opacity,
                        scales=scales, rotations=rotations, cov3D_precomp=None)
                    images.append(img.clamp(0, 1))
                    alphas.append(al)
            images = torch.stack(images, dim=0).view(B * V, 3, S, S)
            alphas = torch.stack(alphas, dim=0).view(B * V, 1, S, S)
            torch.autograd.backward([images, alphas], [d_img, d_alpha])
            return g.grad

        ms_loop, _ = timed(step_loop, 2, 1)
        extra["per_view_loop"] = {"value": n_views_total / (ms_loop * 1e-3), "unit": UNIT, "ms_per_step": ms_loop,
                                  "what": "same kernels driven as /root/reference/core/gs.py:42-93 drives the external "
                                          "rasterizer: Python loop over B x V, one GaussianRasterizer call and one host "
                                          "readback per view, clamp, torch.stack, one backward"}

    # ---- GPU baseline: the reference-shaped CUDA rasterizer (baseline/, a restatement of the external package's
    # algorithm: per-view launches, CUB scan + radix sort, 256-thread tiles, 10 atomics per pair) driven per view ----
    gpu_baseline = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        try:
            from baseline import ref_rasterizer
            gpu_baseline = ref_rasterizer.bench_leg(g_dev, cv_dev, cvp_dev, cp_dev, bg, d_img, d_alpha, S, float(renderer.inner.tan_half_fov),
                                                    value, UNIT)
        except Exception as e:  # the baseline is test infrastructure: its absence must not take the bench down
            gpu_baseline = {"unavailable": f"{type(e).__name__}: {e}"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _, _ = cpu_reference_run(args.workload, args.kind, 3, 0, budget_s=15.0)

    exchange = getattr(renderer, "exchange", "?")
    scale_sweep = None
    if not args.no_scale_sweep:
        g_dev = d_img = d_alpha = flush_buf = None  # free the headline workload's tensors
        torch.cuda.empty_cache()
        scale_sweep = run_scale_sweep(args, dev, rank, world, timed, barrier)

    if rank == 0:
        line = {
            "metric": METRIC if not forward_only else "rendered views/sec fwd (no_grad)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {cfgname}; per GPU {Bg} scenes x {V} views = {Bg * V} views/step, "
                                   f"{args.kind}-like Gaussians (SURVEY.md 8d), step = {B} scenes view-sharded over {world} GPU(s)",
                       "global_views": n_views_total, "gaussians_per_scene": N, "image": f"{S}x{S}",
                       "parallelism": f"view-sharded x{world}" + (f", rank 0 produces the Gaussians and receives their gradient: {exchange}" if world > 1 else ""),
                       "l2": (f"per-step working set {working_set / 1e6:.0f} MB + instances: L2 flushed (256 MB write) between the "
                              "timed steps, each step timed on its own") if flush_l2 else
                             (f"per-step working set ({working_set / 1e9:.1f} GB of geometry, gradient rows and images, plus the "
                              "instance lists) >> 126 MB L2; no explicit flush")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "issue_roofline": issue_roofline,
            "cpu_baseline": cpu_baseline, "gpu_baseline": gpu_baseline, "scale_sweep": scale_sweep, "stages_ms": stages,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_scale_sweep(args, dev, rank, world, timed, barrier):
    """BASELINE.json configs[4] as the north_star partitions it: ONE scene of 1 M Gaussians, 256 views at
    1024^2, STRONG-scaled — rank 0's Gaussians are broadcast, every rank renders its contiguous block of 256 / N views
    with the local kernels, and the [1, 1M, 14] gradient (56 MB) is combined by one NCCL all-reduce inside autograd
    (lgm_b200.dist._ReplicatedInput).  All N use the same launch groups (--sweep-chunk views), so that N = 1 runs the
    same chunks one after the other.  At N > 1 every rank afterwards renders ALL the views alone (no collective): the
    single-GPU time of the same job on the same box, for the efficiency."""
    import torch
    import torch.distributed as dist
    from lgm_b200 import GaussianRenderer, default_options, ops
    from lgm_b200.dist import ShardedGaussianRenderer, partition_views
    from lgm_b200.synthetic import make_bg, make_cameras, make_gaussians
    Vs, Ns, Ss, fovy, chunk = args.sweep_views, args.sweep_gaussians, 1024, 49.1, args.sweep_chunk
    opt = default_options(output_size=Ss, fovy=fovy)
    sharded = ShardedGaussianRenderer(opt, device=dev)
    g0 = make_gaussians(1, Ns, args.kind, seed=4242).to(dev) if rank == 0 else torch.zeros(1, Ns, 14, device=dev)
    cv, cvp, cp = [t.to(dev) for t in make_cameras(1, Vs, fovy=fovy, seed=4242)]
    bg = make_bg().to(dev)
    b0, b1 = partition_views(Vs, world, rank)
    n_local = b1 - b0
    gen = torch.Generator(device=dev).manual_seed(977 + rank)
    d_img = (torch.rand(n_local, 3, Ss, Ss, device=dev, generator=gen) - 0.5) * (4.0 / (Vs * 3 * Ss * Ss))
    d_alpha = (torch.rand(n_local, 1, Ss, Ss, device=dev, generator=gen) - 0.5) * (4.0 / (Vs * Ss * Ss))
    src = 0 if world > 1 else None

    def step():
        g = g0.detach().requires_grad_(True)
        out = sharded.render(g, cv, cvp, cp, bg_color=bg, broadcast_src=src, return_depth=False, max_views_per_call=chunk)
        torch.autograd.backward([out["image"], out["alpha"]], [d_img, d_alpha])
        return g.grad

    steps = max(2, min(args.steps, 5))
    ms, _ = timed(step, steps, 2, flush=False)
    vps = Vs / (ms * 1e-3)
    # stage times of rank 0's share (the library calls only; collectives are outside them)
    ops.enable_stage_timing(True)
    step()
    ops.stage_times_ms()
    step()
    stages = {k: sum(v) for k, v in ops.stage_times_ms().items()}
    ops.enable_stage_timing(False)
    barrier()
    res = {"config": f"configs[4]: 1 scene x {Ns} Gaussians, {Vs} views at {Ss}^2, {args.kind}-like, fwd+bwd, {chunk} views per launch group",
           "scaling": "strong", "n_gpus": world, "views_per_rank": n_local, "value": vps, "unit": UNIT, "ms_per_step": ms,
           "steps": steps, "exchange": sharded.exchange if world > 1 else "none (single rank)",
           "bin_mode": ops.last_bin_mode["mode"], "stages_ms_rank0": stages,
           "limiter": max(stages, key=stages.get) if stages else None}
    if world > 1:
        # the collectives alone, on the gradient's size
        buf = torch.zeros(Ns * 14, device=dev)

        def t_coll(fn, reps=10):
            for _ in range(3):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        t_ar = t_coll(lambda: dist.all_reduce(buf))
        t_bc = t_coll(lambda: dist.broadcast(buf, src=0))
        nbytes = buf.numel() * 4
        res.update({"allreduce_ms": t_ar, "allreduce_bytes": nbytes,
                    "allreduce_bus_gbs": 2.0 * (world - 1) / world * nbytes / (t_ar * 1e-3) / 1e9,
                    "allreduce_bus_gbs_reference": "725 GB/s measured for an 8-rank all-reduce at 1 GiB (B200_PROFILING.md)",
                    "broadcast_ms": t_bc, "collectives_share_of_step": (t_ar + t_bc) / ms})
        del buf
        # single-GPU time of the whole job, on every rank at once (no collective), for the efficiency
        del d_img, d_alpha
        torch.cuda.empty_cache()
        local = GaussianRenderer(opt, device=dev)
        g_all = g0 if rank == 0 else make_gaussians(1, Ns, args.kind, seed=4242).to(dev)
        d_img = (torch.rand(Vs, 3, Ss, Ss, device=dev, generator=gen) - 0.5) * (4.0 / (Vs * 3 * Ss * Ss))
        d_alpha = (torch.rand(Vs, 1, Ss, Ss, device=dev, generator=gen) - 0.5) * (4.0 / (Vs * Ss * Ss))

        def step_alone():
            g = g_all.detach().requires_grad_(True)
            out = local.render(g, cv, cvp, cp, bg_color=bg, return_depth=False, max_views_per_call=chunk)
            torch.autograd.backward([out["image"].view(Vs, 3, Ss, Ss), out["alpha"].view(Vs, 1, Ss, Ss)], [d_img, d_alpha])
            return g.grad

        ms1_max, _ = timed(step_alone, 2, 1, flush=False, reduce="max")
        ms1_min, _ = timed(step_alone, 1, 0, flush=False, reduce="min")
        res.update({"single_gpu_ms_per_step": ms1_min, "single_gpu_ms_per_step_slowest_rank": ms1_max,
                    "single_gpu_value": Vs / (ms1_min * 1e-3),
                    "efficiency": vps / (world * (Vs / (ms1_min * 1e-3))),
                    "efficiency_note": "views/s at N ranks / (N x views/s of ONE GPU rendering all the views with the same launch "
                                       "groups, measured in this same run on every rank at once; the fastest rank's time is used, "
                                       "the conservative choice)"})
    else:
        res.update({"efficiency": 1.0})
    return res


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
