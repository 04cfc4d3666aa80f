"""GPU parity at the sizes that are BENCHMARKED (BASELINE.json configs[1]..[4]): the CUDA path (through the C-ABI)
against the CPU oracle on a few whole views of each configuration — not properties, the oracle itself:

  bit-exact   radii, tiles touched, xy / depth / conic bits, sorted keys, sorted values (order + stability), tile ranges
  counted     n_contrib mismatches (threshold flips of alpha < 1/255 or T(1-alpha) < 1e-4 between ex2.approx on the
              GPU and glibc expf in the oracle; SURVEY.md §7 "threshold-chaotic") — reported, bounded
  <= 1e-4     RGB / alpha max-abs; depth 1e-4 relative
  <= 1e-3     gradients w.r.t. the [N,14] Gaussians against the fp32 oracle's backward started from the same saved
              forward state (alpha image, n_contrib) — strict; end to end against the fp32 and fp64 oracles the distance
              is bounded by the fp32 reference algorithm's own distance to fp64 (it starts from T_final = 1 - alpha: ill
              conditioned on nearly opaque pixels).  Max error over the column's scale, relative L2 and per-element
              relative error percentiles are all reported

Numbers observed on the B200 are written to gpurun_out/parity_fullsize.json (copied to profiles/ by the builder).
PARITY UNPINNED: the oracle restates SURVEY.md Appendix A; see oracle/splat_oracle.c.
"""
import json
import os
import time

import numpy as np
import pytest
import torch

from conftest import ROOT, split14, tan_half
from lgm_b200.synthetic import make_bg, make_cameras, make_gaussians

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REPORT = {}

# name: (Gaussians, image size, fovy, views of the config, which views are compared)
CONFIGS = {
    "configs[1] lgm_big 65,536 @512^2": (65536, 512, 49.1, 8, (0, 5)),
    "configs[2] zero123g 98,304 @320^2": (98304, 320, 60.0, 26, (0, 13)),
    "configs[3] sharded_step 98,304 @320^2 (20 views)": (98304, 320, 60.0, 20, (3, 11)),
    "configs[4] scale_sweep 1M @1024^2": (1000000, 1024, 49.1, 256, (0, 100)),
}


@pytest.fixture(scope="module", autouse=True)
def _dump_report():
    yield
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_fullsize.json"), "w") as f:
            json.dump(REPORT, f, indent=1)
    except OSError:
        pass


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _percentiles(rel):
    q = np.percentile(rel, [50, 90, 99, 99.9, 100])
    return {k: float(v) for k, v in zip(("p50", "p90", "p99", "p99.9", "max"), q)}


@pytest.mark.parametrize("kind", ["trained", "init"])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_benchmarked_config_against_oracle(oracle32, oracle64, name, kind):
    from lgm_b200 import ops
    N, S, fovy, V, picks = CONFIGS[name]
    if kind == "init" and N >= 1000000:
        # untrained Gaussians at 1M / 1024^2 put ~600 M instances on ONE view: minutes of oracle time and
        # beyond the 2^30-instance call limit with two views; the init-like input is checked at the other three sizes
        N = 150000
    g = make_gaussians(1, N, kind, seed=1234)
    cv_all, cvp_all, _ = make_cameras(1, V, fovy=fovy, seed=1234)
    sel = list(picks)
    cv, cvp = cv_all[:, sel].contiguous(), cvp_all[:, sel].contiguous()
    nv = len(sel)
    bg = make_bg().numpy()
    t = tan_half(fovy)
    gd = g.to(DEV).contiguous()
    vm, pm = cv.reshape(nv, 16).to(DEV), cvp.reshape(nv, 16).to(DEV)
    scene = torch.zeros(nv, dtype=torch.int32, device=DEV)
    off = torch.tensor([0, nv], dtype=torch.int32, device=DEV)
    cfg = ops.ViewConfig(S, S, t, t, 1.0, keep_binning=True)
    bgt = torch.as_tensor(bg).to(DEV)
    img, al, dp, st = ops.forward_views(gd, vm, pm, scene, off, bgt, cfg)
    rng = np.random.RandomState(7)
    # loss-shaped upstream gradients (2 (x - G) / numel, scaled up so that fp32 underflow plays no role)
    d_img = ((rng.rand(nv, 3, S, S) - rng.rand(nv, 3, S, S)) * 2.0).astype(np.float32)
    d_alpha = ((rng.rand(nv, 1, S, S) - rng.rand(nv, 1, S, S)) * 2.0).astype(np.float32)
    d_depth = ((rng.rand(nv, 1, S, S) - 0.5) * 2.0).astype(np.float32)
    dg, _ = ops.backward_views(gd, vm, pm, bgt, st, al, torch.as_tensor(d_img).to(DEV), torch.as_tensor(d_alpha).to(DEV),
                               torch.as_tensor(d_depth).to(DEV))
    torch.cuda.synchronize()
    dg = dg[0].cpu().numpy().astype(np.float64)

    means, opac, scales, rots, cols = split14(g[0].numpy())
    ntiles = ((S + 15) // 16) ** 2
    rep = {"gaussians": N, "image": S, "views_compared": sel, "bin_mode": ops.last_bin_mode["mode"], "per_view": []}
    ref64_sum = np.zeros((N, 14))
    ref32_sum = np.zeros((N, 14))
    ref32_same_state = np.zeros((N, 14))   # the fp32 oracle's backward started from the CUDA forward's saved state
    t0 = time.time()
    for v in range(nv):
        args = (means, scales, rots, opac, cols, cv[0, v].numpy(), cvp[0, v].numpy(), bg, S, S, t, t)
        pre, b, f = oracle32.rasterize(*args)
        sl = slice(v * N, (v + 1) * N)
        # ---- bit-exact: geometry, binning ----
        assert np.array_equal(st.radii[sl].cpu().numpy(), pre["radii"])
        assert np.array_equal(st.tiles_touched[sl].cpu().numpy().view(np.uint32), pre["tiles"])
        assert np.array_equal(_bits(st.depth[sl].cpu().numpy()), _bits(pre["depth"]))
        assert np.array_equal(_bits(st.xy[sl].cpu().numpy()), _bits(pre["xy"]))
        assert np.array_equal(_bits(st.conic_opacity[sl].cpu().numpy()), _bits(pre["conic_opacity"]))
        ranges = st.ranges[v * ntiles:(v + 1) * ntiles].cpu().numpy().astype(np.int64)
        ne = ranges[:, 1] > ranges[:, 0]
        assert b["L"] > 0 and ne.any()
        start = int(ranges[ne, 0].min())
        keys = st.keys[start:start + b["L"]].cpu().numpy().view(np.uint64)
        vals = st.vals[start:start + b["L"]].cpu().numpy().view(np.uint32)
        assert np.array_equal(keys - (np.uint64(v * ntiles) << np.uint64(32)), b["keys"]), "sorted keys differ"
        assert np.array_equal(vals - np.uint32(v * N), b["vals"]), "sorted values differ (order / stability)"
        rel = ranges.copy()
        rel[ne] -= start
        assert np.array_equal(rel, b["ranges"].astype(np.int64)), "tile ranges differ"
        # ---- images: 1e-4, with the threshold flips counted ----
        nc = st.n_contrib[v].cpu().numpy().view(np.uint32)
        flips = int((nc != f["n_contrib"]).sum())
        e_img = np.abs(img[v].cpu().numpy() - f["image"])
        e_al = np.abs(al[v].cpu().numpy() - f["alpha"])
        e_dp = np.abs(dp[v].cpu().numpy() - f["depth"]) - 1e-4 * np.abs(f["depth"])
        bad = int((e_img > 1e-4).sum() + (e_al > 1e-4).sum())
        rep["per_view"].append({"view": sel[v], "instances": int(b["L"]), "longest_tile": int((ranges[:, 1] - ranges[:, 0]).max()),
                                "n_contrib_flips": flips, "pixels": int(nc.size), "image_max_abs": float(e_img.max()),
                                "alpha_max_abs": float(e_al.max()), "depth_max_excess_over_1e-4_rel": float(e_dp.max()),
                                "values_above_1e-4": bad})
        # a flip adds or drops one contribution of at most alpha ~ 1/255 .. or the last one before saturation
        assert flips <= max(2, 2e-5 * nc.size), f"{flips} n_contrib mismatches of {nc.size} pixels"
        assert bad <= max(6, 2e-5 * e_img.size), f"{bad} image / alpha values beyond 1e-4"
        assert e_img.max() <= 1.5 / 255 and e_al.max() <= 1.5 / 255 and e_dp.max() <= 4.0 / 255
        # ---- gradients: fp64 oracle (its own forward), accumulated over the compared views ----
        f_cuda = dict(f, alpha=al[v].cpu().numpy(), n_contrib=np.ascontiguousarray(nc & np.uint32(0x1FFFFFFF)))
        for o, acc, fwd_state in ((oracle64, ref64_sum, None), (oracle32, ref32_sum, f), (oracle32, ref32_same_state, f_cuda)):
            p_, b_, f_ = (pre, b, fwd_state) if o is oracle32 else o.rasterize(*args)
            r = o.rasterize_backward(*args, p_, b_, f_, d_img[v], d_alpha[v, 0], d_depth[v, 0])
            acc[:, 0:3] += r["dL_dmeans"]; acc[:, 3] += r["dL_dopacity"]; acc[:, 4:7] += r["dL_dscales"]
            acc[:, 7:11] += r["dL_drots"]; acc[:, 11:14] += r["dL_dcolor"]
    rep["oracle_seconds"] = time.time() - t0
    grads = {}
    for slc, nm in ((slice(0, 3), "means"), (slice(3, 4), "opacity"), (slice(4, 7), "scales"), (slice(7, 11), "rots"),
                    (slice(11, 14), "rgb")):
        ref, got, r32, r32s = ref64_sum[:, slc], dg[:, slc], ref32_sum[:, slc], ref32_same_state[:, slc]
        scale = np.abs(ref).max() + 1e-300
        sig = np.abs(r32) > 1e-3 * scale   # per-element relative error where the element is not negligible against the scale

        def cmp(a, b_):
            rel = np.abs(a - b_)[sig] / np.abs(b_)[sig].clip(1e-300)
            return {"max_err_over_scale": float(np.abs(a - b_).max() / scale),
                    "rel_l2": float(np.linalg.norm(a - b_) / (np.linalg.norm(b_) + 1e-300)),
                    "per_element_rel": _percentiles(rel) if rel.size else None}

        grads[nm] = {
            # THE backward parity: the fp32 oracle's backward (the reference algorithm in the reference's precision and
            # operation order) started from the SAME saved forward state as the CUDA backward (the CUDA forward's alpha
            # image and n_contrib — what the reference's backward reads, Appendix A.5)
            "vs_f32_oracle_same_forward_state": cmp(got, r32s),
            # end to end against the fp32 oracle's own forward + backward.  The alpha images of the two forwards differ by
            # ~3e-7 (ex2.approx vs expf), and the reference's backward starts from T_final = 1 - alpha: on a pixel with
            # T_final ~ 1e-3 that is already 3e-4 relative — the reference ALGORITHM is that ill-conditioned in fp32
            "vs_f32_oracle": cmp(got, r32),
            # against the fp64 oracle (its own forward: other alpha < 1/255 / T < 1e-4 decisions on threshold pixels)
            "vs_f64_oracle": cmp(got, ref),
            "f32_oracle_vs_f64_oracle": cmp(r32, ref),
            "elements": int(sig.sum())}
    rep["gradients"] = grads
    REPORT[f"{name} / {kind}"] = rep
    for nm, gq in grads.items():
        same, a32, a64, n64 = (gq[k] for k in ("vs_f32_oracle_same_forward_state", "vs_f32_oracle", "vs_f64_oracle", "f32_oracle_vs_f64_oracle"))
        # the bar of the north_star (1e-3 relative on accumulated gradients against the reference on identical inputs) for
        # the backward on identical saved state: relative L2 <= 1e-3, strict (observed: ~1e-6 trained-like, <= 5e-4
        # init-like); largest entry-wise error <= 1e-3 of the column's scale — trained-like strict (observed <= 5e-4, and
        # ~1e-6 where no pixel flipped); in the saturated init-like scenes the backward re-decides alpha >= 1/255 per pair
        # with ex2.approx against the oracle's expf, and ONE flipped pair moves the gradients fed by that pixel by up to
        # 1/255 of their size (the same threshold chaos as in the forward): one such contribution, 4e-3, is allowed there
        assert same["rel_l2"] <= 1e-3, (nm, "same forward state", gq)
        assert same["max_err_over_scale"] <= (1e-3 if kind == "trained" else 4e-3), (nm, "same forward state", gq)
        # end to end the CUDA path may be as far from the fp32 reference as the fp32 reference is from the truth (fp64),
        # not farther; and no farther from fp64 than twice the fp32 reference's own distance
        noise, noise_l2 = n64["max_err_over_scale"], n64["rel_l2"]
        assert a32["max_err_over_scale"] <= max(1e-3, noise) and a32["rel_l2"] <= max(1e-3, noise_l2), (nm, "vs f32 oracle", gq)
        assert a64["max_err_over_scale"] <= max(1e-3, 2 * noise) and a64["rel_l2"] <= max(1e-3, 2 * noise_l2), (nm, "vs f64 oracle", gq)
