#!/usr/bin/env python
"""Developer tool: time of the binning on a scene with thousands of exact duplicates (all keys of a tile tie on depth),
direct path against onesweep.  python scripts/tie_timing.py [copies]"""
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lgm_b200 import ops  # noqa: E402
from lgm_b200.synthetic import make_cameras, make_gaussians  # noqa: E402

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 15000
dev = torch.device("cuda:0")
g = make_gaussians(1, 2000, "trained", seed=8)
g = torch.cat([g, g[:, :1].repeat(1, copies, 1)], dim=1).contiguous().to(dev)
cv, cvp, _ = make_cameras(1, 3, seed=5)
S = 96
t = math.tan(0.5 * math.radians(49.1))
cfg = ops.ViewConfig(S, S, t, t, 1.0)
vm = cv.reshape(-1, 16).contiguous().to(dev)
pm = cvp.reshape(-1, 16).contiguous().to(dev)
scene = torch.zeros(3, dtype=torch.int32, device=dev)
off = torch.tensor([0, 3], dtype=torch.int32, device=dev)
bg = torch.full((3,), 0.5, device=dev)
for mode in ("direct", "onesweep"):
    os.environ["LGM_BIN_MODE"] = mode
    ops.enable_stage_timing(True)
    for _ in range(3):
        with torch.no_grad():
            _, _, _, st = ops.forward_views(g, vm, pm, scene, off, bg, cfg)
    torch.cuda.synchronize()
    tm = ops.stage_times_ms()
    ops.enable_stage_timing(False)
    longest = int((st.ranges[:, 1] - st.ranges[:, 0]).max())
    print(mode, "ran", ops.last_bin_mode["mode"], "instances", st.num_rendered, "longest tile", longest,
          "bin ms", [round(x, 3) for x in tm.get("bin", [])])
