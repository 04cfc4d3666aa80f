"""Developer experiment: time the onesweep sort with parts disabled (results are WRONG on purpose; timing only).
Builds in build/timing/liblgm_ablate_*.so are produced by an ad-hoc patch of a COPY of radix_sort.cu."""
import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lgm_b200 import _lib, ops
from lgm_b200.synthetic import make_bg, make_cameras, make_gaussians
dev = "cuda:0"
B, V, N, S, fovy = 8, 26, 98304, 320, 60.0
g = make_gaussians(B, N, "trained").to(dev)
cv, cvp, _ = make_cameras(B, V, fovy=fovy)
t = math.tan(0.5 * math.radians(fovy))
cfg = ops.ViewConfig(S, S, t, t, 1.0, keep_binning=True)
scene = torch.arange(B, dtype=torch.int32).repeat_interleave(V)
off = torch.arange(0, B * V + 1, V, dtype=torch.int32)
_, _, _, st = ops.forward_views(g, cv.reshape(-1, 16).to(dev), cvp.reshape(-1, 16).to(dev), scene.to(dev), off.to(dev), make_bg().to(dev), cfg)
Lr = st.num_rendered
perm = torch.sort(st.vals[:Lr].long() & 0xFFFFFFFF, stable=True).indices
keys_u, vals_u = st.keys[:Lr][perm].contiguous(), st.vals[:Lr][perm].contiguous()
end_bit = ops.sort_end_bit(B * V * 400)
base = os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "build", "timing")
for name in ["product"] + sorted(f for f in os.listdir(base) if f.startswith("liblgm_ablate")):
    L = _lib.lib() if name == "product" else ctypes.CDLL(os.path.join(base, name))
    for fn in ("lgm_sort_pairs", "lgm_sort_workspace_bytes", "lgm_sort_input_is_tmp"):
        getattr(L, fn).restype, getattr(L, fn).argtypes = _lib._SIGNATURES[fn]
    in_tmp = bool(L.lgm_sort_input_is_tmp(end_bit))
    nb = ctypes.c_size_t(0)
    L.lgm_sort_workspace_bytes(Lr, end_bit, nb)
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    ko, vo = torch.empty_like(keys_u), torch.empty_like(vals_u)
    ts = []
    for it in range(5):
        kin, vin = keys_u.clone(), vals_u.clone()
        a = (ko, vo, kin, vin) if in_tmp else (kin, vin, ko, vo)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.lgm_sort_pairs(torch.cuda.current_stream().cuda_stream, *[_lib.ptr(x) for x in a], Lr, end_bit, 1, _lib.ptr(ws), nb.value)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    print(f"{name:28s} rc {rc}  sort {sum(ts)/len(ts):.2f} ms")
