"""Developer experiment (not part of the product or the tests): phase timing inside the onesweep pass, using the
instrumented copy build/timing/liblgm_timing.so (clock64 stamps by thread 0 of every tile)."""
import ctypes, math, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lgm_b200 import _lib, ops
from lgm_b200.synthetic import make_bg, make_cameras, make_gaussians

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "build", "timing", "liblgm_timing.so")
L = _lib.lib()
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
in_tmp = bool(L.lgm_sort_input_is_tmp(end_bit))
nb = ctypes.c_size_t(0)
L.lgm_sort_workspace_bytes(Lr, end_bit, nb)
ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
ko, vo = torch.empty_like(keys_u), torch.empty_like(vals_u)
for it in range(2):
    kin, vin = keys_u.clone(), vals_u.clone()
    a = (ko, vo, kin, vin) if in_tmp else (kin, vin, ko, vo)
    _lib.check(L.lgm_sort_pairs(torch.cuda.current_stream().cuda_stream, *[_lib.ptr(x) for x in a], Lr, end_bit, 1, _lib.ptr(ws), nb.value), "sort")
    torch.cuda.synchronize()
tile_items = 4096
nt = min(32768, (Lr + tile_items - 1) // tile_items)
buf = np.zeros((nt, 16), np.uint64)
L.lgm_debug_sort_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = L.lgm_debug_sort_timing(buf.ctypes.data_as(ctypes.c_void_p), nt)
d = buf.astype(np.int64)   # stamps of the LAST pass that ran (clock64 is per SM: only differences within a tile matter)
names = ["ticket+zero", "key load wait", "rank+sync", "prefix+lookback(t0)", "wait others+vals", "scatter", "write issue"]
dur = np.diff(d[:, :8], axis=1)[100:nt - 100]
print("tiles", nt, "rc", rc)
for i, nme in enumerate(names):
    print(f"{nme:22s} median {np.median(dur[:, i]):8.0f}  p90 {np.percentile(dur[:, i], 90):8.0f} cycles")
dd = d[100:nt - 100]
print("  prefix loop + publish ", np.median(dd[:, 8] - dd[:, 3]), " hist load+scan+bar", np.median(dd[:, 9] - dd[:, 8]), " lookback+publish", np.median(dd[:, 4] - dd[:, 9]))
print("total median", np.median(d[100:nt - 100, 7] - d[100:nt - 100, 0]))
