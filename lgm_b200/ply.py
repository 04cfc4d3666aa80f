"""On-disk format of the path's input: 3DGS-compatible binary PLY.

Mirrors /root/reference/core/gs.py:101-190 (`GaussianRenderer.save_ply / load_ply`) without plyfile / kiui:
vertex properties x, y, z, f_dc_0..2, opacity, scale_0..2, rot_0..3, all float32 little-endian, in that order;
`compatible=True` stores pre-activation values (inverse sigmoid opacity, log scale, SH DC colour) as the original
3DGS viewers expect; Gaussians with opacity < 0.005 are pruned on save (core/gs.py:116).
"""
import numpy as np
import torch

SH_C0 = 0.28209479177387814
PROPS = ["x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1",
         "rot_2", "rot_3"]


def _inverse_sigmoid(x, eps=1e-8):
    # kiui.op.inverse_sigmoid: log(x / (1 - x))
    return torch.log(x / (1 - x))


def save_ply(gaussians, path, compatible=True):
    """gaussians [1, N, 14] (pos, opacity, scale, rot wxyz, rgb) -> binary little-endian PLY at `path`."""
    assert gaussians.shape[0] == 1, 'only support batch size 1'
    g = gaussians[0].detach().float().cpu()
    means3D, opacity, scales, rotations, shs = g[:, 0:3], g[:, 3:4], g[:, 4:7], g[:, 7:11], g[:, 11:14]
    mask = opacity.squeeze(-1) >= 0.005  # prune by opacity
    means3D, opacity, scales, rotations, shs = means3D[mask], opacity[mask], scales[mask], rotations[mask], shs[mask]
    if compatible:  # invert the activations (core/gs.py:124-127)
        opacity = _inverse_sigmoid(opacity)
        scales = torch.log(scales + 1e-8)
        shs = (shs - 0.5) / SH_C0
    attrs = torch.cat([means3D, shs, opacity, scales, rotations], dim=1).numpy().astype("<f4")
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {attrs.shape[0]}"]
    header += [f"property float {p}" for p in PROPS] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(np.ascontiguousarray(attrs).tobytes())
    return int(attrs.shape[0])


def load_ply(path, compatible=True):
    """Inverse of save_ply; returns a CPU float32 tensor [N, 14] (as /root/reference/core/gs.py:154-190)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").split("\n")
    fmt = [l for l in lines if l.startswith("format")][0].split()[1]
    n = int([l for l in lines if l.startswith("element vertex")][0].split()[2])
    names, types = [], []
    in_vertex = False
    for l in lines:
        if l.startswith("element"):
            in_vertex = l.startswith("element vertex")
        elif l.startswith("property") and in_vertex:
            _, t, nme = l.split()
            names.append(nme)
            types.append({"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1",
                          "int": "i4", "int32": "i4", "uint": "u4", "short": "i2", "ushort": "u2", "char": "i1"}[t])
    if fmt == "ascii":
        arr = np.loadtxt(data[end:].decode("ascii").split("\n")[:n], dtype=np.float64).reshape(n, len(names))
        col = {nme: arr[:, i] for i, nme in enumerate(names)}
    else:
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(nme, order + t) for nme, t in zip(names, types)])
        rec = np.frombuffer(data, dtype=dt, count=n, offset=end)
        col = {nme: rec[nme].astype(np.float64) for nme in names}
    xyz = np.stack([col["x"], col["y"], col["z"]], axis=1)
    opac = col["opacity"][:, None]
    shs = np.stack([col["f_dc_0"], col["f_dc_1"], col["f_dc_2"]], axis=1)
    scales = np.stack([col[k] for k in names if k.startswith("scale_")], axis=1)
    rots = np.stack([col[k] for k in names if k.startswith("rot_")], axis=1)
    gaussians = torch.from_numpy(np.concatenate([xyz, opac, scales, rots, shs], axis=1)).float()
    if compatible:
        gaussians[..., 3:4] = torch.sigmoid(gaussians[..., 3:4])
        gaussians[..., 4:7] = torch.exp(gaussians[..., 4:7])
        gaussians[..., 11:] = SH_C0 * gaussians[..., 11:] + 0.5
    return gaussians
