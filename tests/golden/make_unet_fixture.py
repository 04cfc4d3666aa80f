"""Generates tests/golden/config0_unet_splatter.npz from the REFERENCE's own network code — BASELINE.json configs[0]
("tiny CPU run: LGM UNet on 4 synthetic 256^2 views -> 4 x 64 x 64 = 16,384 Gaussians, decode + 1 view 256^2").

Runs only where /root/reference exists (the build container; the GPU box has no copy), on the CPU:
  reference `UNet(9, 14, ...)` of the `tiny` preset (/root/reference/core/unet.py:234, core/options.py:107-120, built as
  /root/reference/core/models.py:24-31) with torch.manual_seed(0) weights  ->  the 1x1 conv of models.py:34  ->  the
  reshape / permute of models.py:98,107  ->  x [1, 16384, 14], the raw splatter image,
and the reference's five activations + cat applied to it exactly as models.py:40-44,109-115 spells them (including
`rot_act = F.normalize` with its default dim=1)  ->  gaussians [1, 16384, 14].
Stored: x (the input of lgm_b200.activate_gaussians), gaussians (what it must reproduce), the synthetic network input's
seed, and the CPU seconds the UNet forward took here (the "CPU-runnable LGM pipeline" half of the north_star's CPU
baseline; bench.py reports it next to the oracle's render time).  tests/test_config0_plumbing.py consumes the file.

    python tests/golden/make_unet_fixture.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    if not os.path.isdir(os.path.join(REF, "core")):
        print(f"{REF}/core not found: the fixture can only be regenerated where the reference tree exists")
        return 0
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    sys.path.insert(0, REF)
    from core.options import config_defaults      # noqa: E402  (tyro is installed)
    from core.unet import UNet                    # noqa: E402  (xformers optional: plain softmax attention)
    opt = config_defaults["tiny"]
    torch.manual_seed(0)
    unet = UNet(9, 14, down_channels=opt.down_channels, down_attention=opt.down_attention, mid_attention=opt.mid_attention,
                up_channels=opt.up_channels, up_attention=opt.up_attention)       # models.py:24-31
    conv = nn.Conv2d(14, 14, kernel_size=1)                                     # models.py:34
    unet.eval(), conv.eval()
    n_params = sum(p.numel() for p in unet.parameters())
    g = torch.Generator().manual_seed(1)
    # 4 synthetic views: 3 "image" channels (ImageNet-normalised range) + 6 ray-embedding channels (provider_lvis.py:183-197)
    images = torch.cat([torch.randn(1, 4, 3, opt.input_size, opt.input_size, generator=g),
                        torch.randn(1, 4, 6, opt.input_size, opt.input_size, generator=g).clamp(-2, 2)], dim=2)
    B, V, C, H, W = images.shape
    with torch.no_grad():
        t0 = time.time()
        x = unet(images.view(B * V, C, H, W))     # models.py:95-96
        t_unet = time.time() - t0
        x = conv(x)                                # models.py:97
        x = x.reshape(B, 4, 14, opt.splat_size, opt.splat_size)       # models.py:98
        x = x.permute(0, 1, 3, 4, 2).reshape(B, -1, 14)              # models.py:107
        # a random-init network emits values near zero; spread them so that every activation is exercised in its
        # non-linear range (a trained network does that itself).  The scaling is part of the committed input.
        x = (x - x.mean(dim=1, keepdim=True)) / x.std(dim=1, keepdim=True) * torch.tensor(
            [0.6, 0.6, 0.6, 1.5, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.5, 1.5, 1.5]) + torch.tensor(
            [0.0, 0.0, 0.0, -1.0, -3.5, -3.5, -3.5, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
        # the reference's activations, as written (models.py:40-44), then models.py:109-115
        pos_act = lambda t: t.clamp(-1, 1)
        scale_act = lambda t: 0.1 * F.softplus(t)
        opacity_act = lambda t: torch.sigmoid(t)
        rot_act = F.normalize
        rgb_act = lambda t: 0.5 * torch.tanh(t) + 0.5
        pos = pos_act(x[..., 0:3])
        opacity = opacity_act(x[..., 3:4])
        scale = scale_act(x[..., 4:7])
        rotation = rot_act(x[..., 7:11])
        rgbs = rgb_act(x[..., 11:])
        gaussians = torch.cat([pos, opacity, scale, rotation, rgbs], dim=-1)
    out = os.path.join(HERE, "config0_unet_splatter.npz")
    np.savez_compressed(out, x=x.numpy().astype(np.float32), gaussians=gaussians.numpy().astype(np.float32),
                        unet_cpu_seconds=np.float64(t_unet), unet_params=np.int64(n_params),
                        torch_threads=np.int64(torch.get_num_threads()), preset=np.array("tiny"),
                        input_shape=np.array(images.shape), torch_version=np.array(torch.__version__))
    print(f"wrote {out}: x {tuple(x.shape)}, UNet(tiny) {n_params / 1e6:.1f} M params, forward {t_unet:.2f} s on "
          f"{torch.get_num_threads()} CPU threads, {os.path.getsize(out) / 1e6:.2f} MB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
