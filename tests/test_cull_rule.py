"""The culling rule of the compositing kernels (lgm_b200/csrc/composite_common.cuh, patch_mask): which of a tile's four
8x8 pixel patches can a Gaussian reach with alpha >= 1/255?  The kernels test the ellipse d^T Q d <= 2 ln(255 o) against
the rectangle of a patch's pixel centres EXACTLY — the minimum of the convex form over a box lies on the faces that separate
the box from the centre, so it is the smaller of two clamped 1-D minimisations.  This test restates the rule in numpy
(float32, same margins) and checks, on random anisotropic Gaussians, that it is CONSERVATIVE against a brute-force
evaluation of the kernels' pinned per-pixel arithmetic (no patch with a contributing pixel is ever culled — the parity
requirement: a skipped pair must be one the reference would evaluate and discard), that it is tight (within 1 % of
the brute-force answer), and that it is never looser than the bounding-box test it replaced.  CPU only."""
import numpy as np

F = np.float32
CULL_SCALE, CULL_PAD, CULL_PIX = F(1.002), F(2e-3), F(0.02)


def _random_gaussians(n, seed):
    rng = np.random.default_rng(seed)
    # 2D covariance R diag(s1^2, s2^2) R^T + 0.3 I (the low-pass filter of A.1), inverted to the conic
    s1, s2 = np.exp(rng.uniform(-1.0, 3.5, n)), np.exp(rng.uniform(-1.0, 3.5, n))
    th = rng.uniform(0.0, np.pi, n)
    c, s = np.cos(th), np.sin(th)
    a = c * c * s1 ** 2 + s * s * s2 ** 2 + 0.3
    b = c * s * (s1 ** 2 - s2 ** 2)
    d = s * s * s1 ** 2 + c * c * s2 ** 2 + 0.3
    det = a * d - b * b
    cx, cy, cz = (d / det).astype(F), (-b / det).astype(F), (a / det).astype(F)
    op = rng.uniform(0.002, 1.0, n).astype(F)
    gx, gy = rng.uniform(-40, 56, n).astype(F), rng.uniform(-40, 56, n).astype(F)  # centre relative to the tile origin
    return cx, cy, cz, op, gx, gy


def _exact_mask(cx, cy, cz, op, gx, gy):
    k = F(255.0) * op
    t2 = (F(2.0) * np.log(k.astype(np.float64)).astype(F) * CULL_SCALE + CULL_PAD).astype(F)
    kx, ky, cy2 = -cy / cx, -cy / cz, cy + cy
    ax, ay = np.abs(gx) + F(16), np.abs(gy) + F(16)
    t2p = (F(4e-6) * (cx * ax * ax + cz * ay * ay) + t2).astype(F)
    m = np.zeros(len(cx), np.uint32)
    for r in range(2):
        y_lo, y_hi = gy - F(r * 8 + 7) - CULL_PIX, gy - F(r * 8) + CULL_PIX
        Y = np.minimum(np.maximum(F(0), y_lo), y_hi)
        for c in range(2):
            x_lo, x_hi = gx - F(c * 8 + 7) - CULL_PIX, gx - F(c * 8) + CULL_PIX
            X = np.minimum(np.maximum(F(0), x_lo), x_hi)
            dy = np.minimum(np.maximum(X * ky, y_lo), y_hi)
            q1 = (cz * dy + cy2 * X) * dy + cx * X * X
            dx = np.minimum(np.maximum(Y * kx, x_lo), x_hi)
            q2 = (cx * dx + cy2 * Y) * dx + cz * Y * Y
            m |= ((~(q1 > t2p)) | (~(q2 > t2p))).astype(np.uint32) << np.uint32(2 * r + c)
    return np.where(k > 1, m, 0).astype(np.uint32)


def _bbox_mask(cx, cy, cz, op, gx, gy):
    k = F(255.0) * op
    t2 = (F(2.0) * np.log(k.astype(np.float64)).astype(F) * CULL_SCALE + CULL_PAD).astype(F)
    inv = np.maximum(t2, F(0)) / (cx * cz - cy * cy)   # (opacities at or below 1/255 are masked out below)
    hx, hy = np.sqrt(cz * inv) * CULL_SCALE + CULL_PIX, np.sqrt(cx * inv) * CULL_SCALE + CULL_PIX
    m = np.zeros(len(cx), np.uint32)
    for r in range(2):
        row = ~(gy + hy < r * 8) & ~(gy - hy > r * 8 + 7)
        for c in range(2):
            col = ~(gx + hx < c * 8) & ~(gx - hx > c * 8 + 7)
            m |= (row & col).astype(np.uint32) << np.uint32(2 * r + c)
    return np.where(k > 1, m, 0).astype(np.uint32)


def _brute_mask(cx, cy, cz, op, gx, gy):
    """Per pixel, the kernels' pinned power (splat_math.cuh pair_power: fma(fma(cx dx, dx, (cz dy) dy), -0.5, -(cy dx) dy))
    and the reference's test alpha = o exp(power) >= 1/255, power <= 0."""
    m = np.zeros(len(cx), np.uint32)
    for r in range(2):
        for c in range(2):
            hit = np.zeros(len(cx), bool)
            for py in range(r * 8, r * 8 + 8):
                for px in range(c * 8, c * 8 + 8):
                    dx, dy = (gx - F(px)).astype(F), (gy - F(py)).astype(F)
                    t1, t2 = (cx * dx).astype(F), ((cz * dy).astype(F) * dy).astype(F)
                    s = (t1.astype(np.float64) * dx.astype(np.float64) + t2.astype(np.float64)).astype(F)      # fma
                    u = ((cy * dx).astype(F) * dy).astype(F)
                    power = (s.astype(np.float64) * -0.5 - u.astype(np.float64)).astype(F)                       # fma
                    alpha = op.astype(np.float64) * np.exp(power.astype(np.float64))
                    hit |= (power <= 0) & (alpha >= 1.0 / 255.0 * (1 - 1e-6))
            m |= hit.astype(np.uint32) << np.uint32(2 * r + c)
    return np.where(F(255.0) * op > 1, m, 0).astype(np.uint32)


def _bits(m):
    return int(np.unpackbits(m.view(np.uint8)).sum())


def test_exact_patch_cull_is_conservative_and_tight():
    cx, cy, cz, op, gx, gy = _random_gaussians(60000, seed=0)
    exact, bbox, brute = (f(cx, cy, cz, op, gx, gy) for f in (_exact_mask, _bbox_mask, _brute_mask))
    assert np.count_nonzero(brute & ~exact) == 0, "a patch with a contributing pixel was culled"
    assert np.count_nonzero(exact & ~bbox) == 0, "the exact test kept a patch the bounding box excludes"
    n_brute, n_exact, n_bbox = _bits(brute), _bits(exact), _bits(bbox)
    assert n_brute > 20000
    assert n_exact <= 1.01 * n_brute, (n_exact, n_brute)      # tight: the excess is the rounding margins
    assert n_bbox >= 1.3 * n_brute, (n_bbox, n_brute)         # what the bounding-box test let through


def test_exact_patch_cull_degenerate_inputs():
    """Opacity at or below 1/255: never visible.  A Gaussian centred inside a patch always reaches it."""
    cx, cy, cz, op, gx, gy = _random_gaussians(2000, seed=1)
    op_low = np.full_like(op, F(1.0 / 255.0))
    assert not _exact_mask(cx, cy, cz, op_low, gx, gy).any()
    gx_in, gy_in = np.full_like(gx, F(3.5)), np.full_like(gy, F(12.25))   # inside patch (row 1, column 0) = bit 2
    m = _exact_mask(cx, cy, cz, np.full_like(op, F(0.5)), gx_in, gy_in)
    assert ((m >> np.uint32(2)) & 1).all()
