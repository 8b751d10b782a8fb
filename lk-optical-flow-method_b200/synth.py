"""Deterministic synthetic dashcam-like clips, integer math only.

`zoom` clips have a known focus of expansion (FOE): frame t samples a 2Wx2H texture through the
inverse map  src = s*dst + foe_tex*(1 - s/2),  s = 2 / rate**(t mod period), so the FOE sits at
(0.48 W, 0.60 H) in frame coordinates and a scene cut happens every `period` frames.

Everything after the (seeded, numpy) texture synthesis is fixed-point integer arithmetic, written once
against the operator subset that numpy arrays and torch tensors share, so the same frames are produced
on the host (tests, CPU baseline) and on the device (bench inputs generated in HBM).  No cv2, no files.
"""
import numpy as np

_Q = 12  # fractional bits of the sampling coordinates


def _bilinear_q(T, sxq, syq):
    """T: (Ht, Wt[, C]) uint8-valued integer array; sxq (W,), syq (H,) coordinates in Q12 (int64).
    Returns (H, W[, C]) int64 in 0..255, rounding half up."""
    Ht, Wt = T.shape[0], T.shape[1]
    one = 1 << _Q
    ix = sxq >> _Q
    iy = syq >> _Q
    fx = sxq & (one - 1)
    fy = syq & (one - 1)
    ix0 = ix.clip(0, Wt - 1)
    ix1 = (ix + 1).clip(0, Wt - 1)
    iy0 = iy.clip(0, Ht - 1)
    iy1 = (iy + 1).clip(0, Ht - 1)
    if T.ndim == 3:
        fxb = fx[None, :, None]
        fyb = fy[:, None, None]
    else:
        fxb = fx[None, :]
        fyb = fy[:, None]
    def g(iy_, ix_):
        v = T[iy_[:, None], ix_[None, :]]
        return v.astype(np.int64) if isinstance(v, np.ndarray) else v.long()
    r0 = g(iy0, ix0) * (one - fxb) + g(iy0, ix1) * fxb
    r1 = g(iy1, ix0) * (one - fxb) + g(iy1, ix1) * fxb
    return (r0 * (one - fyb) + r1 * fyb + (1 << (2 * _Q - 1))) >> (2 * _Q)


def _upsample(grid, Ht, Wt):
    """Integer bilinear up-sampling of a small int64 grid to (Ht, Wt)."""
    gh, gw = grid.shape
    sx = (np.arange(Wt, dtype=np.int64) * (((gw - 1) << _Q) // max(Wt - 1, 1)))
    sy = (np.arange(Ht, dtype=np.int64) * (((gh - 1) << _Q) // max(Ht - 1, 1)))
    return _bilinear_q(grid, sx, sy)


def make_texture(W, H, seed=0):
    """(2H, 2W, 3) uint8 BGR texture: value noise over 6 octaves + dark/bright rectangles (road
    markings / windows: they supply Shi-Tomasi corners) + a small high-frequency speckle."""
    rng = np.random.default_rng(seed)
    Ht, Wt = 2 * H, 2 * W
    acc = np.zeros((Ht, Wt), np.int64)
    wsum = 0
    for o, wgt in zip(range(6), (32, 24, 20, 16, 12, 10)):
        cell = max(4, (Wt // 8) >> o)
        gh, gw = Ht // cell + 2, Wt // cell + 2
        acc += wgt * _upsample(rng.integers(0, 256, (gh, gw)).astype(np.int64), Ht, Wt)
        wsum += wgt
    base = acc // wsum                      # 0..255, mid-grey centred
    base = (base - 128) * 3 // 2 + 128      # stretch contrast
    n_rect = 400 + (W * H) // 1500
    for _ in range(n_rect):
        rw = int(rng.integers(max(6, Wt // 300), max(8, Wt // 40)))
        rh = int(rng.integers(max(6, Ht // 300), max(8, Ht // 40)))
        x0 = int(rng.integers(0, Wt - rw))
        y0 = int(rng.integers(0, Ht - rh))
        val = int(rng.integers(0, 70)) if rng.integers(0, 2) else int(rng.integers(185, 256))
        base[y0:y0 + rh, x0:x0 + rw] = (base[y0:y0 + rh, x0:x0 + rw] + 3 * val) // 4
    speck = rng.integers(-10, 11, (Ht, Wt)).astype(np.int64)
    base = np.clip(base + speck, 0, 255)
    # 3x3 box smoothing (integer) so bilinear sampling does not alias
    p = np.pad(base, 1, mode="edge")
    base = (p[:-2, :-2] + p[:-2, 1:-1] + p[:-2, 2:] + p[1:-1, :-2] + p[1:-1, 1:-1] + p[1:-1, 2:]
            + p[2:, :-2] + p[2:, 1:-1] + p[2:, 2:] + 4) // 9
    tint = [_upsample(rng.integers(-24, 25, (5, 7)).astype(np.int64), Ht, Wt) for _ in range(3)]
    tex = np.stack([np.clip(base + t, 0, 255) for t in tint], axis=-1)
    return tex.astype(np.uint8)


def zoom_scale_q(t, rate=1.015, period=50):
    """s = 2/rate**(t mod period) in Q12 (python int)."""
    return int(round((2.0 / (rate ** (t % period))) * (1 << _Q)))


def zoom_frame(tex, W, H, t, rate=1.015, period=50, foe=(0.48, 0.60)):
    """Frame t of the zoom clip as (H, W, 3) uint8.  `tex` may be a numpy array or a torch tensor
    (any device); the result lives where `tex` lives."""
    is_torch = not isinstance(tex, np.ndarray)
    sq = zoom_scale_q(t, rate, period)
    # FOE in texture coordinates, Q12 ; offset = foe_tex*(1 - s/2)
    fxq = int(round(2 * foe[0] * W * (1 << _Q)))
    fyq = int(round(2 * foe[1] * H * (1 << _Q)))
    offx = fxq - ((fxq * sq) >> (_Q + 1))
    offy = fyq - ((fyq * sq) >> (_Q + 1))
    if is_torch:
        import torch
        xs = torch.arange(W, dtype=torch.int64, device=tex.device)
        ys = torch.arange(H, dtype=torch.int64, device=tex.device)
        T = tex
    else:
        xs = np.arange(W, dtype=np.int64)
        ys = np.arange(H, dtype=np.int64)
        T = tex
    out = _bilinear_q(T, xs * sq + offx, ys * sq + offy)
    if is_torch:
        import torch
        return out.to(torch.uint8)
    return out.astype(np.uint8)


def zoom_clip(W, H, n_frames, seed=0, rate=1.015, period=50, foe=(0.48, 0.60), tex=None, t0=0):
    """List of n_frames BGR frames (numpy)."""
    if tex is None:
        tex = make_texture(W, H, seed)
    return [zoom_frame(tex, W, H, t0 + t, rate, period, foe) for t in range(n_frames)]


def foe_px(W, H, foe=(0.48, 0.60)):
    return (foe[0] * W, foe[1] * H)


def adversarial_points(W, H, n, seed=0):
    """Points that exercise the status paths of the tracker: outside the image, on the borders."""
    rng = np.random.default_rng(seed + 1234)
    pts = np.empty((n, 2), np.float32)
    for i in range(n):
        k = i % 4
        if k == 0:      # outside
            pts[i] = (rng.uniform(-60, -1), rng.uniform(0, H))
        elif k == 1:    # right / bottom border
            pts[i] = (W - 1 - rng.uniform(0, 2), H - 1 - rng.uniform(0, 2))
        elif k == 2:    # top-left border, fractional
            pts[i] = (rng.uniform(0, 3), rng.uniform(0, 3))
        else:           # far outside
            pts[i] = (W + rng.uniform(1, 80), H + rng.uniform(1, 80))
    return pts
