"""Synthetic slide tiles and nucleus polygons. Seeded, numpy only."""
from __future__ import annotations

import numpy as np


def synth_tile(h: int, w: int, seed: int = 1, out=None, band: int = 1024):
    """H&E-like u8 interleaved RGB [h,w,3]: low-frequency pink/purple field + per-pixel noise.
    Row 0 carries a 0..255 ramp in every channel so that every 8-bit value occurs."""
    rng = np.random.default_rng(seed)
    if out is None:
        out = np.empty((h, w, 3), dtype=np.uint8)
    g = 16
    ch, cw = h // g + 2, w // g + 2
    coarse = rng.random((ch, cw), dtype=np.float32)
    coarse2 = rng.random((ch, cw), dtype=np.float32)
    pink = np.array([232, 160, 204], dtype=np.float32)
    purple = np.array([96, 60, 150], dtype=np.float32)
    xs = np.arange(w, dtype=np.float32) / g
    x0 = xs.astype(np.int64)
    fx = (xs - x0)[None, :]
    for r0 in range(0, h, band):
        r1 = min(h, r0 + band)
        ys = np.arange(r0, r1, dtype=np.float32) / g
        y0 = ys.astype(np.int64)
        fy = (ys - y0)[:, None]

        def bil(c):
            a = c[y0][:, x0] * (1 - fx) + c[y0][:, x0 + 1] * fx
            b = c[y0 + 1][:, x0] * (1 - fx) + c[y0 + 1][:, x0 + 1] * fx
            return a * (1 - fy) + b * fy

        mix = bil(coarse)[..., None]
        shade = (0.75 + 0.5 * bil(coarse2))[..., None]
        base = (purple + (pink - purple) * mix) * shade
        noise = rng.integers(-16, 17, size=(r1 - r0, w, 3), dtype=np.int16)
        out[r0:r1] = np.clip(base + noise, 0, 255).astype(np.uint8)
    ramp = (np.arange(w) % 256).astype(np.uint8)
    out[0, :, 0] = ramp
    out[0, :, 1] = ramp[::-1]
    out[0, :, 2] = np.roll(ramp, 85)
    return out


def synth_polygons(n: int, h: int, w: int, seed: int = 1, patch: int = 64, r0_range=(6.0, 26.0),
                   v_range=(12, 48), border_frac: float = 0.01, rough: float = 0.25, harmonics=(2, 3, 5)):
    """Star-shaped closed rings (first vertex repeated, as GeoJSON stores them) in CSR form.
    Returns (poly_xy f32 [sum(V+1), 2], poly_off int64 [n+1])."""
    rng = np.random.default_rng(seed)
    V = rng.integers(v_range[0], v_range[1] + 1, size=n)
    r0 = rng.uniform(r0_range[0], r0_range[1], size=n)
    m = patch / 2 + 1
    cx = rng.uniform(m, max(w - m, m + 1), size=n)
    cy = rng.uniform(m, max(h - m, m + 1), size=n)
    nb = int(round(border_frac * n))
    if nb:
        idx = rng.choice(n, size=nb, replace=False)
        side = rng.integers(0, 4, size=nb)
        cx[idx] = np.where(side == 0, rng.uniform(0, m, nb), np.where(side == 1, rng.uniform(w - m, w, nb), cx[idx]))
        cy[idx] = np.where(side == 2, rng.uniform(0, m, nb), np.where(side == 3, rng.uniform(h - m, h, nb), cy[idx]))
    amp = rng.uniform(0.2, 1.0, size=(n, len(harmonics)))
    amp /= amp.sum(axis=1, keepdims=True)
    ph = rng.uniform(0, 2 * np.pi, size=(n, len(harmonics)))
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(V + 1, out=off[1:])
    total = int(off[-1])
    nuc = np.repeat(np.arange(n), V + 1)
    k = np.arange(total) - off[nuc]
    k = np.where(k == V[nuc], 0, k)                 # closing duplicate
    theta = 2 * np.pi * k / V[nuc]
    noise = np.zeros(total)
    for j, hm in enumerate(harmonics):
        noise += amp[nuc, j] * np.cos(hm * theta + ph[nuc, j])
    r = r0[nuc] * (1 + rough * noise)
    xy = np.empty((total, 2), dtype=np.float32)
    xy[:, 0] = (cx[nuc] + r * np.cos(theta)).astype(np.float32)
    xy[:, 1] = (cy[nuc] + r * np.sin(theta)).astype(np.float32)
    return xy, off


def synth_polygons_pool(n: int, h: int, w: int, seed: int = 1, patch: int = 64, pool: int = 200_000, **kw):
    """Same rings as synth_polygons for the first `pool` nuclei; nucleus i >= pool re-uses the SHAPE of nucleus i % pool at
    a fresh uniformly random centre (features are independent per nucleus, so millions of nuclei do not need millions of
    distinct shapes: 5M rings take ~3 s instead of ~40 s of host time). Returns (poly_xy f32, poly_off int64)."""
    if n <= pool:
        return synth_polygons(n, h, w, seed, patch=patch, **kw)
    bxy, boff = synth_polygons(pool, h, w, seed, patch=patch, **kw)
    lens = np.diff(boff)
    # shape relative to its first vertex-mean (float64), so that it can be moved
    cen = np.add.reduceat(bxy.astype(np.float64), boff[:-1], axis=0) / lens[:, None]
    rel = bxy.astype(np.float64) - np.repeat(cen, lens, axis=0)
    rng = np.random.default_rng(seed + 104729)
    m = patch / 2 + 1
    blocks, offs, base = [bxy], [boff[1:]], int(boff[-1])
    for k in range(pool, n, pool):
        cnt = min(pool, n - k)
        nv = int(boff[cnt])
        c = np.stack([rng.uniform(m, max(w - m, m + 1), size=cnt), rng.uniform(m, max(h - m, m + 1), size=cnt)], axis=1)
        blocks.append((rel[:nv] + np.repeat(c, lens[:cnt], axis=0)).astype(np.float32))
        offs.append(boff[1:cnt + 1] + base)
        base += nv
    return np.concatenate(blocks), np.concatenate([np.zeros(1, np.int64)] + offs)


def rings_of(poly_xy, poly_off):
    return [poly_xy[poly_off[i]:poly_off[i + 1]] for i in range(len(poly_off) - 1)]
