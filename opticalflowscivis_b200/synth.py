"""Synthetic inputs of the BASELINE.json configs.  numpy on the host; no reference code involved.

cfg 1  2-D textured rectangle 160x224      rectangle2d()    (recipe of Datasets/create_rectangle_2d.py:81-121)
cfg 2  2-D droplet-shaped 160x224, N=64    droplet2d()
cfg 3  3-D textured rectangle 128^3, N=4   rectangle3d()    (Datasets/create_data_3d.py:41-105 scaled x2)
cfg 4  3-D droplet 256^3 uint8 {0,255}     droplet3d_u8()   (README.md:24-25 ; Datasets/read_data.py:116-119)
"""
from __future__ import annotations

import numpy as np


def _textured_box(rng, shape, box, shift):
    nd = len(shape)
    tiles = rng.integers(30, 256, size=tuple(-(-b // 10) for b in box)).astype(np.float32) / 255.0
    tex = tiles
    for ax in range(nd):
        tex = np.repeat(tex, 10, axis=ax)
    tex = tex[tuple(slice(0, b) for b in box)]
    out = []
    for k in (0, 1, 2):                      # img0, ground-truth middle, img1
        vol = np.zeros(shape, np.float32)
        org = [(s - b) // 2 + (k * v) // 2 for s, b, v in zip(shape, box, shift)]
        vol[tuple(slice(o, o + b) for o, b in zip(org, box))] = tex
        out.append(vol)
    return out


def rectangle2d(n=1, h=160, w=224, seed=1234):
    rng = np.random.default_rng(seed)
    vols = [_textured_box(rng, (h, w), (60, 80), rng.integers(-6, 7, size=2) * 2) for _ in range(n)]
    return tuple(np.stack([v[k] for v in vols])[:, None] for k in range(3))


def rectangle3d(n=4, s=128, seed=1234):
    rng = np.random.default_rng(seed)
    sc = s / 64.0
    box = tuple(int(b * sc) for b in (20, 30, 40))
    vols = [_textured_box(rng, (s, s, s), box, rng.integers(-4, 5, size=3) * 2) for _ in range(n)]
    return tuple(np.stack([v[k] for v in vols])[:, None] for k in range(3))


def droplet2d(n=64, h=160, w=224, seed=1234):
    """Disk r in [8,24] falling onto a horizontal film band, box-blurred; pair = disk moved 2..8 px down."""
    out = [[], [], []]
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        rng = np.random.default_rng(seed + i)
        r, cx, cy = rng.uniform(8, 24), rng.uniform(40, w - 40), rng.uniform(30, 70)
        dy = 2 * rng.integers(1, 5)
        film = (yy > h - 30).astype(np.float32)
        for k in range(3):
            img = np.maximum(film, ((xx - cx) ** 2 + (yy - cy - k * dy / 2) ** 2 < r * r).astype(np.float32))
            img = (img + np.roll(img, 1, 0) + np.roll(img, -1, 0) + np.roll(img, 1, 1) + np.roll(img, -1, 1)) / 5.0
            out[k].append(img)
    return tuple(np.stack(o)[:, None].astype(np.float32) for o in out)


def droplet3d_u8(n=1, s=256, seed=1234):
    """uint8 {0,255} volumes: sphere (r = 40 at s = 256) above a film slab (thickness 24); second volume has the
    sphere centre 6 voxels closer to the film.  Returns (vol0, vol_mid, vol1) each (n,1,s,s,s) uint8."""
    sc = s / 256.0
    out = [[], [], []]
    z = np.arange(s, dtype=np.float32)[:, None, None]
    y = np.arange(s, dtype=np.float32)[None, :, None]
    x = np.arange(s, dtype=np.float32)[None, None, :]
    for i in range(n):
        rng = np.random.default_rng(seed + i)
        r = 40.0 * sc * rng.uniform(0.8, 1.2)
        c = np.array([s * 0.55, s * 0.5, s * 0.5]) + rng.uniform(-10, 10, 3) * sc
        film = z < 24 * sc
        for k in range(3):
            cz = c[0] - k * 3.0 * sc
            ball = (z - cz) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2 < r * r
            out[k].append(np.where(film | ball, np.uint8(255), np.uint8(0)))
    return tuple(np.stack(o)[:, None] for o in out)
