"""Seeded input generators shared by the parity tests."""
import numpy as np


def edge_mix(rng, n, W, H, short_px=80.0, long_frac=0.125):
    """n edges as float32 (ps.x, ps.y, pe.x, pe.y): mostly short k-NN-like edges plus a share of
    uniform long ones; coordinates in [0,W) x [0,H) like rand()/float(RAND_MAX)*shape with u<1."""
    a = rng.random((n, 2), dtype=np.float32) * np.array([W, H], np.float32)
    d = (rng.random((n, 2), dtype=np.float32) - np.float32(0.5)) * np.float32(2 * short_px)
    b = a + d
    nl = int(n * long_frac)
    b[:nl] = rng.random((nl, 2), dtype=np.float32) * np.array([W, H], np.float32)
    hi = np.array([np.nextafter(np.float32(W), np.float32(0)), np.nextafter(np.float32(H), np.float32(0))], np.float32)
    b = np.clip(b, 0, hi).astype(np.float32)
    a = np.minimum(a, hi)
    return np.ascontiguousarray(np.concatenate([a, b], 1), dtype=np.float32)


def dyn_plane(shape, rng, n_blobs=6, value=127):
    """A dynamic-obstacle byte plane (0/127) with a few discs and boxes."""
    H, W = shape
    D = np.zeros(shape, np.uint8)
    yy, xx = np.ogrid[:H, :W]
    for i in range(n_blobs):
        cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
        r = int(rng.integers(20, 120))
        if i % 2:
            D[max(0, cy - r): cy + r, max(0, cx - r): cx + r] = value
        else:
            y0, y1, x0, x1 = max(0, cy - r), min(H, cy + r + 1), max(0, cx - r), min(W, cx + r + 1)
            sub = (yy[y0:y1] - cy) ** 2 + (xx[:, x0:x1] - cx) ** 2 <= r * r
            D[y0:y1, x0:x1][sub] = value
    return D
