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


def synthetic_diffusion(W, G=100, dim=10, seed=5):
    """A synthetic diffusion-map table like AMRRT/diffusion/*: G x G grid cells (some missing),
    `dim` floats per cell; ratio = G / shapeX narrowed to float (AMRRT/test/test.cpp:56)."""
    rng = np.random.default_rng(seed)
    cells = np.array([(i, j) for i in range(G) for j in range(G) if (i * 7 + j * 3) % 11 != 0], np.int32)
    mat = rng.standard_normal((len(cells), dim)).astype(np.float32)
    return cells, mat, np.float32(G / W)


def golden_cases():
    """The scenario definitions shared by tools/make_golden.py (reference run) and the tests.
    Start / goal points come from the reference's scripted scenarios (RTRRT/test/testD.cpp:36-47)."""
    return {
        "rtrrt": [
            dict(name="mess6", map="MESS6", agent=(720.0, 720.0), goal=(70.0, 70.0), seed=11, n_rand=3_000_000, dt=5000, ticks=8,
                 dobs_ticks=(2, 5), move_ticks=(3,)),
            dict(name="game0", map="GAME0", agent=(350.0, 2350.0), goal=(1400.0, 2050.0), seed=3, n_rand=3_000_000, dt=2000, ticks=6,
                 dobs_ticks=(2,), move_ticks=()),
            # the two largest bundled maps (2339 x 2338, 3318 x 3319: the longest edges and, on MAZE0, rows != columns)
            dict(name="floor5", map="FLOOR5", agent=(1000.0, 1200.0), goal=(1920.0, 600.0), seed=21, n_rand=3_000_000, dt=5000, ticks=6,
                 dobs_ticks=(1, 4), move_ticks=(2,)),
            dict(name="maze0", map="MAZE0", agent=(1500.0, 1850.0), goal=(3100.0, 3100.0), seed=22, n_rand=3_000_000, dt=5000, ticks=6,
                 dobs_ticks=(2,), move_ticks=(3,)),
        ],
        "rtfmt": [
            dict(name="mess6", map="MESS6", agent=(720.0, 720.0), goal=(70.0, 70.0), seed=7, n_rand=100_000, samples=2000, rounds=8,
                 dobs_rounds=(0, 4), move_rounds=(5,)),
            dict(name="floor5", map="FLOOR5", agent=(1000.0, 1200.0), goal=(1460.0, 2100.0), seed=23, n_rand=100_000, samples=2000, rounds=6,
                 dobs_rounds=(1, 3), move_rounds=(4,)),
            dict(name="maze0", map="MAZE0", agent=(1500.0, 1850.0), goal=(200.0, 3100.0), seed=24, n_rand=100_000, samples=2000, rounds=6,
                 dobs_rounds=(2,), move_rounds=(3,)),
        ],
        "amrrt": [
            dict(name="mess6", map="MESS6", agent=(720.0, 720.0), goal=(1240.0, 1240.0), seed=9, n_rand=3_000_000, dt=500000, ticks=25,
                 dobs_ticks=(0, 12), diffusion=synthetic_diffusion),
            dict(name="floor5", map="FLOOR5", agent=(1000.0, 1200.0), goal=(1750.0, 900.0), seed=25, n_rand=3_000_000, dt=500000, ticks=16,
                 dobs_ticks=(0, 8), diffusion=synthetic_diffusion),
            dict(name="maze0", map="MAZE0", agent=(1500.0, 1850.0), goal=(3100.0, 200.0), seed=26, n_rand=3_000_000, dt=500000, ticks=16,
                 dobs_ticks=(3,), diffusion=synthetic_diffusion),
        ],
    }


# outlines of the reference's scripted dynamic obstacles (RTRRT/test/testD.cpp:57-69)
TESTD_OUTLINES = [
    [(-65, 65), (65, 65), (65, -65), (-65, -65)],
    [(-54, -28), (-28, -28), (-28, -54), (28, -54), (28, -28), (54, -28), (54, 28), (28, 28), (28, 54), (-28, 54), (-28, 28), (-54, 28)],
    [(0, -81), (-81, 81), (81, 81)],
    [(0, -57), (57, 0), (57, 57), (-57, 57), (-57, 0)],
    [(-57, -57), (57, -57), (57, 57), (-57, 57)],
    [(-52, 0), (-31, -49), (31, -49), (52, 0), (31, 49), (-31, 49)],
]


def raster_cases(seed=3):
    """(kind, H, W, arg): filled circles (also clipped by the border) and in-image polygons."""
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(40):
        H, W = int(rng.integers(60, 300)), int(rng.integers(60, 300))
        cases.append(("circle", H, W, (int(rng.integers(-30, W + 30)), int(rng.integers(-30, H + 30)), int(rng.integers(0, 120)))))
    for _ in range(40):
        H, W = int(rng.integers(60, 300)), int(rng.integers(60, 300))
        n = int(rng.integers(3, 13))
        cases.append(("poly", H, W, np.stack([rng.integers(2, W - 2, n), rng.integers(2, H - 2, n)], 1).astype(np.int32).tolist()))
    for k, outline in enumerate(TESTD_OUTLINES):
        sh = np.array(outline, np.float32)
        for a in (0.0, 0.78539816, 1.1, 2.5, 4.0):
            c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
            pts = np.stack([(c * sh[:, 0] - s * sh[:, 1] + np.float32(150)).astype(np.int32),
                            (s * sh[:, 0] + c * sh[:, 1] + np.float32(150)).astype(np.int32)], 1)
            cases.append(("poly", 300, 300, pts.tolist()))
    return cases
