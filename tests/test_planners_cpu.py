"""CPU: host-side pieces of the C++ planner layer that need no GPU: the product rasteriser
(rtpp_b200/planners/raster.hpp) against planes drawn by the real OpenCV."""
import os

import numpy as np
import pytest

from helpers import raster_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raster_cv2.npz")


def test_host_rasteriser_matches_opencv_golden():
    from rtpp_b200 import capi

    g = np.load(GOLD)
    for i, (kind, H, W, arg) in enumerate(raster_cases()):
        want = np.unpackbits(g[f"c{i}"])[: H * W].reshape(H, W)
        img = np.zeros((H, W), np.uint8)
        if kind == "circle":
            capi.host_raster_circle(img, arg[0], arg[1], arg[2])
        else:
            capi.host_raster_fillpoly(img, np.asarray(arg, np.int32))
        assert np.array_equal(img != 0, want != 0), f"case {i} {kind}"


def test_host_rasteriser_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    from rtpp_b200 import capi

    rng = np.random.default_rng(4)
    for _ in range(200):
        H, W = int(rng.integers(60, 400)), int(rng.integers(60, 400))
        a, b = np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8)
        cx, cy, r = int(rng.integers(-30, W + 30)), int(rng.integers(-30, H + 30)), int(rng.integers(0, 150))
        cv2.circle(a, (cx, cy), r, 127, cv2.FILLED)
        capi.host_raster_circle(b, cx, cy, r)
        n = int(rng.integers(3, 12))
        px, py = int(rng.integers(25, W - 25)), int(rng.integers(25, H - 25))
        ang = np.sort(rng.random(n) * 2 * np.pi)
        rad = rng.random(n) * (min(px, py, W - px, H - py) - 3) + 2
        pts = np.stack([px + rad * np.cos(ang), py + rad * np.sin(ang)], 1).astype(np.int32)
        cv2.fillPoly(a, [pts], 127)
        capi.host_raster_fillpoly(b, pts)
        assert np.array_equal(a, b)
