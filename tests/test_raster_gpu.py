"""D1 parity: the GPU rasteriser (rtpp_map_draw_dynamic) against planes drawn by the real OpenCV
(golden file from python cv2 4.13), the oracle rasteriser, and a full scripted obstacle set."""
import os

import numpy as np
import pytest

from helpers import TESTD_OUTLINES, raster_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raster_cv2.npz")


def test_gpu_raster_matches_opencv_golden(ctx):
    g = np.load(GOLD)
    for i, (kind, H, W, arg) in enumerate(raster_cases()):
        want = np.unpackbits(g[f"c{i}"])[: H * W].reshape(H, W)
        ctx.map_set_plane(1, np.zeros((H, W), np.uint8))
        if kind == "circle":
            ctx.map_draw_dynamic(circles=[arg])
        else:
            ctx.map_draw_dynamic(polys=[arg])
        got = ctx.map_get_plane(2)
        assert np.array_equal(got != 0, want != 0), f"case {i} {kind}"


def test_gpu_raster_obstacle_set_and_merge(ctx, oracle):
    """GAME0 with its scripted obstacle set (RTRRT/test/testD.cpp:57-60): D plane equals the oracle
    rasteriser's, S|D is refreshed, redrawing after a move clears the previous plane."""
    S = oracle.load_map("GAME0")
    ctx.map_set_plane(1, S)
    for shift in (0, 37):
        circles = [(370 + shift, 388, 57), (2370, 630 + shift, 114)]
        polys = []
        for (x, y, ang, outline) in ((1364 + shift, 1364, 0.78539816, TESTD_OUTLINES[3]), (1340, 2520 - shift, 0.0, TESTD_OUTLINES[4])):
            c, s = np.float32(np.cos(np.float32(ang))), np.float32(np.sin(np.float32(ang)))
            o = np.array(outline, np.float32)
            polys.append(np.stack([(c * o[:, 0] - s * o[:, 1] + np.float32(x)).astype(np.int32),
                                   (s * o[:, 0] + c * o[:, 1] + np.float32(y)).astype(np.int32)], 1))
        want = np.zeros(S.shape, np.uint8)
        for cx, cy, r in circles:
            oracle.raster_circle(want, cx, cy, r)
        for p in polys:
            oracle.raster_fillpoly(want, p)
        ctx.map_draw_dynamic(circles=circles, polys=polys)
        assert np.array_equal(ctx.map_get_plane(2), (want != 0).astype(np.uint8))
        assert np.array_equal(ctx.map_get_plane(3), ((want != 0) | (S != 0)).astype(np.uint8))
    ctx.map_draw_dynamic()  # no obstacles: cv::Mat::zeros
    assert not ctx.map_get_plane(2).any()


def test_gpu_raster_random_shapes_vs_oracle(ctx, oracle):
    rng = np.random.default_rng(12)
    H, W = 700, 900
    ctx.map_set_plane(1, np.zeros((H, W), np.uint8))
    for _ in range(30):
        circles = [(int(rng.integers(-40, W + 40)), int(rng.integers(-40, H + 40)), int(rng.integers(0, 130))) for _ in range(5)]
        polys = []
        for _ in range(4):
            n = int(rng.integers(3, 14))
            px, py = int(rng.integers(140, W - 140)), int(rng.integers(140, H - 140))
            ang = np.sort(rng.random(n) * 2 * np.pi)
            rad = rng.random(n) * 130 + 2
            polys.append(np.stack([px + rad * np.cos(ang), py + rad * np.sin(ang)], 1).astype(np.int32))
        want = np.zeros((H, W), np.uint8)
        for cx, cy, r in circles:
            oracle.raster_circle(want, cx, cy, r)
        for p in polys:
            oracle.raster_fillpoly(want, p)
        ctx.map_draw_dynamic(circles=circles, polys=polys)
        assert np.array_equal(ctx.map_get_plane(2), (want != 0).astype(np.uint8))
