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


@pytest.mark.parametrize("name", ["MESS6", "MAZE0"])
def test_host_fill_convex_polygon_matches_cv2_planes(oracle, name):
    """The planner layer's own fillConvexPoly (planners/raster.hpp, product code) redraws the id planes of DrawBImap
    (CDRT.cpp:121-175) exactly as the real cv2.fillConvexPoly drew them (tests/golden/cdrt/<MAP>.npz)."""
    import ctypes as C

    from rtpp_b200 import capi

    L = capi.planners_lib()
    fx = oracle.load_cdrt_fixture(name)
    H, W = fx["H"], fx["W"]
    imap = np.full((H, W), 0xFFFF, np.uint16)
    g = fx["graphs"][0]
    for pi in range(len(g["poly_core"])):
        v = np.ascontiguousarray(g["pv"][g["pv_off"][pi]:g["pv_off"][pi + 1]].astype(np.int32))
        L.pl_fill_convex16(imap.ctypes.data_as(C.c_void_p), H, W, v.ctypes.data_as(C.c_void_p), len(v), pi)
    assert np.array_equal(imap, fx["imap"])
