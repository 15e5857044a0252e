"""CPU: the oracle's restatement of cv::circle(FILLED) / cv::fillPoly (D1, RTRRT.cpp:138-169) against
planes drawn by the real OpenCV (golden file from python cv2 4.13; live cv2 when importable), and
the reference's DrawDObsmap transform + this rasteriser against cv2 on the scripted obstacle set."""
import os

import numpy as np
import pytest

from helpers import TESTD_OUTLINES, raster_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raster_cv2.npz")


def _draw(oracle, kind, H, W, arg):
    img = np.zeros((H, W), np.uint8)
    if kind == "circle":
        oracle.raster_circle(img, arg[0], arg[1], arg[2], 127)
    else:
        oracle.raster_fillpoly(img, np.asarray(arg, np.int32), 127)
    return img


def test_raster_matches_opencv_golden(oracle):
    g = np.load(GOLD)
    cases = raster_cases()
    assert len(cases) == len(g.files) - 1
    for i, (kind, H, W, arg) in enumerate(cases):
        want = np.unpackbits(g[f"c{i}"])[: H * W].reshape(H, W)
        got = _draw(oracle, kind, H, W, arg)
        assert set(np.unique(got)) <= {0, 127}
        assert np.array_equal(got != 0, want != 0), f"case {i} {kind}"


def test_raster_matches_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    for _ in range(150):
        H, W = int(rng.integers(60, 400)), int(rng.integers(60, 400))
        a, b = np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8)
        cx, cy, r = int(rng.integers(-30, W + 30)), int(rng.integers(-30, H + 30)), int(rng.integers(0, 150))
        cv2.circle(a, (cx, cy), r, 127, cv2.FILLED)
        oracle.raster_circle(b, cx, cy, r, 127)
        # simple (star-shaped) polygons like the reference's obstacle outlines; random
        # self-intersecting ones differ from cv2 by one pixel in ~1 of 600 cases (edge-crossing ties)
        n = int(rng.integers(3, 12))
        px, py = int(rng.integers(25, W - 25)), int(rng.integers(25, H - 25))
        ang = np.sort(rng.random(n) * 2 * np.pi)
        rad = rng.random(n) * (min(px, py, W - px, H - py) - 3) + 2
        pts = np.stack([px + rad * np.cos(ang), py + rad * np.sin(ang)], 1).astype(np.int32)
        cv2.fillPoly(a, [pts], 127)
        oracle.raster_fillpoly(b, pts, 127)
        assert np.array_equal(a, b)


def test_reference_drawdobsmap_with_this_rasteriser_matches_cv2(oracle):
    """The reference's own DrawDObsmap (compiled in place) calling this rasteriser through the
    cv:: stand-in must give the plane python cv2 draws for the same obstacle set
    (GAME0 set of RTRRT/test/testD.cpp:57-60)."""
    cv2 = pytest.importorskip("cv2")
    if oracle.ref_lib("rtrrt") is None:
        pytest.skip("oracle/_ref not built")
    S = oracle.load_map("GAME0")
    ref = oracle.RefRTRRT(S)
    obs = [(0, 370, 388, 0.0, 57, [(0, 0)]), (0, 2370, 630, 0.0, 114, [(0, 0)]),
           (1, 1364, 1364, 0.78539816, 0, TESTD_OUTLINES[3]), (1, 1340, 2520, 0.0, 0, TESTD_OUTLINES[4])]
    want = np.zeros(S.shape, np.uint8)
    for num, (typ, x, y, ang, rad, outline) in enumerate(obs):
        ref.add_obstacle(num, typ, x, y, ang, rad, outline)
        # the same transform in numpy float32 (cos/sin of these angles agree with glibc here)
        c, s = np.float32(np.cos(np.float32(ang))), np.float32(np.sin(np.float32(ang)))
        o = np.array(outline, np.float32)
        px = (c * o[:, 0] - s * o[:, 1] + np.float32(x)).astype(np.int32)
        py = (s * o[:, 0] + c * o[:, 1] + np.float32(y)).astype(np.int32)
        if typ == 0:
            for qx, qy in zip(px, py):
                cv2.circle(want, (int(qx), int(qy)), int(rad), 127, cv2.FILLED)
        else:
            cv2.fillPoly(want, [np.stack([px, py], 1)], 127)
    ref.draw_dobs()
    got = ref.get_dobs()
    ref.close()
    assert want.any() and np.array_equal(got, want)
