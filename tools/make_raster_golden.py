#!/usr/bin/env python
"""Golden planes for the D1 rasteriser from the REAL OpenCV (python cv2): cv2.circle(FILLED) and
cv2.fillPoly on seeded in-image shapes, incl. the reference's scripted obstacle outlines
(RTRRT/test/testD.cpp:57-69) at several rotations.  Output: tests/golden/raster_cv2.npz."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import raster_cases  # noqa: E402


def main():
    out = {"cv2_version": cv2.__version__}
    for i, (kind, H, W, arg) in enumerate(raster_cases()):
        img = np.zeros((H, W), np.uint8)
        if kind == "circle":
            cv2.circle(img, (arg[0], arg[1]), arg[2], 127, cv2.FILLED)
        else:
            cv2.fillPoly(img, [np.asarray(arg, np.int32)], 127)
        out[f"c{i}"] = np.packbits(img != 0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "raster_cv2.npz"), **out)
    print(len(out) - 1, "cases, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
