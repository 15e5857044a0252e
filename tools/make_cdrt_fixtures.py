#!/usr/bin/env python
"""CDRT map fixtures: polygon groups, convex decomposition graph and the two 16-bit id planes.

Build-container only (needs /root/reference, cv2 and oracle/_ref/libref_cdrt.so).  Pipeline, all of it
the reference's own code or the real OpenCV:

  1. approx_work.py's call sequence (CDRT2/approx_work.py:43-64) with cv2 4.x (findContours returns
     two values there; the author ran 3.4.9): cvtColor -> threshold(thresh) -> erode(ellipse
     robotpixel/2) -> findContours(RETR_TREE, CHAIN_APPROX_SIMPLE) -> approxPolyDP(robotpixel/3)
     reversed -> even-depth contours with their holes.  This replaces MaptoBInavi's UDP round trip.
  2. BImap::StartCut on every group — the reference's compiled CDRTmapcd.cpp (oracle/_ref).
  3. BIamap / BIimap by the real cv2.fillConvexPoly in DrawBImap's order (CDRT.cpp:121-175); the
     planes the reference's own DrawBImap produces through the oracle rasteriser are compared with
     them (that is the pin of oracle/raster.hpp:fill_convex_poly).

Output: tests/golden/cdrt/<MAP>.npz (graph arrays of every group, planes as compressed uint16).
"""
import ctypes as C
import hashlib
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "cdrt")
# CDRT2/test/testD.cpp:52-55 (InitTaskMapD): {ratio, robotsize, resolution}
MAPS = {"MESS6": (0.1, 0.5, 11), "FLOOR5": (0.1, 0.5, 20), "GAME0": (0.1, 0.5, 23), "MAZE0": (0.1, 1.1, 28)}


def approx_groups(path, ratio, robotsize, thresh=128):
    img = cv2.imread(path)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    _, binary = cv2.threshold(gray, thresh, 255, cv2.THRESH_BINARY)
    robotpixel = int(robotsize / ratio)
    erode_r = int(robotpixel / 2)
    element = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (erode_r, erode_r))
    erosion = cv2.erode(binary, element, iterations=1)
    contours, hierarchy = cv2.findContours(erosion, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    eps = int(robotpixel / 3)
    approx = [cv2.approxPolyDP(c, eps, True)[:, 0, :][::-1] for c in contours]
    hchy = hierarchy[0]
    groups = []
    for i in range(len(hchy)):
        depth, p = 0, i
        while hchy[p][3] != -1:
            p = hchy[p][3]
            depth += 1
        if depth % 2 == 0:
            g = [approx[i].tolist()]
            c = hchy[i][2]
            if c != -1:
                g.append(approx[c].tolist())
                while hchy[c][0] != -1:
                    c = hchy[c][0]
                    g.append(approx[c].tolist())
            groups.append(g)
    return groups, binary.shape


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def main():
    os.makedirs(OUT, exist_ok=True)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_cdrt.so"))
    lib.ref_cdrt_create.restype = C.c_void_p
    index = {}
    for name, (ratio, rsize, res) in MAPS.items():
        groups, (H, W) = approx_groups(f"{REF}/CDRT2/map/{name}.png", ratio, rsize)
        h = C.c_void_p(lib.ref_cdrt_create(W, H, C.c_double(ratio), C.c_double(rsize)))
        out = {"shape": np.array([H, W], np.int32), "n_groups": np.array(len(groups), np.int32), "resolution": np.array(res, np.float64)}
        amap = np.full((H, W), 0xFFFF, np.uint16)
        imap = np.full((H, W), 0xFFFF, np.uint16)
        for gi, g in enumerate(groups):
            pts = np.array([p for poly in g for p in poly], np.float64)
            sizes = np.array([len(poly) for poly in g], np.int32)
            assert lib.ref_cdrt_add_group(h, ptr(pts), ptr(sizes), len(g)) == gi
            s = np.zeros(7, np.int32)
            lib.ref_cdrt_graph_sizes(h, gi, ptr(s))
            ncut, npoly = int(s[0]), int(s[1])
            a = dict(cut_se=np.zeros((ncut, 4)), cut_core=np.zeros((ncut, 2)), cp_off=np.zeros(ncut + 1, np.int32), cp=np.zeros(s[2], np.int32),
                     cc_off=np.zeros(ncut + 1, np.int32), cc=np.zeros(s[3], np.int32), pv_off=np.zeros(npoly + 1, np.int32), pv=np.zeros((s[4], 2)),
                     poly_core=np.zeros((npoly, 2)), pc_off=np.zeros(npoly + 1, np.int32), pc=np.zeros(s[5], np.int32),
                     pp_off=np.zeros(npoly + 1, np.int32), pp=np.zeros(s[6], np.int32))
            lib.ref_cdrt_graph_dump(h, gi, *[ptr(a[k]) for k in ["cut_se", "cut_core", "cp_off", "cp", "cc_off", "cc", "pv_off", "pv", "poly_core",
                                                                "pc_off", "pc", "pp_off", "pp"]])
            for k, v in a.items():
                out[f"g{gi}_{k}"] = v
            out[f"g{gi}_poly_pts"] = pts
            out[f"g{gi}_poly_sizes"] = sizes
            # DrawBImap with the real OpenCV (vertices truncate to int: CDRT.cpp:140-141)
            for pi in range(npoly):
                v = a["pv"][a["pv_off"][pi]:a["pv_off"][pi + 1]].astype(np.int32)
                cv2.fillConvexPoly(amap, v, gi)
                cv2.fillConvexPoly(imap, v, pi)
            print(f"{name} group {gi}: {len(g)} polygons -> {ncut} cut lines, {npoly} free convex polygons")
        ra = np.zeros((H, W), np.uint16)
        ri = np.zeros((H, W), np.uint16)
        lib.ref_cdrt_draw_bimap(h, ptr(ra), ptr(ri))
        same = bool(np.array_equal(ra, amap) and np.array_equal(ri, imap))
        print(f"{name}: reference DrawBImap through oracle fill_convex_poly == cv2.fillConvexPoly planes: {same}",
              f"(differing px: {int((ra != amap).sum())}, {int((ri != imap).sum())})")
        out["amap"], out["imap"] = amap, imap
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
        index[name] = {"rows": int(H), "cols": int(W), "groups": len(groups), "ratio": ratio, "robotsize": rsize, "resolution": res,
                       "sha256_amap": hashlib.sha256(amap.tobytes()).hexdigest(), "sha256_imap": hashlib.sha256(imap.tobytes()).hexdigest(),
                       "oracle_raster_matches_cv2": same, "bytes": os.path.getsize(os.path.join(OUT, f"{name}.npz"))}
        lib.ref_cdrt_destroy(h)
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())
