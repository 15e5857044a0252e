#!/usr/bin/env python
"""Convert the reference's bundled PNG maps into obstacle bit-planes (test fixtures).

Runs only in the build container (needs /root/reference and cv2).  Follows the exact call
sequence of RTRRTplanner::InitializeInvironment (reference RTRRT/lib/RTrrt/RTRRT.cpp:57-80):
    imread(GRAYSCALE) -> threshold(th=128, 255, THRESH_BINARY) -> bitwise_not
so Obsmap[y][x] != 0  <=>  pixel is an obstacle.  The plane is stored as np.packbits(obs != 0,
axis=1, bitorder='little') in a compressed .npz with the shape and a SHA-256 of the byte plane.
"""
import hashlib, json, os, sys
import cv2
import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "maps")
MAPS = {name: f"{REF}/RTRRT/map/{name}.png" for name in
        ["FLOOR2", "FLOOR5", "GAME0", "GAME6", "MAZE0", "MAZE5", "MESS6", "Trap", "dogmap"]}
MAPS["dilated"] = f"{REF}/AMRRT/diffusion/dilated.png"


def obstacle_plane(path, th=128):
    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    assert img is not None, path
    _, free = cv2.threshold(img, th, 255, cv2.THRESH_BINARY)
    assert free.dtype == np.uint8
    return cv2.bitwise_not(free)


def main():
    os.makedirs(OUT, exist_ok=True)
    index = {}
    for name, path in MAPS.items():
        obs = obstacle_plane(path)
        h, w = obs.shape
        bits = np.packbits(obs != 0, axis=1, bitorder="little")
        sha = hashlib.sha256(obs.tobytes()).hexdigest()
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), bits=bits, shape=np.array([h, w], np.int32))
        index[name] = {"rows": int(h), "cols": int(w), "obstacle_fraction": float((obs != 0).mean()),
                       "sha256_byteplane": sha, "source": path.replace(REF + "/", "")}
        print(name, h, w, index[name]["obstacle_fraction"], sha[:12])
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())
