#!/usr/bin/env python
"""Tile-split timing on one GPU without torch: config 3 (c2048 as 64 x 256^2 colour tiles, q=30) and config 4
(g4096 as 64 x 512^2 grey tiles, q=20) through the C ABI, kernel time and md5 of a digest over all automata.
usage: python tools/tiles_quick.py [c4|c3|c2t] [repeats]"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fiasco_b200 as F  # noqa: E402
from fiasco_b200 import ffi  # noqa: E402
import gen_frames  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c4"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if which == "c4":
        img, tile, q, bands = gen_frames.frame("g4096"), 512, 20.0, 1
    elif which == "c3":
        img, tile, q, bands = gen_frames.frame("c2048"), 256, 30.0, 3
    else:
        img, tile, q, bands = gen_frames.frame("g1024"), 256, 20.0, 1
    crops = gen_frames.crops(img, tile)
    p = ffi.make_params(tile, tile, bands, q, 0)
    enc = F.TileEncoder(p, len(crops))
    if bands == 1:
        planes = [ffi.pixels_from_grey(c).reshape(-1) for c in crops]
    else:
        import oracle_lib as O
        planes = [O.planes_of(c) for c in crops]
    for r in range(reps):
        t0 = time.perf_counter()
        wfas, _ = enc.encode(planes)
        dt = time.perf_counter() - t0
        st = enc.stats()
        dig = hashlib.md5()
        for w in wfas:
            dig.update("\n".join(F.wfa_lines(w)).encode())
        print(json.dumps({"workload": which, "tiles": len(crops), "cluster": os.environ.get("FB200_CLUSTER", "auto"),
                          "kernel_ms": st["kernel_ms"], "wall_ms": dt * 1e3, "mp_calls": st["mp_calls"],
                          "automata_md5": dig.hexdigest()}), flush=True)
    enc.close()


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    main()
