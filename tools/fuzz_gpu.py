#!/usr/bin/env python
"""Randomised parity campaign (runs on the GPU box): seeded random crops of the synthetic frames at
random sizes, qualities and optimisation levels, grey and colour, device vs oracle, bit for bit
(automaton lines, which include the weights' bit patterns).

    python tools/fuzz_gpu.py [cases] [seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fiasco_b200 as F  # noqa: E402
from fiasco_b200 import ffi  # noqa: E402
import oracle_lib as O  # noqa: E402
import gen_frames  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    grey = gen_frames.frame("g1024")
    col = gen_frames.frame("c256")
    bad = 0
    t0 = time.time()
    for case in range(n):
        colour = rng.random() < 0.25
        src = col if colour else grey
        H, W = src.shape[:2]
        w = int(rng.integers(17, min(W, 320) // 2 + 1)) * 2          # even, >= 34 (lib/image.c:194)
        h = int(rng.integers(17, min(H, 320) // 2 + 1)) * 2
        x = int(rng.integers(0, W - w + 1))
        y = int(rng.integers(0, H - h + 1))
        q = float(rng.choice([8, 12, 20, 30, 45, 60, 90]))
        z = int(rng.choice([0, 0, 0, 1, 2]))
        img = np.ascontiguousarray(src[y:y + h, x:x + w])
        if rng.random() < 0.1:
            img = np.full_like(img, int(rng.integers(0, 256)))       # flat block
        ow = O.encode(img, quality=q, optimize=z)
        p = ffi.make_params(w, h, 3 if colour else 1, q, z)
        enc = F.TileEncoder(p, 1)
        try:
            gw = enc.encode(O.planes_of(img))[0][0]
        finally:
            enc.close()
        lvl = ow["level"]
        a = O.mask_virtual(F.wfa_lines(gw), lvl) if colour else F.wfa_lines(gw)
        b = O.mask_virtual(O.wfa_lines(ow), lvl) if colour else O.wfa_lines(ow)
        ok = a == b and gw["states"] == ow["states"]
        if not ok:
            bad += 1
            print("MISMATCH case %d: %s %dx%d@%d,%d q=%g z=%d states gpu %d oracle %d"
                  % (case, "colour" if colour else "grey", w, h, x, y, q, z, gw["states"], ow["states"]), flush=True)
    print("fuzz: %d cases, %d mismatches, %.1f s" % (n, bad, time.time() - t0), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
