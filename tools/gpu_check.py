#!/usr/bin/env python
"""Debug helper (runs on the GPU box): encode seeded frames on the device and on the oracle,
print the first divergence of the per-range traces / automata.
usage: python tools/gpu_check.py [case ...]   case = name:WxH@x0,y0 from a seeded frame
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fiasco_b200 as F  # noqa: E402
from fiasco_b200 import ffi  # noqa: E402
import oracle_lib as O  # noqa: E402
import gen_frames  # noqa: E402


def trace_line(i, r):
    def fb(x):
        return int(np.float32(x).view(np.uint32))
    s = "lc %d %d %d %d %d %d %d %d %08x %08x %08x" % (i, r.level, r.image, r.address, r.x, r.y, r.y_state, r.states,
                                                      fb(r.max_costs), fb(r.price), fb(r.costs))
    if r.n_edges >= 0 and r.into[0] >= 0:
        s += " %08x %08x %08x :" % (fb(r.err), fb(r.matrix_bits), fb(r.weights_bits))
        for e in range(r.n_edges):
            s += " %d:%08x" % (r.into[e], fb(r.weight[e]))
    return s


def run_case(img, quality=20.0, optimize=0, label="", cap=0):
    h, w = img.shape[:2]
    t0 = time.time()
    ow = O.encode(img, quality=quality, optimize=optimize, want_trace=True)
    t1 = time.time()
    p = ffi.make_params(w, h, 3 if img.ndim == 3 else 1, quality, optimize, cap)
    enc = F.TileEncoder(p, 1)
    gw, tr = enc.encode(O.planes_of(img), trace_cap=200000)
    st = enc.stats()
    enc.close()
    gw = gw[0]
    olc = O.lc_lines(ow["trace"])
    glc = [trace_line(i, r) for i, r in enumerate(tr)]
    ok = True
    for i, (a, b) in enumerate(zip(olc, glc)):
        if a != b:
            print("  TRACE DIVERGES at lc %d:\n    oracle %s\n    gpu    %s" % (i, a, b))
            for j in range(max(0, i - 2), i):
                print("    prev   %s" % olc[j])
            ok = False
            break
    if ok and len(olc) != len(glc):
        print("  trace length differs: oracle %d gpu %d" % (len(olc), len(glc)))
        ok = False
    ol, gl = O.wfa_lines(ow), F.wfa_lines(gw)
    if ol != gl:
        ok = False
        for i, (a, b) in enumerate(zip(ol, gl)):
            if a != b:
                print("  WFA DIVERGES at line %d:\n    oracle %s\n    gpu    %s" % (i, a, b))
                break
        print("  wfa lines: oracle %d gpu %d; states oracle %d gpu %d" % (len(ol), len(gl), ow["states"], gw["states"]))
    print("%s %dx%d q=%g z=%d: %s  states %d  oracle %.3fs  gpu kernel %.3f ms (h2d %.2f d2h %.2f ms) mp %d steps %d"
          % (label, w, h, quality, optimize, "OK" if ok else "MISMATCH", gw["states"], t1 - t0, st["kernel_ms"],
             st["h2d_ms"], st["d2h_ms"], st["mp_calls"], st["mp_steps"]) + " pass2 %d" % st.get("pass2", -1), flush=True)
    names = ["ctrl", "pix", "dots", "upsweep", "enter", "mp_pro", "mp_p1", "mp_waves", "mp_commit", "mp_ortho", "ar_epi",
             "ap_img", "ap_direct", "ap_staged", "decide", "cluster"]
    tot = float(sum(st["lap"])) or 1.0
    print("   laps: " + " ".join("%s=%.1f%%" % (n, 100 * v / tot) for n, v in zip(names, st["lap"])) +
          "  total %.1f Mcyc" % (tot / 1e6), flush=True)
    return ok


def main():
    print(F.load().fb200_version().decode(), "devices:", F.device_count(), flush=True)
    g1024 = gen_frames.frame("g1024")
    cases = [
        ("tile64", g1024[0:64, 0:64], 20, 0),
        ("tile64b", g1024[512:576, 256:320], 20, 0),
        ("tile128", g1024[128:256, 640:768], 20, 0),
        ("rag200x136", g1024[0:136, 0:200], 20, 0),
        ("g256", gen_frames.frame("g256"), 20, 0),
        ("g256q60", gen_frames.frame("g256"), 60, 0),
        ("g512", gen_frames.frame("g512"), 20, 0),
    ]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        cases = [("g1024", g1024, 20, 0)]
    if len(sys.argv) > 1 and sys.argv[1] == "colour":
        c256 = gen_frames.frame("c256")
        cases = [("c64", c256[0:64, 0:64], 20, 0), ("c128", c256[64:192, 128:256], 20, 0), ("c256", c256, 20, 0)]
    if len(sys.argv) > 1 and sys.argv[1] == "z":
        cases = [("g256z1", gen_frames.frame("g256"), 20, 1), ("g256z2", gen_frames.frame("g256"), 20, 2)]
    allok = True
    for label, img, q, z in cases:
        try:
            allok &= run_case(np.ascontiguousarray(img), q, z, label)
        except Exception as e:  # noqa: BLE001
            print(label, "EXCEPTION", repr(e), flush=True)
            allok = False
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
