#!/usr/bin/env python
"""Experiment helper (runs on the GPU box): device-resident throughput of the tile kernel for
one thread-block shape, chosen through the environment (FB200_NT, FB200_BIG) by the caller.

    FB200_NT=64 FB200_BIG=2 python tools/sweep_nt.py [tiles_per_sm] [distinct_frames]

Encodes B = SMs x tiles_per_sm frames (distinct_frames seeded frames, repeated), checks frame 0
against the reference automaton of the g1024 golden case, prints Mpixels/s and lap shares.
"""
import gzip
import hashlib
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
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

NAMES = ["ctrl", "pix", "dots", "upsweep", "enter", "mp_pro", "mp_p1", "mp_waves", "mp_commit", "mp_ortho", "ar_epi",
         "ap_img", "ap_direct", "ap_staged", "decide"]


def main():
    if os.environ.get("FB200_LIB"):              # experiment builds (tools/build_variant.sh)
        ffi.lib_path = lambda: os.path.join(ROOT, os.environ["FB200_LIB"])
    per_sm = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    distinct = int(sys.argv[2]) if len(sys.argv) > 2 else 74
    p = ffi.make_params(1024, 1024, 1, 20.0, 0)
    probe = F.TileEncoder(p, 1)
    res = probe.resident_tiles()
    probe.close()
    sms = 148
    B = sms * per_sm if per_sm else res
    B = min(B, 8192)
    planes = [ffi.pixels_from_grey(gen_frames.chan(1024, 1024, 3 + k)).reshape(-1) for k in range(distinct)]
    planes = [planes[i % distinct] for i in range(B)]
    enc = F.TileEncoder(p, B)
    enc.upload(planes)
    enc.launch(B)
    enc.sync()
    t = []
    for _ in range(2):
        t0 = time.perf_counter()
        enc.launch(B)
        enc.sync()
        t.append(time.perf_counter() - t0)
    w = enc.download(B)[0]
    st = enc.stats()
    enc.close()
    lines = F.wfa_lines(w)
    ok = lines == O.golden_wfa_lines("g1024_q20_z0")
    tot = float(sum(st["lap"])) or 1.0
    print("%s NT=%s BIG=%s resident=%d B=%d (%.1f/SM): %.1f Mpx/s (launch %.1f ms) parity(frame0)=%s states0=%d md5=%s" % (
        os.environ.get("FB200_LIB", "product"), os.environ.get("FB200_NT", "-"), os.environ.get("FB200_BIG", "-"), res, B, B / sms,
        B * 1.048576 / min(t), 1e3 * min(t), ok, w["states"], hashlib.md5("\n".join(lines).encode()).hexdigest()[:8]),
        flush=True)
    print("   work/frame: mp_calls %.0f steps %.0f pass2 evaluations %.0f (%.1f per step, incl. failing last steps: %.1f per find)"
          % (st["mp_calls"] / B, st["mp_steps"] / B, st["pass2"] / B, st["pass2"] / max(1, st["mp_steps"]),
             st["pass2"] / max(1, st["mp_steps"] + st["mp_calls"])), flush=True)
    print("   laps: " + " ".join("%s=%.1f" % (n, 100 * v / tot) for n, v in zip(NAMES, st["lap"])), flush=True)


if __name__ == "__main__":
    main()
