#!/usr/bin/env python
"""One fiasco_coder() call on synthetic frames, for profiling runs (ncu -k regex:fiasco_tile_kernel ...).

    python tools/coder_once.py nd512            a 512^2 grey still with --prediction (one launch)
    python tools/coder_once.py colour N         N colour frames 720x576, pattern IPPP (N launches)
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from fiasco_b200 import ffi, hostlib  # noqa: E402
import gen_frames  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "nd512"
    L = hostlib.load()
    o = hostlib.cli_options(0)
    with tempfile.TemporaryDirectory() as tmp:
        if what == "nd512":
            names = [os.path.join(tmp, "nd.pgm")]
            gen_frames.write_pnm(names[0], gen_frames.nd_still())
            L.fiasco_c_options_set_prediction(o, 1, 6, 10)
            q = 80.0
        else:
            n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
            names = []
            for i, f in enumerate(gen_frames.colour_video(n, 720, 576)):
                names.append(os.path.join(tmp, "c%02d.ppm" % i))
                gen_frames.write_pnm(names[-1], f)
            L.fiasco_c_options_set_frame_pattern(o, b"ippp")
            q = 20.0
        ffi.counters(reset=True)
        t0 = time.perf_counter()
        ok, msg = hostlib.coder(names, os.path.join(tmp, "out.fco"), q, options=o)
        dt = time.perf_counter() - t0
        c = ffi.counters()
        print("%s: ok=%s %.3f s, kernels %.1f ms in %d launches, %d bytes %s" % (
            what, ok, dt, c["kernel_ms"], c["launches"], os.path.getsize(os.path.join(tmp, "out.fco")) if ok else 0, msg))


if __name__ == "__main__":
    main()
