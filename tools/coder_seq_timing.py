#!/usr/bin/env python
"""fiasco_coder() on a --pattern=i sequence of N seeded 1024^2 frames with FIASCO_TIMINGS=1 (phase times on stderr).
usage: coder_seq_timing.py [frames] [repeats]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
from fiasco_b200 import hostlib  # noqa: E402
import gen_frames  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tmp = tempfile.mkdtemp(prefix="fbseq_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
base = [gen_frames.chan(1024, 1024, 3 + k) for k in range(min(n, 37))]
for k in range(n):
    gen_frames.write_pnm(os.path.join(tmp, "f%04d.pgm" % k), base[k % len(base)])
os.environ["FIASCO_TIMINGS"] = "1"
L = hostlib.load()
for r in range(reps):
    o = hostlib.cli_options(0)
    L.fiasco_c_options_set_progress_meter(o, 0)
    L.fiasco_c_options_set_frame_pattern(o, b"i")
    t0 = time.perf_counter()
    ok, msg = hostlib.coder([os.path.join(tmp, "f[0000-%04d].pgm" % (n - 1))], os.path.join(tmp, "seq.fco"), options=o)
    dt = time.perf_counter() - t0
    L.fiasco_c_options_delete(o)
    print("run %d: ok=%s %.3f s -> %.1f Mpx/s" % (r, ok, dt, n * 1.048576 / dt), msg if not ok else "", flush=True)
import shutil
shutil.rmtree(tmp, ignore_errors=True)
