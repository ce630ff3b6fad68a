#!/usr/bin/env python
"""BASELINE config 5 on ONE GPU without torch (a fresh box pays up to a minute for the first `import torch`):
30 frames 720x576 grey, q=20, pattern IPPP -> 8 groups of pictures, the I frames in one launch, then three
steps of 8 P frames each (one thread block per group); stream written by the host writer.  Prints one JSON
line per run (wall clock around encode + write, kernel milliseconds, md5 of the stream).

    python tools/video_quick.py [runs] [expected md5]
"""
import hashlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from fiasco_b200 import ffi, hostlib, video  # noqa: E402
import gen_frames  # noqa: E402


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    want = sys.argv[2] if len(sys.argv) > 2 else None
    n, w, h, pattern, q = int(os.environ.get("FBQ_FRAMES", "30")), 720, 576, "ippp", 20.0
    planes = [ffi.pixels_from_grey(f) for f in gen_frames.video(n, w, h)]
    p = ffi.make_params(w, h, 1, q, 0)
    # FBQ_REPLICAS = R: R copies of the sequence as independent sequences (R x the groups in flight)
    reps = int(os.environ.get("FBQ_REPLICAS", "1"))
    gl = [(s + k * n, e + k * n) for k in range(reps) for s, e in video.groups(n, pattern)]
    for r in range(runs):
        t0 = time.perf_counter()
        done, kernel_ms = video.encode_groups({f: planes[f % n] for f in range(n * reps)}, gl, p)
        t1 = time.perf_counter()
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "v.fco")
            hostlib.write_video_stream(out, p, [done[f] for f in range(n)])
            data = open(out, "rb").read()
        t2 = time.perf_counter()
        md5 = hashlib.md5(data).hexdigest()
        print(json.dumps({"workload": "%d frames %dx%d grey q=%g pattern %s, %d groups, 1 GPU" % (n, w, h, q, pattern, len(gl)),
                          "run": r, "encode_s": t1 - t0, "write_s": t2 - t1, "kernel_ms": kernel_ms,
                          "sequences": reps, "groups_in_flight": len(gl),
                          "mpixels_per_s_e2e": reps * n * w * h / 1e6 / (t2 - t0),
                          "mpixels_per_s_kernels": reps * n * w * h / 1e3 / kernel_ms if kernel_ms else None,
                          "bytes": len(data), "fco_md5": md5, "md5_matches_reference": (md5 == want) if want else None,
                          "states": [int(done[f]["states"]) for f in range(min(n, 8))]}), flush=True)


if __name__ == "__main__":
    main()
