#!/usr/bin/env python
"""A/B helper: device-resident batch throughput of the library named by FB200_LIB (default: the product),
the same frames and launch shape as bench.py's `value` leg, without torch.  usage: ab_batch.py [frames] [steps]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cache = "/tmp/ab_frames_%d.npy" % n
if os.path.exists(cache):
    arr = np.load(cache)
else:
    arr = np.stack([gen_frames.chan(1024, 1024, 3 + k % 74) for k in range(min(n, 74))])
    arr = np.stack([arr[k % len(arr)] for k in range(n)])
    np.save(cache, arr)
p = ffi.make_params(1024, 1024, 1, 20.0, 0)
enc = F.TileEncoder(p, n)
enc.upload([ffi.pixels_from_grey(a).reshape(-1) for a in arr])
ms = []
for i in range(steps + 2):
    enc.launch(n)
    enc.sync()
    ms.append(enc.stats()["kernel_ms"])
enc.download(n)
print("%s: %d frames, kernel ms %s -> %.1f Mpx/s" % (os.environ.get("FB200_LIB", "product"), n,
      " ".join("%.1f" % m for m in ms), n * 1.048576 / (np.mean(ms[2:]) / 1e3)), flush=True)
enc.close()
