#!/usr/bin/env python
"""Tile-split encoding of one large frame over 1..8 GPUs (BASELINE.json configs[3]: 4096x4096 grey,
q=20, 64 tiles of 512x512): every tile is an independent FIASCO stream (identical to running the
reference cfiasco on the crop), tiles are sharded over the ranks, the finished .fco byte strings
are gathered on rank 0 with one collective.

    python tools/encode_tiled.py [--frame g4096 --tile 512 --out /tmp/tiles]            # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/encode_tiled.py
"""
import argparse
import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fiasco_b200 as F  # noqa: E402
from fiasco_b200 import ffi, hostlib, distributed as D  # noqa: E402
import gen_frames  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frame", default="g4096")
    ap.add_argument("--tile", type=int, default=512)
    ap.add_argument("--quality", type=float, default=20.0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    img = gen_frames.frame(a.frame)
    crops = gen_frames.crops(img, a.tile)
    mine = D.shard(len(crops), rank, world)
    p = ffi.make_params(a.tile, a.tile, 1, a.quality, 0)
    enc = F.TileEncoder(p, max(1, len(mine)), device=local)
    planes = [ffi.pixels_from_grey(crops[i]).reshape(-1) for i in mine]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    wfas, _ = enc.encode(planes)
    streams = {}
    with tempfile.TemporaryDirectory() as tmp:
        for i, w in zip(mine, wfas):
            path = os.path.join(tmp, "t%03d.fco" % i)
            hostlib.write_stream(path, p, [w])
            streams[i] = open(path, "rb").read()
    allb = D.gather_streams(streams, len(crops), rank, world, device="cuda" if world > 1 else "cpu")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        if a.out:
            os.makedirs(a.out, exist_ok=True)
            for i, b in enumerate(allb):
                open(os.path.join(a.out, "tile%03d.fco" % i), "wb").write(b)
        print(json.dumps({"frame": a.frame, "tiles": len(crops), "tile": a.tile, "n_gpus": world,
                          "seconds": float(t[0]), "mpixels_per_s": img.shape[0] * img.shape[1] / 1e6 / float(t[0]),
                          "bytes": sum(len(b) for b in allb), "tile0_md5": hashlib.md5(allb[0]).hexdigest()}))
    enc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
