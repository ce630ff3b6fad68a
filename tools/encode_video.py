#!/usr/bin/env python
"""A sequence with P frames over 1..8 GPUs (BASELINE.json configs[4]: 30 frames 720x576 grey, q=20,
pattern IPPP, 2 GPUs): the groups of pictures are independent chains, group g goes to rank g mod world
(fiasco_b200/video.py); one gather brings the finished automata to rank 0, which writes the single stream.

    python tools/encode_video.py [--frames 30 --width 720 --height 576 --pattern ippp --out v.fco]     # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/encode_video.py
"""
import argparse
import hashlib
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from fiasco_b200 import ffi, hostlib, video  # noqa: E402
import gen_frames  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--width", type=int, default=720)
    ap.add_argument("--height", type=int, default=576)
    ap.add_argument("--pattern", default="ippp")
    ap.add_argument("--quality", type=float, default=20.0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    planes = [ffi.pixels_from_grey(f) for f in gen_frames.video(a.frames, a.width, a.height)]
    p = ffi.make_params(a.width, a.height, 1, a.quality, 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    seq, kernel_ms = video.encode_sequence(planes, a.pattern, p, rank, world,
                                           gather_device="cuda" if world > 1 else "cpu", device=local)
    md5 = None
    if rank == 0:
        with tempfile.TemporaryDirectory() as tmp:
            out = a.out or os.path.join(tmp, "v.fco")
            hostlib.write_video_stream(out, p, seq)
            data = open(out, "rb").read()
            md5 = hashlib.md5(data).hexdigest()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, kernel_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": "%d frames %dx%d grey q=%g pattern %s" % (a.frames, a.width, a.height, a.quality, a.pattern),
                          "groups": len(video.groups(a.frames, a.pattern)), "n_gpus": world,
                          "wall_s": float(t[0]), "kernel_ms_max_rank": float(t[1]),
                          "mpixels_per_s": a.frames * a.width * a.height / 1e6 / float(t[0]),
                          "bytes": len(data), "fco_md5": md5}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
