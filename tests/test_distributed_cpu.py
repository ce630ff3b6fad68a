"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: round-robin tile sharding and the single
gather of the per-tile FIASCO streams.  The encoder itself is exercised on the GPU; here the
streams are produced by the host writer from oracle automata (CPU)."""
import hashlib
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_tiles, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fiasco_b200 import ffi, hostlib, distributed as D
    import oracle_lib as O
    import gen_frames
    import tempfile
    crops = gen_frames.crops(gen_frames.frame("g256"), 64)[:n_tiles]
    p = ffi.make_params(64, 64, 1, 20.0, 0)
    local = {}
    with tempfile.TemporaryDirectory() as tmp:
        for i in D.shard(n_tiles, rank, world):
            w = O.encode(crops[i], quality=20, optimize=0)
            path = os.path.join(tmp, "t%d.fco" % i)
            hostlib.write_stream(path, p, [w])
            local[i] = open(path, "rb").read()
    allb = D.gather_streams(local, n_tiles, rank, world)
    if rank == 0:
        q.put([hashlib.md5(b).hexdigest() for b in allb])
    else:
        assert allb is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_is_a_partition():
    from fiasco_b200 import distributed as D
    for n in (1, 5, 16, 64):
        for world in (1, 2, 3, 8):
            got = sorted(i for r in range(world) for i in D.shard(n, r, world))
            assert got == list(range(n))
            sizes = [len(D.shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_equals_single_process():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from fiasco_b200 import ffi, hostlib
    import oracle_lib as O
    import gen_frames
    n_tiles = 5                                       # odd: ranks own 3 and 2 tiles
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_tiles, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    got = q.get(timeout=180)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    # single-process reference of the same streams
    import tempfile
    crops = gen_frames.crops(gen_frames.frame("g256"), 64)[:n_tiles]
    p = ffi.make_params(64, 64, 1, 20.0, 0)
    want = []
    with tempfile.TemporaryDirectory() as tmp:
        for i, c in enumerate(crops):
            path = os.path.join(tmp, "t%d.fco" % i)
            hostlib.write_stream(path, p, [O.encode(c, quality=20, optimize=0)])
            want.append(hashlib.md5(open(path, "rb").read()).hexdigest())
    assert got == want


# ---------------------------------------------------------------- sequences: groups of pictures over ranks

def _video_worker(rank, world, port, name, q):
    """A rank of the GOP-sharded sequence encoder.  No GPU here: the device library is the emulated build
    of the same sources (tests/emu, test infrastructure), so the whole path -- sharding, the chains on
    the 'device', finishing / regenerating on the host, the gather, the stream writer -- runs on CPU."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["FB200_NT"] = "128"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fiasco_b200 import ffi, hostlib, video
    emu = os.path.join(ROOT, "tests", "emu", "_build")
    ffi.lib_path = lambda: os.path.join(emu, "libfiasco_b200_emu.so")
    hostlib.lib_path = lambda: os.path.join(emu, "libfiasco_emu.so")
    import oracle_lib as O
    import gen_frames
    import tempfile
    m = O.manifest()[name]
    planes = [ffi.pixels_from_grey(f) for f in gen_frames.video(m["frames"], m["width"], m["height"])]
    p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
    seq, _ = video.encode_sequence(planes, m["pattern"], p, rank, world)
    if rank == 0:
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "v.fco")
            hostlib.write_video_stream(out, p, seq)
            q.put(hashlib.md5(open(out, "rb").read()).hexdigest())
    else:
        assert seq is None
    dist.barrier()
    dist.destroy_process_group()


def test_groups_of_pictures():
    from fiasco_b200 import video
    assert video.groups(5, "ippip") == [(0, 3), (3, 5)]
    assert video.groups(30, "ippp")[:3] == [(0, 4), (4, 8), (8, 12)] and len(video.groups(30, "ippp")) == 8
    assert video.groups(3, "p") == [(0, 3)]                     # frame 0 is always intra
    import pytest
    with pytest.raises(ValueError):
        video.groups(4, "ibbp")


def test_two_rank_sequence_equals_reference_stream():
    """BASELINE config 5's multi-GPU form at world size 2: the two groups of pictures of the golden IPPIP
    sequence on two ranks, one gather, one stream -- the reference coder's bytes."""
    import platform
    import subprocess
    import pytest
    if platform.machine() != "x86_64":
        pytest.skip("the emulator's context switch is x86-64")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    name = "v352_q30_ippip"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_video_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    got = q.get(timeout=300)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert got == O.manifest()[name]["fco_md5"]
