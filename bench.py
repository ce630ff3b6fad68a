#!/usr/bin/env python
"""bench.py -- FIASCO encoder hot path throughput, Mpixels/s at fixed PSNR.

Workload (BASELINE.json configs[1]): 1024x1024 greyscale frames, quality 20, CLI defaults
(-z 0), encoded MONOLITHICALLY -- the form the reference really encodes whatever
--tiling-exponent says (tiling is dead code there) -- so every automaton is
bit-identical to the reference CPU coder's and the PSNR is the reference's by construction.
One "step" = one batch of B independent seeded synthetic frames (one thread block per frame,
B defaults to the SM count, per GPU).  Multi-GPU: frames are independent, so each rank takes
its own B frames, no data-path collective (weak scaling); rank 0 gathers the per-rank times.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Prints ONE JSON line (rank 0).  --impl reference times the reference's own CPU implementation
(oracle/_ref/cfiasco, the unmodified reference built in place; else the oracle port) on all
host cores for the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# one process per GPU: every rank sees only its own device (as device 0), so that the library's own
# entry points -- fiasco_coder() included, which uses devices 0 .. FIASCO_GPUS - 1 -- run on it
if "LOCAL_RANK" in os.environ and "reference" not in sys.argv[1:]:
    _vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    _ids = _vis.split(",") if _vis else None
    _lr = int(os.environ["LOCAL_RANK"])
    os.environ["CUDA_VISIBLE_DEVICES"] = _ids[_lr] if _ids and _lr < len(_ids) else str(_lr)

W_ = H_ = 1024
QUALITY = 20.0
LAP_NAMES = ["ctrl", "pix", "dots", "upsweep", "enter", "mp_pro", "mp_p1", "mp_waves", "mp_commit", "mp_ortho",
             "ar_epi", "ap_img", "ap_direct", "ap_staged", "decide", "cluster"]
METRIC = "encoder Mpixels/s at fixed PSNR (1024^2 grey, q=20)"


_GEN = """
import sys, numpy as np, multiprocessing as mp
sys.path.insert(0, sys.argv[1])
import gen_frames
def one(seed):
    return gen_frames.chan(int(sys.argv[2]), int(sys.argv[3]), seed)
if __name__ == "__main__":
    seeds = range(int(sys.argv[4]), int(sys.argv[4]) + int(sys.argv[5]))
    with mp.get_context("fork").Pool(int(sys.argv[6])) as pool:
        np.save(sys.argv[7], np.stack(pool.map(one, seeds, chunksize=8)))
"""


def frames(first, count):
    """Seeded synthetic frames (SURVEY.md 8d value model); frame k uses seed 3 + k so frame 0 is
    the g1024 golden frame.  75 ms of numpy per frame: large counts are generated on all host cores
    by a helper process (a clean interpreter, so that forking never meets an initialised CUDA)."""
    import gen_frames
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = min(16, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))))
    if count < 32 or workers < 2:
        return [gen_frames.chan(W_, H_, 3 + first + k) for k in range(count)]
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "frames.npy")
        subprocess.run([sys.executable, "-c", _GEN, os.path.join(ROOT, "oracle"), str(W_), str(H_), str(3 + first),
                        str(count), str(workers), out], check=True)
        arr = np.load(out)
    return [arr[k] for k in range(count)]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                pass
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm

def ref_tools():
    c = os.path.join(ROOT, "oracle", "_ref", "cfiasco")
    return c if os.path.exists(c) else None


def cpu_encode_frames(imgs, workers):
    """Encode frames on the host with the reference binary (kind 'reference') or, if it is not
    there, the oracle port (kind 'port'); `workers` concurrent single-threaded processes.
    Returns (seconds, kind)."""
    import gen_frames
    cf = ref_tools()
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i, im in enumerate(imgs):
            p = os.path.join(tmp, "f%03d.pgm" % i)
            gen_frames.write_pnm(p, im)
            paths.append(p)
        env = dict(os.environ, FIASCO_DATA=os.path.join(ROOT, "oracle", "_ref", "data"), FIASCO_IMAGES=tmp)
        if cf:
            cmds = [[cf, "--progress-meter=0", "-V", "0", "-q", str(QUALITY), "-i", p, "-o", p + ".fco"] for p in paths]
            kind = "reference"
        else:
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
            oe = os.path.join(ROOT, "oracle", "_build", "oracle_enc")
            cmds = [[oe, p, str(QUALITY), "0", os.devnull] for p in paths]
            kind = "port"
        t0 = time.perf_counter()
        running, todo = [], list(cmds)
        while todo or running:
            while todo and len(running) < workers:
                running.append(subprocess.Popen(todo.pop(0), env=env, stdout=subprocess.DEVNULL,
                                                stderr=subprocess.DEVNULL))
            running[0].wait()
            if running[0].returncode != 0:
                raise RuntimeError("CPU encoder failed")
            running.pop(0)
        return time.perf_counter() - t0, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    imgs = frames(0, cores)                       # one frame per host core per step
    for _ in range(args.warmup):
        cpu_encode_frames(imgs[:cores], cores)
    t = 0.0
    kind = "port"
    for _ in range(args.steps):
        dt, kind = cpu_encode_frames(imgs, cores)
        t += dt
    mpx = len(imgs) * W_ * H_ / 1e6
    value = mpx * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1024x1024 grey frames, q=20, cfiasco defaults (-z 0), monolithic; one frame per "
                               "host core per step, one single-threaded reference process per core",
                   "frames_per_step": len(imgs)},
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": cores, "kind": kind,
                         "sample": "%d frames/step x %d steps, %d concurrent processes" % (len(imgs), args.steps, cores)},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm

def md5_of(path):
    import hashlib
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def scratch_dir(need_bytes):
    """A directory for the input frames 'on disk': tmpfs when it has the room (the metric excludes the
    medium: SURVEY 8d 'image already on disk'), else the default temporary directory."""
    import shutil
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            if shutil.disk_usage(d).free > 1.5 * need_bytes:
                return tempfile.mkdtemp(prefix="fb200_bench_", dir=d)
        except OSError:
            pass
    return tempfile.mkdtemp(prefix="fb200_bench_")


def coder_call(inputs, out, quality, pattern=None, optimize=0, env=None, prediction=False):
    """One fiasco_coder() call (the reference's public entry point, include/fiasco.h) with the CLI's
    default options.  Returns (seconds, work counters of the call)."""
    from fiasco_b200 import ffi, hostlib
    L = hostlib.load()
    o = hostlib.cli_options(optimize)
    L.fiasco_c_options_set_progress_meter(o, 0)
    if pattern:
        L.fiasco_c_options_set_frame_pattern(o, pattern.encode())
    if prediction:                                 # cfiasco --prediction (nd_prediction, codec/prediction.c:371)
        L.fiasco_c_options_set_prediction(o, 1, 6, 10)
    saved = {k: os.environ.get(k) for k in (env or {})}
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    ffi.counters(reset=True)
    t0 = time.perf_counter()
    try:
        ok, msg = hostlib.coder(inputs, out, quality=quality, options=o)
    finally:
        dt = time.perf_counter() - t0
        L.fiasco_c_options_delete(o)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if not ok:
        raise RuntimeError("fiasco_coder: " + msg)
    return dt, ffi.counters()


def config_records(tmp):
    """BASELINE.json configs 2 (tile-split form), 3, 4 and 5 through fiasco_coder(): files on disk ->
    .fco files, the library's FIASCO_TILE_SPLIT mode, bytes compared with the reference
    coder's (tests/golden/manifest.json: the reference binary run on the crops / the sequence)."""
    import gen_frames
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    out = {}
    for key, mkey, split in (("c2_tiles16", "tiles_g1024_256", 4), ("c3", "tiles_c2048_256", 6),
                             ("c4", "tiles_g4096_512", 6)):
        m = man[mkey]
        img = gen_frames.frame(m["frame"])
        pnm = os.path.join(tmp, key + (".pgm" if img.ndim == 2 else ".ppm"))
        gen_frames.write_pnm(pnm, img)
        dst = os.path.join(tmp, key + ".fco")
        best, cnt = None, None
        for _ in range(3):
            dt, c = coder_call([pnm], dst, float(m["quality"]), env={"FIASCO_TILE_SPLIT": split, "FIASCO_GPUS": 1})
            if best is None or dt < best:
                best, cnt = dt, c
        n = 1 << split
        got = [md5_of(os.path.join(tmp, "%s.t%02d.fco" % (key, t))) for t in range(n)]
        mpx = img.shape[0] * img.shape[1] / 1e6
        out[key] = {"workload": "%s as %d streams of %dx%d, q=%g" % (m["frame"], n, m["tile"], m["tile"], m["quality"]),
                    "mpixels_per_s": mpx / best, "seconds": best, "kernel_ms": cnt["kernel_ms"],
                    "kernel_launches": cnt["launches"], "md5_matches_reference": got == m["fco_md5"],
                    "through": "fiasco_coder(), FIASCO_TILE_SPLIT=%d, best of 3 calls (context creation included)" % split}
    m = man["v720_q20_ippp"]
    vd = os.path.join(tmp, "video")
    os.makedirs(vd, exist_ok=True)
    for i, f in enumerate(gen_frames.video(m["frames"], m["width"], m["height"])):
        gen_frames.write_pnm(os.path.join(vd, "w%02d.pgm" % i), f)
    dst = os.path.join(tmp, "c5.fco")
    best, cnt = None, None
    for _ in range(2):
        dt, c = coder_call([os.path.join(vd, "w[00-%02d].pgm" % (m["frames"] - 1))], dst, float(m["quality"]),
                           pattern=m["pattern"], env={"FIASCO_GPUS": 1})
        if best is None or dt < best:
            best, cnt = dt, c
    mpx = m["frames"] * m["width"] * m["height"] / 1e6
    out["c5"] = {"workload": "%d frames %dx%d, pattern %s, q=%g" % (m["frames"], m["width"], m["height"], m["pattern"],
                                                                      m["quality"]),
                 "mpixels_per_s": mpx / best, "seconds": best, "kernel_ms": cnt["kernel_ms"],
                 "kernel_launches": cnt["launches"], "md5_matches_reference": md5_of(dst) == m["fco_md5"],
                 "through": "fiasco_coder(), one process, one GPU, best of 2 calls"}
    # beyond BASELINE's list: the same sequence geometry in colour (every frame of a colour sequence depends on the
    # frame coded before it -- range levels, coder.c:797, y_column history -- so the frames run one after the other),
    # and a still with nondeterministic prediction (cfiasco --prediction)
    if "cv720_q20_ippp" in man:
        m = man["cv720_q20_ippp"]
        cd = os.path.join(tmp, "cvideo")
        os.makedirs(cd, exist_ok=True)
        for i, f in enumerate(gen_frames.colour_video(m["frames"], m["width"], m["height"])):
            gen_frames.write_pnm(os.path.join(cd, "c%02d.ppm" % i), f)
        dst = os.path.join(tmp, "c5c.fco")
        dt, cnt = coder_call([os.path.join(cd, "c[00-%02d].ppm" % (m["frames"] - 1))], dst, float(m["quality"]),
                             pattern=m["pattern"], env={"FIASCO_GPUS": 1})
        out["c5_colour"] = {"workload": "%d COLOUR frames %dx%d, pattern %s, q=%g" % (m["frames"], m["width"], m["height"],
                                                                                    m["pattern"], m["quality"]),
                            "mpixels_per_s": m["frames"] * m["width"] * m["height"] / 1e6 / dt, "seconds": dt,
                            "kernel_ms": cnt["kernel_ms"], "kernel_launches": cnt["launches"],
                            "md5_matches_reference": md5_of(dst) == m["fco_md5"],
                            "through": "fiasco_coder(), one process, one GPU, one call"}
    if "nd512_q80" in man:
        m = man["nd512_q80"]
        pnm = os.path.join(tmp, "nd512.pgm")
        gen_frames.write_pnm(pnm, gen_frames.nd_still())
        dst = os.path.join(tmp, "nd512.fco")
        best, cnt = None, None
        for _ in range(3):
            dt, c = coder_call([pnm], dst, float(m["quality"]), env={"FIASCO_GPUS": 1}, prediction=True)
            if best is None or dt < best:
                best, cnt = dt, c
        out["nd512"] = {"workload": "512x512 grey still, q=%g, --prediction (%d ranges predicted)" % (m["quality"],
                                                                                                   m["nd_ranges"]),
                        "mpixels_per_s": 512 * 512 / 1e6 / best, "seconds": best, "kernel_ms": cnt["kernel_ms"],
                        "kernel_launches": cnt["launches"], "md5_matches_reference": md5_of(dst) == m["fco_md5"],
                        "through": "fiasco_coder(), best of 3 calls"}
    return out


def strong_records(rank, world, dist, torch):
    """Strong scaling: ONE job split over the ranks, the gather inside the timed region.  Config 4: the 64
    tiles of the 4096^2 frame dealt to the ranks, every rank codes its tiles and writes their .fco bytes,
    two all_gathers (sizes, padded payloads; NCCL) bring them to rank 0.  Config 5: the 8 groups of pictures
    dealt to the ranks, the automata gathered the same way, rank 0 writes the stream."""
    import gen_frames
    import fiasco_b200 as F
    from fiasco_b200 import ffi, hostlib, distributed as D, video as V
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    dev = "cuda" if world > 1 else "cpu"
    rec = {}

    def timed(fn, reps):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            res = fn()
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if best is None or float(t[0]) < best[0]:
                best = (float(t[0]), res)
        return best

    # ---- config 4
    m = man["tiles_g4096_512"]
    img = gen_frames.frame(m["frame"])
    crops = gen_frames.crops(img, m["tile"])
    mine = D.shard(len(crops), rank, world)
    p = ffi.make_params(m["tile"], m["tile"], 1, float(m["quality"]), 0)
    enc = F.TileEncoder(p, max(1, len(mine)), device=0)
    planes = [ffi.pixels_from_grey(crops[i]).reshape(-1) for i in mine]
    tmp = tempfile.mkdtemp(prefix="fb200_strong_")

    def job4():
        wfas, _ = enc.encode(planes)
        streams = {}
        for i, w in zip(mine, wfas):
            path = os.path.join(tmp, "t%03d.fco" % i)
            hostlib.write_stream(path, p, [w])
            streams[i] = open(path, "rb").read()
        return D.gather_streams(streams, len(crops), rank, world, device=dev), enc.stats()["kernel_ms"]

    job4()
    sec, (allb, kms) = timed(job4, 3)
    enc.close()
    if rank == 0:
        import hashlib
        ok = [hashlib.md5(b).hexdigest() for b in allb] == m["fco_md5"]
        rec["c4"] = {"workload": "g4096 as 64 streams of 512x512 dealt to %d rank(s)" % world, "seconds": sec,
                     "mpixels_per_s": img.shape[0] * img.shape[1] / 1e6 / sec, "kernel_ms_rank0": kms,
                     "md5_matches_reference": ok, "collective": "2 x all_gather (sizes, padded .fco payloads), "
                     + ("nccl" if world > 1 else "none (one rank)"), "timed": "planes on the host -> tile kernel -> "
                     ".fco bytes of every tile on rank 0, max over ranks, best of 3"}
    # ---- config 5
    m = man["v720_q20_ippp"]
    seq = [ffi.pixels_from_grey(f) for f in gen_frames.video(m["frames"], m["width"], m["height"])]
    pv = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)

    def job5():
        auto, ms = V.encode_sequence(seq, m["pattern"], pv, rank=rank, world=world, gather_device=dev)
        if rank == 0:
            path = os.path.join(tmp, "v.fco")
            hostlib.write_video_stream(path, pv, auto)
            return md5_of(path), ms
        return None, ms

    job5()
    sec, (digest, kms) = timed(job5, 2)
    if rank == 0:
        rec["c5"] = {"workload": "30 frames 720x576 IPPP: 8 groups of pictures dealt to %d rank(s)" % world,
                     "seconds": sec, "mpixels_per_s": m["frames"] * m["width"] * m["height"] / 1e6 / sec,
                     "kernel_ms_rank0": kms, "md5_matches_reference": digest == m["fco_md5"],
                     "collective": "2 x all_gather (sizes, packed automata), " + ("nccl" if world > 1 else "none (one rank)"),
                     "timed": "planes on the host -> I launch + 3 P steps per rank -> automata gathered -> "
                              "stream written by rank 0, max over ranks, best of 2"}
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return rec


def run_ours(args):
    import torch
    import fiasco_b200 as F
    from fiasco_b200 import ffi
    import gen_frames

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
    if not torch.cuda.is_available() or F.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    if world > 1 and "FIASCO_HOST_THREADS" not in os.environ:
        # the ranks of one box share its cores: the library's host threads (PNM readers, stream writers) are dealt out
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        os.environ["FIASCO_HOST_THREADS"] = str(max(2, min(16, cores // world)))
    local = 0                                       # CUDA_VISIBLE_DEVICES holds this rank's GPU only
    torch.cuda.set_device(local)
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    p = ffi.make_params(W_, H_, 1, QUALITY, 0)
    probe = F.TileEncoder(p, 1, device=local)
    resident = probe.resident_tiles() or sms         # SM count x resident thread blocks per SM
    probe.close()
    # default: three waves of frames per launch -- the frames take different times, later waves fill
    # the SMs that finish early (the big tables exist once per RESIDENT frame, see ffi.cu)
    B = args.batch if args.batch else args.waves * resident
    imgs = frames(rank * B, B)
    planes = [ffi.pixels_from_grey(im).reshape(-1) for im in imgs]
    # inputs of the C-ABI end-to-end leg live in pinned host memory
    pinned = torch.empty((B, W_ * H_), dtype=torch.int16).pin_memory()
    pinned.numpy()[:] = np.stack(planes)
    host_planes = [pinned.numpy()[i] for i in range(B)]

    enc = F.TileEncoder(p, B, device=local)
    # a dedicated (non-default) torch stream: the kernel, the L2 flush and the timing events all go
    # through it, so torch.cuda.Event brackets exactly our launches
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    # ---- device-resident leg: inputs already in HBM, kernel only ----
    enc.upload(host_planes)
    l2_flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(max(args.warmup, 3)):
        enc.launch(B, stream)
    barrier()
    sampler = ClockSampler(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.perf_counter()
    for a, b in ev:
        l2_flush.zero_()                          # flush L2 between timed iterations (not timed)
        a.record()
        enc.launch(B, stream)
        b.record()
    barrier()
    wall = time.perf_counter() - t0
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    clocks = sampler.stop() if rank == 0 else None
    step_ms = float(np.sum(kernel_ms))
    enc.download(B)
    st = enc.stats()

    # ---- end to end through the C ABI: pinned host planes -> fb200_encode_tiles -> automata on the host ----
    enc.encode(host_planes)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        enc.encode(host_planes)
    torch.cuda.synchronize()
    abi_s = time.perf_counter() - t0
    st2 = enc.stats()
    enc.close()
    del pinned, host_planes, l2_flush
    torch.cuda.empty_cache()

    # ---- end to end through the reference's own entry point: fiasco_coder() ----
    # (ii) throughput: the B frames of this rank as PGM files on disk, one --pattern=i sequence -> one .fco
    tmp = scratch_dir(B * (W_ * H_ + 64))
    try:
        for k, im in enumerate(imgs):
            gen_frames.write_pnm(os.path.join(tmp, "f%04d.pgm" % k), im)
        template = os.path.join(tmp, "f[0000-%04d].pgm" % (B - 1))
        seq_out = os.path.join(tmp, "seq.fco")
        coder_call([template], seq_out, QUALITY, pattern="i")          # warm-up (page cache, CUDA context)
        barrier()
        t0 = time.perf_counter()
        seq_cnt = None
        for _ in range(e2e_steps):
            _, seq_cnt = coder_call([template], seq_out, QUALITY, pattern="i")
        seq_s = time.perf_counter() - t0
        seq_bytes = os.path.getsize(seq_out)
        # (i) latency: one 1024^2 frame per call (what cfiasco does), context creation included
        one = os.path.join(tmp, "f0000.pgm")
        one_out = os.path.join(tmp, "one.fco")
        coder_call([one], one_out, QUALITY)
        lat_steps = 20
        barrier()
        lat, lat_kernel = [], []
        for _ in range(lat_steps):
            dt, c = coder_call([one], one_out, QUALITY)
            lat.append(dt)
            lat_kernel.append(c["kernel_ms"])
        one_md5_ok = (md5_of(one_out) == json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
                      ["g1024_q20_z0"]["fco_md5"]) if rank == 0 else None
        configs = config_records(tmp) if rank == 0 and world == 1 and not args.no_configs else None
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    strong = strong_records(rank, world, dist, torch) if not args.no_configs else None

    step_ms_max, abi_max, seq_max = max_over_ranks(step_ms, abi_s, seq_s)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    mpx_step = world * B * W_ * H_ / 1e6
    value = mpx_step * args.steps / (step_ms_max / 1e3)
    abi_value = mpx_step * e2e_steps / abi_max
    seq_value = mpx_step * e2e_steps / seq_max
    peak, peak_src = peaks()
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):                        # DRAM bytes per frame measured by ncu (never under this run)
        t_ = json.load(open(tj))
        traffic, traffic_src = t_["dram_bytes_per_frame"] * B, t_["source"]
    alg_bytes = st["ip_bytes"] + st["mp_bytes"] + st["ss_bytes"]      # per launch, this rank
    launch_ms = step_ms / args.steps
    achieved = alg_bytes / (launch_ms / 1e3) / 1e9
    # the inner-product phase (pixels -> range x state products, SURVEY 8d "k_leaf_ip") seen alone: its
    # algorithmic bytes over the time the resident thread blocks spent in it (in-kernel cycle counters)
    ip_phase = None
    if st["cyc_T"] and clocks and clocks.get("sm_mhz"):
        ip_s = st["cyc_T"] / (clocks["sm_mhz"] * 1e6) / min(B, resident)
        ip_gbs = st["ip_bytes"] / ip_s / 1e9
        ip_phase = {"achieved": ip_gbs, "unit": "GB/s", "frac": ip_gbs / peak, "share_of_launch": ip_s / (launch_ms / 1e3),
                    "how": "ip_bytes of one launch / (sum of the blocks' cycles in the phase / SM clock / resident blocks)"}
    # CPU baseline: the reference (or the port) on ONE core, a bounded sample of the same workload
    cpu_s, kind = cpu_encode_frames(imgs[:2], 1)
    cpu_value = 2 * W_ * H_ / 1e6 / cpu_s
    lat_ms = float(np.median(lat)) * 1e3
    line = {
        "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch of %d independent 1024x1024 grey frames per GPU, q=20, cfiasco defaults "
                               "(-z 0), monolithic (bit-identical to the reference coder); one thread block per "
                               "frame, %d frames resident per SM, %.1f waves per launch"
                               % (B, max(1, min(B, resident) // sms), B / float(resident)),
                   "frames_per_step_per_gpu": B, "timing": "CUDA events on the launch stream, L2 flushed "
                   "(192 MiB memset) between timed launches", "states_per_frame": st["states"] / B,
                   "wall_s_timed_region": wall},
        "e2e": {"value": seq_value, "unit": "Mpixels/s", "h2d_bytes_per_step": int(seq_cnt["h2d_bytes"]),
                "d2h_bytes_per_step": int(seq_cnt["d2h_bytes"]), "steps": e2e_steps,
                "through": "fiasco_coder() (include/fiasco.h, the reference's public entry point): %d PGM files on disk "
                           "(%s), frame pattern `i' -> one .fco file (%d bytes); PNM parsing, context creation, "
                           "host-to-device copies, tile kernel, device-to-host copies and the stream writer all inside "
                           "the timed region; wall clock, max over ranks" % (B, "tmpfs" if tmp.startswith("/dev/shm")
                                                                             else "temporary directory", seq_bytes),
                "kernel_ms_per_step": seq_cnt["kernel_ms"], "kernel_launches_per_step": seq_cnt["launches"]},
        "e2e_latency": {"value": W_ * H_ / 1e6 / (lat_ms / 1e3), "unit": "Mpixels/s", "ms_per_frame": lat_ms,
                        "kernel_ms_per_frame": float(np.median(lat_kernel)), "steps": lat_steps,
                        "md5_matches_reference": one_md5_ok,
                        "through": "fiasco_coder() on ONE 1024x1024 PGM per call (what cfiasco does): file -> .fco, "
                                   "context creation included; the frame runs on a cluster of thread blocks; median"},
        "e2e_abi": {"value": abi_value, "unit": "Mpixels/s", "h2d_bytes_per_step": int(st2["h2d_bytes"]),
                    "d2h_bytes_per_step": int(st2["d2h_bytes"]), "steps": e2e_steps,
                    "through": "fb200_encode_tiles() (include/fiasco_b200.h): pinned host int16 planes -> H2D -> tile "
                               "kernel -> D2H automata"},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "fiasco_tile_kernel",
                     "algorithmic_bytes_per_launch": int(alg_bytes),
                     "bytes_model": "sum over lc_max blocks (4*2^lc_max + 376*S_b) + 8D per pursuit + 4D per "
                                    "Gram-Schmidt step + 4*levels*(s+1) per new state (SURVEY.md 8d)",
                     "note": "the path is bound by its serial chain (~20k dependent pursuits per frame; ncu: barrier waits, "
                             "instruction fetch, fixed-latency dependencies), not by HBM; traffic > algorithmic bytes "
                             "because the tables of the frames in flight exceed L2 (profiles/README.md)"},
        "cpu_baseline": {"value": cpu_value, "unit": "Mpixels/s", "cores": 1, "kind": kind,
                         "sample": "2 of the %d frames, one single-threaded process (%.1f s)" % (B, cpu_s)},
        "clocks": clocks,
        "phase_cycles": {k: int(st[k]) for k in ("cyc_total", "cyc_T", "cyc_mp", "cyc_append")},
        "ip_phase": ip_phase,
        "work": {k: int(st[k]) for k in ("mp_calls", "mp_steps", "blocks", "states")},
        "configs": configs,
        "strong": strong,
    }
    if sum(st["lap"]):                            # only a -DFB200_LAPS diagnostics build fills the lap timers
        line["lap_share"] = {n: round(v / float(sum(st["lap"])), 4) for n, v in zip(LAP_NAMES, st["lap"])}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (default: waves x resident frames)")
    ap.add_argument("--waves", type=int, default=3, help="frames per step in units of the resident frames per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config and strong-scaling sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
