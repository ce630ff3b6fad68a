#!/usr/bin/env python
"""Randomised campaign for the motion path WITHOUT a GPU: fiasco_coder() with the device sources run
under the thread emulator (tests/emu, test infrastructure) against the unmodified reference binary
(oracle/_ref/cfiasco) on random short sequences -- sizes, qualities, frame patterns with P and B frames,
grey and colour frames, with and without `--prediction' (nondeterministic prediction of the intra
frames), thread scheduling orders of the emulator.  The streams must be identical byte for byte.

    python tools/fuzz_video_emu.py [cases] [seed] [largest side, default 200]
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from fiasco_b200 import ffi, hostlib  # noqa: E402
import gen_frames  # noqa: E402

EMU = os.path.join(ROOT, "tests", "emu", "_build")
REF = os.path.join(ROOT, "oracle", "_ref")


def sequence(rng, n, w, h):
    """A drifting textured background with a few moving rectangles and noise."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    a, b, c = rng.uniform(7, 40, 3)
    base = 128 + 50 * np.sin(x / a) * np.cos(y / b) + 30 * np.sin((x + y) / c)
    rects = [(rng.integers(0, w - 16), rng.integers(0, h - 16), rng.integers(8, max(9, w // 3)),
              rng.integers(8, max(9, h // 3)), rng.integers(-60, 60), rng.integers(-4, 5), rng.integers(-4, 5))
             for _ in range(rng.integers(1, 6))]
    drift = int(rng.integers(-3, 4))
    sigma = float(rng.uniform(0, 3))
    for f in range(n):
        img = np.roll(base, shift=f * drift, axis=1).copy()
        for (x0, y0, ww, hh, v, dx, dy) in rects:
            xx = int(np.clip(x0 + dx * f, 0, max(0, w - ww)))
            yy = int(np.clip(y0 + dy * f, 0, max(0, h - hh)))
            img[yy:yy + hh, xx:xx + ww] += v
        img += rng.normal(0, sigma, size=img.shape)
        yield np.clip(img, 0, 255).astype(np.uint8)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    side = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    # FB200_EMU_ASAN=1 (after `make -C tests/emu asan`, with LD_PRELOAD=libasan.so): the sanitizer build
    dev = os.path.join(os.path.dirname(EMU), "_asan") if os.environ.get("FB200_EMU_ASAN") else EMU
    ffi.lib_path = lambda: os.path.join(dev, "libfiasco_b200_emu.so")
    hostlib.lib_path = lambda: os.path.join(dev, "libfiasco_emu.so")
    os.environ.setdefault("FB200_NT", "128")
    L = hostlib.load()
    cf = os.path.join(REF, "cfiasco")
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(cases):
        w = int(rng.integers(16, side // 2)) * 2
        h = int(rng.integers(16, side // 2)) * 2
        n = int(rng.integers(2, 7))
        q = float(rng.choice([5, 10, 20, 35, 60]))
        pattern = str(rng.choice(["ippp", "ip", "ippip", "ibbp", "ibp", "ibbbp", "ipbp", "ippppppppp", "i", "iip"]))
        order = int(rng.integers(0, 3))
        colour = bool(rng.integers(0, 2))
        nd = bool(rng.integers(0, 3) == 0)
        if colour and rng.integers(0, 4):
            # (colour frames of odd sizes make the reference coder itself fail most of the time -- "Can't write
            # more than N weights" -- which is reproduced but exercises little: mostly multiples of 32 here)
            w, h = max(64, w // 32 * 32), max(64, h // 32 * 32)
        os.environ["FB200_EMU_ORDER"] = str(order)
        only = os.environ.get("FUZZ_ONLY")         # replay one case of a campaign (the generator is still advanced)
        with tempfile.TemporaryDirectory() as tmp:
            names = []
            seq = list(sequence(rng, n, w, h))
            yy, xx = np.mgrid[0:h, 0:w]
            ca, cb = rng.uniform(15, 40, 2)
            for i in range(n):
                f = seq[i]
                if colour:
                    # correlated channels (three independent ones make the reference itself fail: "Can't write
                    # more than N weights", output/weights.c:137 -- reproduced, but that says little)
                    g = f.astype(np.int32)
                    r = np.clip(np.roll(g, 5, axis=1) + 40 * np.sin((xx + 4 * i) / ca), 0, 255).astype(np.uint8)
                    b = np.clip(255 - np.roll(g, -3, axis=0) + 30 * np.cos((yy - 2 * i) / cb), 0, 255).astype(np.uint8)
                    f = np.stack([r, f, b], axis=-1)
                names.append(os.path.join(tmp, "f%02d.%s" % (i, "ppm" if colour else "pgm")))
                gen_frames.write_pnm(names[-1], f)
            if only is not None and int(only) != case:
                continue
            print("case %d: %dx%d x%d %s q=%g pattern=%s%s order=%d ..." % (case, w, h, n, "colour" if colour else "grey",
                                                                          q, pattern, " --prediction" if nd else "",
                                                                          order), file=sys.stderr, flush=True)
            o = hostlib.cli_options(0)
            L.fiasco_c_options_set_frame_pattern(o, pattern.encode())
            L.fiasco_c_options_set_prediction(o, int(nd), 6, 10)
            out = os.path.join(tmp, "ours.fco")
            ok, msg = hostlib.coder(names, out, q, options=o)
            L.fiasco_c_options_delete(o)
            ref = os.path.join(tmp, "ref.fco")
            env = dict(os.environ, FIASCO_DATA=os.path.join(REF, "data"), FIASCO_IMAGES=tmp)
            r = subprocess.run([cf, "--progress-meter=0", "-V", "0", "-q", str(q), "--pattern=" + pattern]
                               + (["--prediction"] if nd else []) + ["-o", ref] + names, env=env, capture_output=True)
            if r.returncode != 0:
                status = "reference failed (rc %d), ours: %s" % (r.returncode, "ok" if ok else "refused: " + msg.splitlines()[0])
            elif not ok:
                status = "MISMATCH: ours refused: " + msg
                bad += 1
            else:
                same = hashlib.md5(open(out, "rb").read()).hexdigest() == hashlib.md5(open(ref, "rb").read()).hexdigest()
                status = "identical (%d bytes)" % os.path.getsize(ref) if same else "MISMATCH"
                bad += 0 if same else 1
        print("case %d: %dx%d x%d %s q=%g pattern=%s%s order=%d: %s" % (case, w, h, n, "colour" if colour else "grey", q,
                                                                    pattern, " --prediction" if nd else "", order,
                                                                    status), flush=True)
    print("%d cases, %d mismatches" % (cases, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
