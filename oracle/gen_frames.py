#!/usr/bin/env python
"""Seeded synthetic PGM/PPM frames for parity tests and the benchmark (test tooling).

Value model: smooth sinusoid field + 40 random
rectangles + N(0,4) noise, clipped to u8, numpy default_rng(seed).  Frames are the
workload BASELINE.json's configs are quoted on:
  g256 (seed 1)  g512 (2)  g1024 (3)  g2048 (5)  g4096 (6)
  c256 (R,G,B = seeds 4,5,6)  c2048 (seeds 11,12,13)
Crops are plain array slices, row-major tile index k = ty*n + tx.
"""
import hashlib
import os
import sys

import numpy as np


def chan(w, h, s):
    """One 8-bit plane."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    r = np.random.default_rng(s)
    img = 128 + 60 * np.sin(x / (17 + s % 5)) * np.cos(y / (23 + s % 7)) + 40 * np.sin((x + y) / (61 + s % 3))
    for _ in range(40):
        x0, y0 = r.integers(0, w), r.integers(0, h)
        ww, hh = r.integers(8, w // 4), r.integers(8, h // 4)
        img[y0:y0 + hh, x0:x0 + ww] += r.integers(-50, 50)
    img += r.normal(0, 4, size=img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


GREY = {"g256": (256, 1), "g512": (512, 2), "g1024": (1024, 3), "g2048": (2048, 5), "g4096": (4096, 6)}
COLOUR = {"c256": (256, (4, 5, 6)), "c2048": (2048, (11, 12, 13))}


def frame(name):
    """Return the u8 array of a named frame: (h, w) grey or (h, w, 3) RGB."""
    if name in GREY:
        n, s = GREY[name]
        return chan(n, n, s)
    n, seeds = COLOUR[name]
    return np.stack([chan(n, n, s) for s in seeds], axis=-1)


def crops(img, tile):
    """Row-major list of tile x tile crops."""
    h, w = img.shape[:2]
    return [np.ascontiguousarray(img[y:y + tile, x:x + tile]) for y in range(0, h, tile) for x in range(0, w, tile)]


def pnm_bytes(img):
    if img.ndim == 2:
        return b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]) + img.tobytes()
    return b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]) + img.tobytes()


def write_pnm(path, img):
    with open(path, "wb") as f:
        f.write(pnm_bytes(img))


def video(n=30, w=720, h=576):
    """Config-5 frames: drifting background + 12 moving rectangles + N(0,2)."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    r = np.random.default_rng(7)
    base = 128 + 50 * np.sin(x / 19) * np.cos(y / 27) + 30 * np.sin((x + y) / 53)
    rects = [(r.integers(0, w - 120), r.integers(0, h - 120), r.integers(30, 120), r.integers(30, 120),
              r.integers(-60, 60), r.integers(-3, 4), r.integers(-3, 4)) for _ in range(12)]
    for f in range(n):
        img = np.roll(base, shift=f * 2, axis=1).copy()
        for (x0, y0, ww, hh, v, dx, dy) in rects:
            xx = int(np.clip(x0 + dx * f, 0, w - ww))
            yy = int(np.clip(y0 + dy * f, 0, h - hh))
            img[yy:yy + hh, xx:xx + ww] += v
        img += np.random.default_rng(100 + f).normal(0, 2, size=img.shape)
        yield np.clip(img, 0, 255).astype(np.uint8)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "."
    names = sys.argv[2:] or ["g256", "g512", "g1024", "c256"]
    os.makedirs(out, exist_ok=True)
    for nm in names:
        img = frame(nm)
        p = os.path.join(out, nm + (".pgm" if img.ndim == 2 else ".ppm"))
        write_pnm(p, img)
        print(nm, hashlib.md5(open(p, "rb").read()).hexdigest())


def box_blur(a, k):
    """Mean over a (2k)x(2k) window with replicated borders, back to u8."""
    a = a.astype(np.float32)
    c = np.cumsum(np.pad(a, ((k, k), (0, 0)), mode="edge"), axis=0)
    a = (c[2 * k:] - c[:-2 * k]) / (2 * k)
    c = np.cumsum(np.pad(a, ((0, 0), (k, k)), mode="edge"), axis=1)
    a = (c[:, 2 * k:] - c[:, :-2 * k]) / (2 * k)
    return np.clip(a, 0, 255).astype(np.uint8)


def colour_sequence(n=2, w=128, h=128, k=8):
    """A colour sequence whose FIRST frame is smooth (its luminance band stops above the finest range
    level), followed by detailed frames: the case where the reference's chroma set-up of frame j
    changes how frame j + 1 is coded (codec/coder.c:797).  With k = 12 the smooth frame is one the
    reference coder refuses ("Can't write more than N weights.", output/weights.c:137)."""
    out = [np.stack([box_blur(chan(w, h, s), k) for s in (4, 5, 6)], axis=-1)]
    for j in range(1, n):
        out.append(np.stack([chan(w, h, s + 20 * j) for s in (4, 5, 6)], axis=-1))
    return out


def nd_sequence(n=4, w=160, h=128, seed=7):
    """Bright frames with a fine moving texture: the kind of picture on which `cfiasco --prediction'
    takes nondeterministic (DC) prediction (codec/prediction.c:371) on a number of ranges."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    return [np.clip(190 + 25 * np.sin((xx + 3 * t) / 3.0) * np.sin((yy + t) / 5.0) + rng.normal(0, 5, (h, w)), 0,
                    255).astype(np.uint8) for t in range(n)]


def nd_still(n=512, seed=11):
    """A still for `--prediction': slow luminance ramp, fine texture, some noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n]
    return np.clip(150 + 60 * np.sin(xx / 40.0) + 25 * np.sin(xx / 3.0) * np.sin(yy / 5.0) + rng.normal(0, 6, (n, n)),
                   0, 255).astype(np.uint8)


def colour_video(n=4, w=160, h=128):
    """A colour sequence with motion: the frames of video() as the green channel, shifted / inverted
    copies of them plus a moving colour ramp as red and blue."""
    out = []
    yy, xx = np.mgrid[0:h, 0:w]
    for f, g in enumerate(video(n, w, h)):
        r = np.clip(np.roll(g, 5, axis=1).astype(np.int32) + 40 * np.sin((xx + 4 * f) / 23.0), 0, 255).astype(np.uint8)
        b = np.clip(255 - np.roll(g, -3, axis=0).astype(np.int32) + 30 * np.cos((yy - 2 * f) / 17.0), 0, 255).astype(np.uint8)
        out.append(np.stack([r, g, b], axis=-1))
    return out
