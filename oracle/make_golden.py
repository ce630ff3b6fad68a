#!/usr/bin/env python
"""Generate tests/golden/ from the REFERENCE binaries in oracle/_ref (test tooling).

Run in the build container (needs oracle/_ref, i.e. /root/reference):
    make -C oracle ref && python oracle/make_golden.py
Everything written is an OUTPUT of the unmodified reference (its .fco md5, the WFA its own
reader returns, the per-range trace at its inner seams, PSNR from its own dfiasco +
pnmpsnr), never reference source.
"""
import gzip
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_frames  # noqa: E402

REF = os.path.join(HERE, "_ref")
GOLD = os.path.join(HERE, "..", "tests", "golden")

# name -> (frame, crop spec or None, quality, optimize, keep trace?)
CASES = {
    "g256_q20_z0": ("g256", None, 20, 0, True),
    "g256_q20_z1": ("g256", None, 20, 1, True),
    "g256_q20_z2": ("g256", None, 20, 2, False),
    "g512_q20_z0": ("g512", None, 20, 0, False),
    "g1024_q20_z0": ("g1024", None, 20, 0, False),
    "c256_q20_z0": ("c256", None, 20, 0, True),
    "c256_q30_z0": ("c256", None, 30, 0, False),
    "g1024t0_q20_z0": ("g1024", (256, 0), 20, 0, False),
    "g1024t15_q20_z0": ("g1024", (256, 15), 20, 0, False),
    "c2048t0_q30_z0": ("c2048", (256, 0), 30, 0, False),
    "g4096t0_q20_z0": ("g4096", (512, 0), 20, 0, False),
    # ragged geometry: not a power of two, partly outside the bintree canvas
    "g1024r_q20_z0": ("g1024", ("rect", 0, 0, 200, 136), 20, 0, True),
    "g1024s_q40_z0": ("g1024", ("rect", 300, 500, 96, 64), 40, 0, False),
}


# predicted sequences: name -> (frames, width, height, quality, pattern)
VIDEOS = {
    "v160_q20_ippp": (4, 160, 128, 20, "ippp"),
    "v352_q30_ippip": (5, 352, 288, 30, "ippip"),   # the regenerated frames of this one: md5 only
    "v160_q20_ibbp": (7, 160, 128, 20, "ibbp"),     # B frames: forward, backward and interpolated prediction
}


def md5(b):
    return hashlib.md5(b).hexdigest()


def main():
    os.makedirs(GOLD, exist_ok=True)
    env = dict(os.environ, FIASCO_DATA=os.path.join(REF, "data"))
    manifest = {}
    cache = {}
    with tempfile.TemporaryDirectory() as tmp:
        env["FIASCO_IMAGES"] = tmp
        for name, (fr, crop, q, z, keep_trace) in CASES.items():
            if fr not in cache:
                cache[fr] = gen_frames.frame(fr)
            img = cache[fr]
            if crop is not None:
                if crop[0] == "rect":
                    _, x0, y0, w, h = crop
                    img = img[y0:y0 + h, x0:x0 + w].copy()
                else:
                    img = gen_frames.crops(img, crop[0])[crop[1]]
            ext = ".pgm" if img.ndim == 2 else ".ppm"
            pnm = os.path.join(tmp, name + ext)
            fco = os.path.join(tmp, name + ".fco")
            trace = os.path.join(tmp, name + ".trace")
            dec = os.path.join(tmp, name + ".dec" + ext)
            gen_frames.write_pnm(pnm, img)
            subprocess.run([os.path.join(REF, "seamdump"), pnm, fco, str(q), str(z), trace], check=True, env=env)
            # the same stream must come out of the reference CLI itself
            fco2 = os.path.join(tmp, name + ".cli.fco")
            subprocess.run([os.path.join(REF, "cfiasco"), "--progress-meter=0", "-V", "0", "-q", str(q), "-z", str(z),
                            "-i", pnm, "-o", fco2], check=True, env=env, stderr=subprocess.DEVNULL)
            fb = open(fco, "rb").read()
            assert fb == open(fco2, "rb").read(), name
            dump = subprocess.run([os.path.join(REF, "wfadump"), fco], check=True, env=env, capture_output=True).stdout
            subprocess.run([os.path.join(REF, "dfiasco"), "-o", dec, fco], check=True, env=env, stderr=subprocess.DEVNULL)
            pr = subprocess.run([os.path.join(REF, "pnmpsnr"), pnm, dec], check=True, env=env, capture_output=True)
            ps = (pr.stdout + pr.stderr).decode()
            psnr = [float(v) for v in re.findall(r"([0-9.]+) dB", ps)]
            with gzip.GzipFile(os.path.join(GOLD, name + ".wfa.gz"), "wb", mtime=0) as f:
                f.write(dump)
            # the frame as the coder itself regenerates it (decode_image, codec/coder.c:647)
            raw = os.path.join(tmp, name + ".raw")
            subprocess.run([os.path.join(REF, "decdump"), fco, raw], check=True, env=env, capture_output=True)
            decoded_md5 = md5(open(raw, "rb").read())
            tr = open(trace, "rb").read()
            lc = b"".join(l for l in tr.splitlines(True) if l.startswith(b"lc "))
            if keep_trace:
                with gzip.GzipFile(os.path.join(GOLD, name + ".trace.gz"), "wb", mtime=0) as f:
                    f.write(tr)
            manifest[name] = {
                "frame": fr, "crop": crop, "quality": q, "optimize": z,
                "width": int(img.shape[1]), "height": int(img.shape[0]), "color": int(img.ndim == 3),
                "pnm_md5": md5(gen_frames.pnm_bytes(img)), "fco_md5": md5(fb), "fco_bytes": len(fb),
                "lc_trace_md5": md5(lc), "lc_calls": lc.count(b"\n"),
                "psnr_db": psnr, "decoded_md5": decoded_md5,
            }
            print(name, manifest[name]["fco_md5"], len(fb), psnr, flush=True)
        # a short predicted sequence (BASELINE config 5 in small): I P P P, CLI defaults
        for name, (n, w, h, q, pattern) in VIDEOS.items():
            frames = list(gen_frames.video(n, w, h))
            for i, fr_ in enumerate(frames):
                gen_frames.write_pnm(os.path.join(tmp, "%s_%02d.pgm" % (name, i)), fr_)
            fco = os.path.join(tmp, name + ".fco")
            subprocess.run([os.path.join(REF, "cfiasco"), "--progress-meter=0", "-V", "0", "-q", str(q),
                            "--pattern=" + pattern, "-i", "%s_[00-%02d+1].pgm" % (name, n - 1), "-o", fco],
                           check=True, env=env, stderr=subprocess.DEVNULL)
            fb = open(fco, "rb").read()
            if len(frames) * w * h < 200000 and "b" not in pattern.lower():      # seam trace of the unmodified coder (oracle/seamdump.c)
                fco2, trace = os.path.join(tmp, name + ".seam.fco"), os.path.join(tmp, name + ".trace")
                subprocess.run([os.path.join(REF, "seamdump"), "%s_[00-%02d+1].pgm" % (name, n - 1), fco2, str(q), "0",
                                trace], check=True, env=env)
                assert open(fco2, "rb").read() == fb, name
                assert pattern.lower() == "ippppppppp"[:len(pattern)], "seamdump codes with the CLI's default pattern"
                with gzip.GzipFile(os.path.join(GOLD, name + ".trace.gz"), "wb", mtime=0) as f:
                    f.write(open(trace, "rb").read())
            dump = subprocess.run([os.path.join(REF, "wfadump"), fco], check=True, env=env, capture_output=True).stdout
            raw = os.path.join(tmp, name + ".raw")
            subprocess.run([os.path.join(REF, "decdump"), fco, raw], check=True, env=env, capture_output=True)
            rb = open(raw, "rb").read()
            with gzip.GzipFile(os.path.join(GOLD, name + ".wfa.gz"), "wb", mtime=0) as f:
                f.write(dump)
            if len(rb) < 200000:
                with gzip.GzipFile(os.path.join(GOLD, name + ".decoded.raw.gz"), "wb", mtime=0) as f:
                    f.write(rb)
            per = len(rb) // n
            manifest[name] = {
                "video": True, "frames": n, "width": w, "height": h, "quality": q, "pattern": pattern,
                "fco_md5": md5(fb), "fco_bytes": len(fb),
                "decoded_md5": [md5(rb[i * per:(i + 1) * per]) for i in range(n)],
            }
            print(name, manifest[name]["fco_md5"], len(fb), flush=True)
    kat = subprocess.run([os.path.join(REF, "seamdump"), "--kat"], check=True, capture_output=True).stdout
    with gzip.GzipFile(os.path.join(GOLD, "kat.txt.gz"), "wb", mtime=0) as f:
        f.write(kat)
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
