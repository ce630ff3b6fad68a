#!/usr/bin/env python
"""Round-2 additions to tests/golden/manifest.json, made by the REFERENCE binaries in oracle/_ref
(test tooling; run in the build container: make -C oracle ref && python oracle/make_golden_r02.py).
md5 sums of reference streams only (small), plus two WFA dumps:

  cseq128_*         2-frame colour sequences, all intra, smooth frame first (lc_min_level carry-over,
                    codec/coder.c:797), and the same frames in the other order
  g1024_q20_z1/z2   the 1024^2 frame of BASELINE config 2 at optimisation levels 1 and 2
  tiles_g1024_256   BASELINE config 2 as 16 independent streams (256^2 crops, row-major)
  tiles_c2048_256   BASELINE config 3: 64 crops of 256^2, colour, q = 30
  tiles_g4096_512   BASELINE config 4: 64 crops of 512^2, q = 20
  v720_q20_ippp     BASELINE config 5: 30 frames 720x576, IPPP
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_frames  # noqa: E402

REF = os.path.join(HERE, "_ref")
GOLD = os.path.join(HERE, "..", "tests", "golden")


def md5(b):
    return hashlib.md5(b).hexdigest()


def cfiasco(env, out, names, q, extra=()):
    subprocess.run([os.path.join(REF, "cfiasco"), "--progress-meter=0", "-V", "0", "-q", str(q), *extra, "-o", out,
                    *names], check=True, env=env, stderr=subprocess.DEVNULL)
    return open(out, "rb").read()


def main():
    path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(path))
    only = set(sys.argv[1:])
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, FIASCO_DATA=os.path.join(REF, "data"), FIASCO_IMAGES=tmp)

        def want(k):
            return not only or k in only

        # colour sequences of intra frames
        seq = gen_frames.colour_sequence(2, 128, 128)
        for order, tag in (((0, 1), "sd"), ((1, 0), "ds")):
            names = []
            for i, k in enumerate(order):
                names.append(os.path.join(tmp, "cseq_%s_%d.ppm" % (tag, i)))
                gen_frames.write_pnm(names[-1], seq[k])
            for q, z in ((25, 0), (25, 1), (40, 1)):
                key = "cseq128_%s_q%d_z%d" % (tag, q, z)
                if not want(key):
                    continue
                fb = cfiasco(env, os.path.join(tmp, key + ".fco"), names, q, ("-z", str(z), "--pattern=i"))
                single = [md5(cfiasco(env, os.path.join(tmp, key + ".%d.fco" % i), [nm], q, ("-z", str(z))))
                          for i, nm in enumerate(names)]
                manifest[key] = {"colour_sequence": True, "order": list(order), "quality": q, "optimize": z,
                                 "width": 128, "height": 128, "pattern": "i", "fco_md5": md5(fb), "fco_bytes": len(fb),
                                 "single_frame_md5": single}
                print(key, md5(fb), len(fb), flush=True)
        # a frame the reference refuses: the failure must be reproduced, not "fixed"
        if want("csmooth128_q25_refused"):
            nm = os.path.join(tmp, "csmooth.ppm")
            gen_frames.write_pnm(nm, gen_frames.colour_sequence(1, 128, 128, 12)[0])
            r = subprocess.run([os.path.join(REF, "cfiasco"), "--progress-meter=0", "-V", "0", "-q", "25", "-o",
                                os.path.join(tmp, "csmooth.fco"), nm], env=env, capture_output=True)
            msg = r.stderr.decode().strip().splitlines()[-1]
            assert r.returncode != 0 and msg.startswith("Can't write more than"), (r.returncode, msg)
            manifest["csmooth128_q25_refused"] = {"refused": True, "quality": 25, "message": msg}
            print("csmooth128_q25_refused", msg, flush=True)
        # the big frame at -z 1 / -z 2
        for z in (1, 2):
            key = "g1024_q20_z%d" % z
            if not want(key):
                continue
            pnm = os.path.join(tmp, "g1024.pgm")
            gen_frames.write_pnm(pnm, gen_frames.frame("g1024"))
            fb = cfiasco(env, os.path.join(tmp, key + ".fco"), [pnm], 20, ("-z", str(z)))
            dump = subprocess.run([os.path.join(REF, "wfadump"), os.path.join(tmp, key + ".fco")], check=True, env=env,
                                  capture_output=True).stdout
            with gzip.GzipFile(os.path.join(GOLD, key + ".wfa.gz"), "wb", mtime=0) as f:
                f.write(dump)
            manifest[key] = {"frame": "g1024", "crop": None, "quality": 20, "optimize": z, "width": 1024, "height": 1024,
                             "color": 0, "fco_md5": md5(fb), "fco_bytes": len(fb)}
            print(key, md5(fb), len(fb), flush=True)
        # tile-split forms: every tile is the reference coder run on the crop
        for key, (fr, tile, q) in {"tiles_g1024_256": ("g1024", 256, 20), "tiles_c2048_256": ("c2048", 256, 30),
                                   "tiles_g4096_512": ("g4096", 512, 20)}.items():
            if not want(key):
                continue
            img = gen_frames.frame(fr)
            cr = gen_frames.crops(img, tile)
            ext = ".pgm" if img.ndim == 2 else ".ppm"

            def one(i):
                pnm = os.path.join(tmp, "%s_%02d%s" % (key, i, ext))
                gen_frames.write_pnm(pnm, cr[i])
                fb = cfiasco(env, os.path.join(tmp, "%s_%02d.fco" % (key, i)), [pnm], q)
                return md5(fb), len(fb)

            with ThreadPoolExecutor(8) as ex:
                res = list(ex.map(one, range(len(cr))))
            manifest[key] = {"tiles": True, "frame": fr, "tile": tile, "quality": q, "optimize": 0,
                             "fco_md5": [r[0] for r in res], "fco_bytes": [r[1] for r in res]}
            print(key, res[0], res[-1], sum(r[1] for r in res), flush=True)
        # BASELINE config 5 at full size
        if want("v720_q20_ippp"):
            names = []
            for i, f in enumerate(gen_frames.video(30, 720, 576)):
                names.append(os.path.join(tmp, "w%02d.pgm" % i))
                gen_frames.write_pnm(names[-1], f)
            fb = cfiasco(env, os.path.join(tmp, "v720.fco"), names, 20, ("--pattern=ippp",))
            manifest["v720_q20_ippp"] = {"video": True, "frames": 30, "width": 720, "height": 576, "quality": 20,
                                         "pattern": "ippp", "fco_md5": md5(fb), "fco_bytes": len(fb)}
            print("v720_q20_ippp", md5(fb), len(fb), flush=True)
    with open(path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
