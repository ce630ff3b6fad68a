#!/usr/bin/env python
"""Golden vectors of `cfiasco --prediction' (nondeterministic prediction, codec/prediction.c:371), made by
the REFERENCE binaries in oracle/_ref (test tooling; run in the build container:
make -C oracle ref && python oracle/make_golden_r03.py).  Stream md5 / sizes into
tests/golden/manifest.json, the automata (oracle/wfadump.c) of the small cases beside it.

  nd160_q70_i      4 grey frames 160x128, all intra, q = 70: ranges with ND prediction in every frame
  nd160_q70_ippp   the same frames as IPPP: ND prediction in the I frame, motion compensation after it,
                   the ND tree in every frame's stream; two of its frames carry no edge at all and end
                   off a byte boundary
  nd512_q80        a 512^2 still, q = 80
  g256_q20_nd      the 256^2 frame of the other goldens: --prediction never wins, the stream differs from
                   g256_q20_z0 only by its (empty) ND trees
  c128_q30_nd      a colour still: the flag changes the kind of the delta pool, nothing is predicted

and of colour sequences with predicted frames (codec/coder.c:757-849, subtract_mc codec/mwfa.c:156):

  cv160_q20_ippp      4 colour frames 160x128, IPPP
  cv160_q20_ibbpbbp   7 frames with B frames (coded out of display order)
  cv160_q20_i         5 frames, all intra: the frames still depend on each other (lc_min_level, coder.c:797,
                      and the stale y_column entries of the virtual states, output/matrices.c:491)
  cv160_q25_ippibp    intra frames in the middle of a sequence
  cv160_q20_ippp_nd   IPPP with --prediction
  cv352_q35_ipp       3 frames 352x288
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_frames  # noqa: E402

REF = os.path.join(HERE, "_ref")
GOLD = os.path.join(HERE, "..", "tests", "golden")


def md5(b):
    return hashlib.md5(b).hexdigest()


def nd_labels(dump):
    """Ranges with ND prediction in a dump: labels that are subdivided AND have edges."""
    n, child = 0, {}
    for line in dump.splitlines():
        t = line.split()
        if t[0] == "frame":
            child = {}
        elif t[0] == "s":
            child[int(t[1])] = (int(t[3]), int(t[4]))
        elif t[0] == "e" and child[int(t[1])][int(t[2])] >= 0:
            n += 1
    return n


def cases():
    seq = gen_frames.nd_sequence()
    yield "nd160_q70_i", seq, 70, "i", True
    yield "nd160_q70_ippp", seq, 70, "ippp", True
    yield "nd512_q80", [gen_frames.nd_still()], 80, "i", True
    yield "g256_q20_nd", [gen_frames.frame("g256")], 20, "i", False
    yield "c128_q30_nd", [gen_frames.colour_sequence(2, 128, 128)[1]], 30, "i", False
    cv = gen_frames.colour_video(7, 160, 128)
    yield "cv160_q20_ippp", cv[:4], 20, "ippp", True
    yield "cv160_q20_ibbpbbp", cv, 20, "ibbpbbp", True
    yield "cv160_q20_i", cv[:5], 20, "i", False
    yield "cv160_q25_ippibp", cv, 25, "ippibp", False
    yield "cv160_q20_ippp_nd", cv[:4], 20, "ippp", False
    yield "cv352_q35_ipp", gen_frames.colour_video(3, 352, 288), 35, "ipp", False
    if "big" in sys.argv[1:]:       # BASELINE config 5 in colour: 30 frames 720x576 (minutes of reference time)
        yield "cv720_q20_ippp", gen_frames.colour_video(30, 720, 576), 20, "ippp", False


def main():
    path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(path))
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, FIASCO_DATA=os.path.join(REF, "data"), FIASCO_IMAGES=tmp)
        for key, frames, q, pattern, keep_dump in cases():
            if "big" in sys.argv[1:] and key != "cv720_q20_ippp":
                continue
            names = []
            for i, f in enumerate(frames):
                names.append(os.path.join(tmp, "%s_%d.%s" % (key, i, "pgm" if f.ndim == 2 else "ppm")))
                gen_frames.write_pnm(names[-1], f)
            fco = os.path.join(tmp, key + ".fco")
            nd = not key.startswith("cv") or key.endswith("_nd")
            subprocess.run([os.path.join(REF, "cfiasco"), "--progress-meter=0", "-V", "0", "-q", str(q),
                            *(["--prediction"] if nd else []), "--pattern=" + pattern, "-o", fco, *names], check=True,
                           env=env, stderr=subprocess.DEVNULL)
            fb = open(fco, "rb").read()
            dump = subprocess.run([os.path.join(REF, "wfadump"), fco], check=True, env=env, capture_output=True).stdout
            if keep_dump:
                with gzip.GzipFile(os.path.join(GOLD, key + ".wfa.gz"), "wb", mtime=0) as f:
                    f.write(dump)
            h, w = frames[0].shape[:2]
            manifest[key] = {"nd_prediction": nd, "frames": len(frames), "width": w, "height": h, "quality": q,
                             "pattern": pattern, "color": int(frames[0].ndim == 3), "fco_md5": md5(fb),
                             "fco_bytes": len(fb), "nd_ranges": nd_labels(dump.decode())}
            print(key, manifest[key], flush=True)
    with open(path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
