"""End-to-end through the PUBLIC libfiasco API on the GPU: fiasco_coder() and the unchanged
reference command line front end (cfiasco, linked against our library) must write .fco files
byte-identical to the reference coder's (golden md5 from tests/golden/manifest.json), and the
reference's own decoder must reproduce the reference PSNR from them."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

from fiasco_b200 import hostlib
import oracle_lib as O
import gen_frames

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("name", ["g256_q20_z0", "g256_q20_z1", "g256_q20_z2", "g1024t15_q20_z0", "g1024r_q20_z0",
                                  "g512_q20_z0"])
def test_fiasco_coder_stream_md5(name, tmp_path):
    m = O.manifest()[name]
    pnm = str(tmp_path / (name + ".pgm"))
    gen_frames.write_pnm(pnm, O.case_image(name))
    out = str(tmp_path / (name + ".fco"))
    ok, msg = hostlib.coder([pnm], out, quality=float(m["quality"]), optimize=m["optimize"])
    assert ok, msg
    assert md5(out) == m["fco_md5"]


@pytest.mark.parametrize("name", ["c256_q20_z0", "c256_q30_z0", "c2048t0_q30_z0"])
def test_fiasco_coder_colour_stream_md5(name, tmp_path):
    m = O.manifest()[name]
    pnm = str(tmp_path / (name + ".ppm"))
    gen_frames.write_pnm(pnm, O.case_image(name))
    out = str(tmp_path / (name + ".fco"))
    ok, msg = hostlib.coder([pnm], out, quality=float(m["quality"]), optimize=m["optimize"])
    assert ok, msg
    assert md5(out) == m["fco_md5"]


def test_fiasco_coder_1024_md5(tmp_path):
    """BASELINE.json config[1] end to end: same bytes as the reference cfiasco (c5f96a1d...)."""
    name = "g1024_q20_z0"
    pnm = str(tmp_path / "g1024.pgm")
    gen_frames.write_pnm(pnm, O.case_image(name))
    out = str(tmp_path / "g1024.fco")
    ok, msg = hostlib.coder([pnm], out, quality=20.0, optimize=0)
    assert ok, msg
    assert md5(out) == O.manifest()[name]["fco_md5"] == "c5f96a1d79dcaeef269c40c668def7d9"


def test_unchanged_reference_cli_on_our_library(tmp_path):
    """cfiasco built from the reference's bin/*.c, linked against libfiasco.so (B200)."""
    exe = os.path.join(ROOT, "fiasco_b200", "lib", "cfiasco")
    assert os.path.exists(exe)
    name = "g256_q20_z0"
    pnm = str(tmp_path / "g256.pgm")
    gen_frames.write_pnm(pnm, O.case_image(name))
    out = str(tmp_path / "g256.fco")
    # small.fco must be findable for fiasco_c_options_set_basisfile (as with the reference)
    data = tmp_path / "data"
    data.mkdir()
    (data / "small.fco").write_text("Fiasco\n")
    env = dict(os.environ, FIASCO_DATA=str(data), FIASCO_IMAGES=str(tmp_path))
    r = subprocess.run([exe, "--progress-meter=0", "-V", "0", "-q", "20", "-i", pnm, "-o", out], env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert md5(out) == O.manifest()[name]["fco_md5"] == "3b393f57b00d4e6e64d1158c985fb5ba"
    # the reference decoder + PSNR tool agree with the numbers recorded from the reference run
    if os.path.exists(os.path.join(REF, "dfiasco")):
        dec = str(tmp_path / "dec.pgm")
        env["FIASCO_DATA"] = os.path.join(REF, "data")
        subprocess.run([os.path.join(REF, "dfiasco"), "-o", dec, out], env=env, check=True, capture_output=True)
        pr = subprocess.run([os.path.join(REF, "pnmpsnr"), pnm, dec], env=env, capture_output=True, text=True)
        psnr = [float(v) for v in re.findall(r"([0-9.]+) dB", pr.stdout + pr.stderr)]
        assert psnr == O.manifest()[name]["psnr_db"]


def test_intra_sequence_matches_reference_cli(tmp_path):
    """Three intra frames in one stream (pattern 'i'): all frames go to the device in one launch;
    bytes equal the reference CLI's stream when the reference binary is available on the box."""
    frames = [gen_frames.frame("g256"), gen_frames.frame("g256")[::-1].copy(), gen_frames.frame("g256").T.copy()]
    names = []
    for i, f in enumerate(frames):
        p = str(tmp_path / ("f%02d.pgm" % i))
        gen_frames.write_pnm(p, np.ascontiguousarray(f))
        names.append(p)
    out = str(tmp_path / "seq.fco")
    L = hostlib.load()
    o = hostlib.cli_options(0)
    L.fiasco_c_options_set_frame_pattern(o, b"i")
    ok, msg = hostlib.coder([str(tmp_path / "f0[0-2].pgm")], out, options=o)
    L.fiasco_c_options_delete(o)
    assert ok, msg
    cf = os.path.join(REF, "cfiasco")
    if os.path.exists(cf):
        ref_out = str(tmp_path / "ref.fco")
        env = dict(os.environ, FIASCO_DATA=os.path.join(REF, "data"), FIASCO_IMAGES=str(tmp_path))
        subprocess.run([cf, "--progress-meter=0", "-V", "0", "-q", "20", "--pattern=i", "-o", ref_out] + names,
                       env=env, check=True, capture_output=True)
        assert md5(out) == md5(ref_out)


def test_frames_without_reference_are_refused(tmp_path):
    """A B frame whose future reference is an I frame: the reference coder drops the past frame there
    (codec/coder.c:581-591) and reads through the NULL pointer; we refuse with a message."""
    p = str(tmp_path / "f.pgm")
    gen_frames.write_pnm(p, gen_frames.frame("g256")[:64, :64])
    L = hostlib.load()
    o = hostlib.cli_options(0)
    L.fiasco_c_options_set_frame_pattern(o, b"ibi")
    ok, msg = hostlib.coder([p, p, p], str(tmp_path / "o.fco"), options=o)
    L.fiasco_c_options_delete(o)
    assert not ok and "no reference frame" in msg
