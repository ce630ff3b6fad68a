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


def test_half_pixel_vectors_are_refused(tmp_path):
    """Half-pixel motion compensation (and the cross-B search the reference ties to it, codec/coder.c:359): the
    reference takes the vector as `unsigned' and halves it (extract_mc_block, codec/motion.c:232-260), so a
    negative vector reads far outside the frame -- undefined there, refused here with a message; the same
    sequence with full-pixel vectors is coded."""
    names = []
    for i, f in enumerate(gen_frames.video(2, 160, 128)):
        names.append(str(tmp_path / ("h%d.pgm" % i)))
        gen_frames.write_pnm(names[-1], f)
    L = hostlib.load()
    o = hostlib.cli_options(0)
    L.fiasco_c_options_set_frame_pattern(o, b"ip")
    L.fiasco_c_options_set_video_param(o, 25, 1, 0, 1)
    ok, msg = hostlib.coder(names, str(tmp_path / "o.fco"), options=o)
    assert not ok and "Half pixel" in msg
    L.fiasco_c_options_set_video_param(o, 25, 0, 0, 1)
    ok, msg = hostlib.coder(names, str(tmp_path / "o.fco"), options=o)
    L.fiasco_c_options_delete(o)
    assert ok, msg


# ---------------------------------------------------------------------------------------------
# round 2: colour sequences, the big frame at -z 1 / -z 2, tile-split mode and several GPUs inside
# the library, every tile of BASELINE configs 2, 3, 4 and config 5 at its full size
# ---------------------------------------------------------------------------------------------

def _with_env(**kv):
    saved = {k: os.environ.get(k) for k in kv}
    for k, v in kv.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    return saved


def _restore_env(saved):
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("name", ["cseq128_sd_q25_z0", "cseq128_sd_q25_z1", "cseq128_sd_q40_z1", "cseq128_ds_q25_z0",
                                  "cseq128_ds_q40_z1"])
def test_colour_sequence_of_intra_frames_hands_lc_min_level_on(name, tmp_path):
    """codec/coder.c:797: the chroma set-up of a frame raises c->options.lc_min_level and nothing sets
    it back, so frame k + 1 of a colour sequence starts where frame k ended.  Smooth frame first
    ("sd"): the second frame's bytes differ from coding it alone; the reference's stream is the pin."""
    m = O.manifest()[name]
    seq = gen_frames.colour_sequence(2, m["width"], m["height"])
    names = []
    for i, k in enumerate(m["order"]):
        names.append(str(tmp_path / ("f%d.ppm" % i)))
        gen_frames.write_pnm(names[-1], seq[k])
    L = hostlib.load()
    o = hostlib.cli_options(m["optimize"])
    L.fiasco_c_options_set_frame_pattern(o, b"i")
    out = str(tmp_path / "seq.fco")
    ok, msg = hostlib.coder(names, out, quality=float(m["quality"]), options=o)
    L.fiasco_c_options_delete(o)
    assert ok, msg
    assert md5(out) == m["fco_md5"]


def test_frame_the_reference_refuses_is_refused_alike(tmp_path):
    """A very smooth colour frame: the reference's writer stops with "Can't write more than N weights."
    (output/weights.c:137); so do we, with the same text."""
    m = O.manifest()["csmooth128_q25_refused"]
    p = str(tmp_path / "s.ppm")
    gen_frames.write_pnm(p, gen_frames.colour_sequence(1, 128, 128, 12)[0])
    ok, msg = hostlib.coder([p], str(tmp_path / "s.fco"), quality=float(m["quality"]))
    assert not ok and msg == m["message"]


@pytest.mark.parametrize("z", [1, 2])
def test_fiasco_coder_1024_higher_optimisation_levels(z, tmp_path):
    m = O.manifest()["g1024_q20_z%d" % z]
    pnm = str(tmp_path / "g1024.pgm")
    gen_frames.write_pnm(pnm, gen_frames.frame("g1024"))
    out = str(tmp_path / "g1024.fco")
    ok, msg = hostlib.coder([pnm], out, quality=20.0, optimize=z)
    assert ok, msg
    assert md5(out) == m["fco_md5"]


def _tile_split_case(key, split, gpus, tmp_path):
    m = O.manifest()[key]
    img = gen_frames.frame(m["frame"])
    pnm = str(tmp_path / ("img" + (".pgm" if img.ndim == 2 else ".ppm")))
    gen_frames.write_pnm(pnm, img)
    out = str(tmp_path / "out.fco")
    saved = _with_env(FIASCO_TILE_SPLIT=split, FIASCO_GPUS=gpus)
    try:
        ok, msg = hostlib.coder([pnm], out, quality=float(m["quality"]), optimize=m["optimize"])
    finally:
        _restore_env(saved)
    assert ok, msg
    n = 1 << split
    got = [md5(str(tmp_path / ("out.t%02d.fco" % t))) for t in range(n)]
    bad = [t for t in range(n) if got[t] != m["fco_md5"][t]]
    assert not bad, "tiles %s differ from the reference coder run on the crops" % bad
    assert not os.path.exists(out)


def test_tile_split_config2_all_16_tiles(tmp_path):
    """BASELINE config 2 in its tile-split form (FIASCO_TILE_SPLIT=4): every one of the 16 streams has the
    bytes the reference writes for the 256^2 crop."""
    _tile_split_case("tiles_g1024_256", 4, 1, tmp_path)


def test_tile_split_config3_all_64_colour_tiles(tmp_path):
    """BASELINE config 3: 2048^2 colour, q = 30, 64 streams."""
    _tile_split_case("tiles_c2048_256", 6, 1, tmp_path)


def test_tile_split_config4_all_64_tiles(tmp_path):
    """BASELINE config 4: 4096^2 grey, q = 20, 64 streams of 512^2."""
    _tile_split_case("tiles_g4096_512", 6, 1, tmp_path)


def test_tile_split_over_all_gpus_of_the_box(tmp_path):
    """FIASCO_GPUS: the tiles are dealt to the devices, one host thread each; same bytes.  (On a box
    with one GPU the library clamps to it and this repeats the single-device case.)"""
    import fiasco_b200 as F
    _tile_split_case("tiles_g1024_256", 4, max(1, min(8, F.device_count())), tmp_path)


def test_config5_full_size_stream_md5(tmp_path):
    """BASELINE config 5 at its own size: 30 frames 720x576, IPPP, q = 20 (e9d88f99...)."""
    m = O.manifest()["v720_q20_ippp"]
    for i, f in enumerate(gen_frames.video(m["frames"], m["width"], m["height"])):
        gen_frames.write_pnm(str(tmp_path / ("w%02d.pgm" % i)), f)
    L = hostlib.load()
    o = hostlib.cli_options(0)
    L.fiasco_c_options_set_frame_pattern(o, m["pattern"].encode())
    out = str(tmp_path / "v.fco")
    saved = _with_env(FIASCO_GPUS=8)
    try:
        ok, msg = hostlib.coder([str(tmp_path / "w[00-29].pgm")], out, quality=20.0, options=o)
    finally:
        _restore_env(saved)
        L.fiasco_c_options_delete(o)
    assert ok, msg
    assert md5(out) == m["fco_md5"] == "e9d88f99690abf5b88c449478ff1dbf3"


def test_progress_meter_output(tmp_path):
    """The meter of subdivide() (codec/subdivide.c:323-349) as the reference CLI shows it: the bar is 50
    marks and a newline per band, the percent counter the values the traversal passes."""
    exe = os.path.join(ROOT, "fiasco_b200", "lib", "cfiasco")
    cf = os.path.join(REF, "cfiasco")
    if not (os.path.exists(exe) and os.path.exists(cf)):
        pytest.skip("needs both command line binaries")
    pnm = str(tmp_path / "g.pgm")
    gen_frames.write_pnm(pnm, gen_frames.frame("g1024")[:136, :200].copy())
    data = tmp_path / "data"
    data.mkdir()
    (data / "small.fco").write_text("Fiasco\n")
    for meter in ("1", "2"):
        outs = []
        for binary, datadir in ((exe, str(data)), (cf, os.path.join(REF, "data"))):
            env = dict(os.environ, FIASCO_DATA=datadir, FIASCO_IMAGES=str(tmp_path))
            r = subprocess.run([binary, "--progress-meter=" + meter, "-V", "1", "-q", "20", "-i", pnm, "-o",
                                str(tmp_path / "o.fco")], env=env, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            outs.append([ln for ln in r.stderr.replace("\r", "\n").split("\n")
                         if ln.strip() and "resource file" not in ln and "params.c" not in ln])
        assert outs[0] == outs[1], meter
