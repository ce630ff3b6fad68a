"""Host side of libfiasco (product code, CPU): the .fco writer must be byte-identical to the
reference's.  Input automata come from the oracle (itself pinned to the reference), the expected
md5 of the stream from the reference binary (tests/golden/manifest.json).  Also checks that the
public API of include/fiasco.h / fiasco_host.h is exported and behaves like the reference's."""
import ctypes as C
import hashlib
import os
import re

import numpy as np

import pytest

from fiasco_b200 import ffi, hostlib
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["g256_q20_z0", "g256_q20_z1", "g256_q20_z2", "g512_q20_z0", "g1024t0_q20_z0",
                                  "g1024t15_q20_z0", "g1024r_q20_z0", "g1024s_q40_z0", "g4096t0_q20_z0",
                                  "c256_q20_z0", "c256_q30_z0", "c2048t0_q30_z0"])
def test_writer_is_byte_identical_to_reference(name, tmp_path):
    m = O.manifest()[name]
    img = O.case_image(name)
    w = O.encode(img, quality=m["quality"], optimize=m["optimize"])
    p = ffi.make_params(m["width"], m["height"], 3 if m["color"] else 1, float(m["quality"]), m["optimize"])
    out = str(tmp_path / (name + ".fco"))
    hostlib.write_stream(out, p, [w])
    data = open(out, "rb").read()
    assert len(data) == m["fco_bytes"]
    assert hashlib.md5(data).hexdigest() == m["fco_md5"]


def test_full_frame_stream_1024(tmp_path):
    name = "g1024_q20_z0"
    m = O.manifest()[name]
    w = O.encode(O.case_image(name), quality=20, optimize=0)
    out = str(tmp_path / "g1024.fco")
    hostlib.write_stream(out, ffi.make_params(1024, 1024, 1, 20.0, 0), [w])
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == m["fco_md5"]   # c5f96a1d...


def test_public_api_exports():
    L = hostlib.load()
    for header in ("fiasco.h", "fiasco_host.h"):
        txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", header)).read(), flags=re.S)
        names = set(re.findall(r"\b(fiasco_[a-z0-9_]+)\s*\(", txt))
        assert names, header
        for n in names:
            assert hasattr(L, n), "missing export " + n
    # the two extra symbols the reference CLI links
    assert hasattr(L, "fiasco_calloc") and hasattr(L, "open_file") and hasattr(L, "fiasco_free")


def test_option_setters_follow_reference_contract():
    L = hostlib.load()
    o = L.fiasco_c_options_new()
    assert L.fiasco_c_options_set_optimizations(o, 6, 10, 3, 10000, 0) == 1
    assert L.fiasco_c_options_set_optimizations(o, 3, 10, 3, 10000, 0) == 0
    assert "at least level 4" in hostlib.error_message()
    assert L.fiasco_c_options_set_optimizations(o, 8, 6, 3, 10000, 0) == 0
    assert L.fiasco_c_options_set_quantization(o, 9, 2, 5, 1) == 0
    assert "[2,8]" in hostlib.error_message()
    assert L.fiasco_c_options_set_frame_pattern(o, b"ixp") == 0
    assert "invalid character `x'" in hostlib.error_message()
    assert L.fiasco_c_options_set_frame_pattern(o, b"IbP") == 1
    assert L.fiasco_c_options_set_smoothing(o, 101) == 0
    assert L.fiasco_c_options_set_chroma_quality(o, 0.0, 40) == 0
    assert L.fiasco_c_options_set_prediction(o, 0, 5, 10) == 0
    assert L.fiasco_c_options_set_tiling(o, 7, 4) == 0
    L.fiasco_c_options_delete(o)


def test_coder_errors_return_zero_with_message(tmp_path):
    ok, msg = hostlib.coder([str(tmp_path / "missing.pgm")], str(tmp_path / "o.fco"))
    assert not ok and "missing.pgm" in msg
    ok, msg = hostlib.coder([str(tmp_path / "missing.pgm")], str(tmp_path / "o.fco"), quality=0.0)
    assert not ok and "positive" in msg


def test_coder_refuses_without_gpu(tmp_path):
    """fiasco_coder() has no CPU fallback: on a box without a CUDA device it fails loudly."""
    import fiasco_b200 as F
    if F.device_count() > 0:
        pytest.skip("a CUDA device is present")
    import gen_frames
    p = str(tmp_path / "g.pgm")
    gen_frames.write_pnm(p, gen_frames.frame("g256")[:64, :64])
    ok, msg = hostlib.coder([p], str(tmp_path / "o.fco"))
    assert not ok and "CUDA" in msg


def test_cfiasco_links_unchanged():
    """The reference CLI (compiled unchanged from /root/reference/bin) links against our library."""
    exe = os.path.join(ROOT, "fiasco_b200", "lib", "cfiasco")
    assert os.path.exists(exe), "run `make product` where /root/reference is available"
    import subprocess
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libfiasco.so" in out and "libfiasco_b200.so" in out
    sy = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for s in ("fiasco_coder", "fiasco_c_options_new", "fiasco_c_options_set_optimizations", "fiasco_calloc", "open_file"):
        assert s in sy


@pytest.mark.parametrize("name", ["v160_q20_ippp", "v352_q30_ippip", "v160_q20_ibbp"])
def test_video_stream_is_byte_identical_to_reference(name, tmp_path):
    """Host half of the motion path: fiasco_write_video_stream() (frame types, the motion tree and
    vectors of output/mc.c, delta contexts of output/weights.c) writes the reference coder's bytes
    for sequences with predicted frames.  The automata come from the test oracle here -- the GPU
    path does not produce predicted frames yet."""
    import hashlib
    m = O.manifest()[name]
    frames = list(O.gen_frames.video(m["frames"], m["width"], m["height"]))
    ws, _ = O.encode_video(frames, quality=m["quality"], pattern=m["pattern"])
    p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
    out = str(tmp_path / "v.fco")
    hostlib.write_video_stream(out, p, [O.struct_dict(w["_struct"]) for w in ws])
    b = open(out, "rb").read()
    assert len(b) == m["fco_bytes"]
    assert hashlib.md5(b).hexdigest() == m["fco_md5"]


def test_host_regenerates_frames_like_the_reference():
    """fiasco_regenerate_frame() (host side of the motion path: decode_image + restore_mc of the
    reference, codec/decoder.c:412, codec/motion.c:37): intra and predicted frames of the golden
    sequence and a ragged still, from the oracle's automata, equal the frames the reference coder
    regenerated (md5 of the shorts, written by oracle/decdump.c)."""
    m = O.manifest()["v160_q20_ippp"]
    frames = list(O.gen_frames.video(m["frames"], m["width"], m["height"]))
    ws, _ = O.encode_video(frames, quality=m["quality"], pattern=m["pattern"])
    past = None
    for f, w in enumerate(ws):
        img = hostlib.regenerate_frame(O.struct_dict(w["_struct"]), m["width"], m["height"], past)
        assert hashlib.md5(img.tobytes()).hexdigest() == m["decoded_md5"][f], "frame %d" % f
        past = img
    for name in ("g1024r_q20_z0", "g256_q20_z0"):
        m = O.manifest()[name]
        w = O.encode(O.case_image(name), quality=m["quality"], optimize=m["optimize"])
        img = hostlib.regenerate_frame(O.struct_dict(w["_struct"]), m["width"], m["height"])
        assert hashlib.md5(img.tobytes()).hexdigest() == m["decoded_md5"], name


def test_host_finishes_predicted_frames_with_holes():
    """The automaton of a predicted frame as the device will leave it (states of losing split
    alternatives left as holes, delta flags not set): fiasco_finish_predicted_frame() closes the
    holes and derives the flags; the stream written from the result is the reference's, byte for byte."""
    name = "v352_q30_ippip"
    m = O.manifest()[name]
    frames = list(O.gen_frames.video(m["frames"], m["width"], m["height"]))
    L = O.lib()
    L.fo_set_holes_mode(1)
    try:
        ws, _ = O.encode_video(frames, quality=m["quality"], pattern=m["pattern"])
    finally:
        L.fo_set_holes_mode(0)
    done, holes = [], 0
    for w in ws:
        d = O.struct_dict(w["_struct"])
        d["delta_state"] = np.zeros_like(d["delta_state"])          # the host has to find them itself
        f = hostlib.finish_predicted_frame(d)
        holes += d["states"] - f["states"]
        done.append(f)
    assert holes > 100
    p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "v.fco")
        hostlib.write_video_stream(out, p, done)
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == m["fco_md5"]


def test_host_regenerates_b_frames_like_the_reference():
    """B frames: forward, backward and interpolated motion compensation against the previous and the
    future regenerated frame, frames in coding order, B frames serving as past references
    (codec/coder.c:571-627): every regenerated frame of the golden IBBP sequence equals the reference's."""
    m = O.manifest()["v160_q20_ibbp"]
    frames = list(O.gen_frames.video(m["frames"], m["width"], m["height"]))
    ws, _ = O.encode_video(frames, quality=m["quality"], pattern=m["pattern"])
    past = future = reconst = None
    future_frame, expected, seen = False, 0, set()
    assert [w["_struct"].frame_number for w in ws] == [0, 3, 1, 2, 4, 6, 5]
    for k, w in enumerate(ws):
        d = O.struct_dict(w["_struct"])
        if d["frame_type"] == 0:
            past = future = reconst = None
        elif d["frame_type"] == 1:
            past, future, reconst = reconst, None, None
        elif future_frame:
            future, reconst = reconst, None
        else:
            past, reconst = reconst, None
        seen.add(d["frame_number"])
        future_frame = d["frame_number"] > expected
        while expected in seen:
            expected += 1
        reconst = hostlib.regenerate_frame(d, m["width"], m["height"], past, future)
        assert hashlib.md5(reconst.tobytes()).hexdigest() == m["decoded_md5"][k], "coded frame %d" % k


@pytest.mark.parametrize("name", ["nd160_q70_i", "nd160_q70_ippp", "nd512_q80"])
def test_nd_prediction_stream_is_byte_identical_to_reference(name, tmp_path):
    """Streams coded with `--prediction': the nondeterminism tree and its DC weights (output/nd.c) in every
    frame, the delta contexts of the states below an ND-predicted range -- from the oracle's automata the
    writer gives the bytes of the reference cfiasco."""
    import hashlib
    from test_emu_device_code import nd_case_frames
    m, frames = nd_case_frames(name)
    L = O.lib()
    L.fo_set_nd_prediction(1)
    try:
        ws, _ = O.encode_video(frames, quality=m["quality"], pattern=m["pattern"])
    finally:
        L.fo_set_nd_prediction(0)
    p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
    out = str(tmp_path / "nd.fco")
    hostlib.write_video_stream(out, p, [O.struct_dict(w["_struct"]) for w in ws], nd_prediction=True)
    b = open(out, "rb").read()
    assert (len(b), hashlib.md5(b).hexdigest()) == (m["fco_bytes"], m["fco_md5"])


@pytest.mark.parametrize("name", ["cv160_q20_ippp", "cv160_q20_ibbpbbp"])
def test_colour_video_stream_and_regenerated_frames(name, tmp_path):
    """Host side of colour sequences with predicted frames, from the oracle's automata: the stream writer
    gives the reference's bytes (the y_column entries of the virtual states included, which the reference
    writes without setting them), and fiasco_regenerate_colour_frame() -- three bands, the luminance
    tree's vectors, chroma clipping -- the frames the oracle regenerates."""
    import hashlib
    from test_emu_device_code import colour_case_frames
    m, frames = colour_case_frames(name)
    ws, rec = O.encode_video(frames, quality=m["quality"], pattern=m["pattern"])
    p = ffi.make_params(m["width"], m["height"], 3, float(m["quality"]), 0)
    out = str(tmp_path / "cv.fco")
    hostlib.write_video_stream(out, p, [O.struct_dict(w["_struct"]) for w in ws])
    b = open(out, "rb").read()
    assert (len(b), hashlib.md5(b).hexdigest()) == (m["fco_bytes"], m["fco_md5"])
    past = future = reconst = None
    future_frame, expected, seen = False, 0, set()
    for k, w in enumerate(ws):
        d = O.struct_dict(w["_struct"])
        if d["frame_type"] == 0:
            past = future = reconst = None
        elif d["frame_type"] == 1:
            past, future, reconst = reconst, None, None
        elif future_frame:
            future, reconst = reconst, None
        else:
            past, reconst = reconst, None
        seen.add(d["frame_number"])
        future_frame = d["frame_number"] > expected
        while expected in seen:
            expected += 1
        reconst = hostlib.regenerate_frame(d, m["width"], m["height"], past, future, colour=True)
        assert np.array_equal(reconst, rec[k]), "coded frame %d" % k
