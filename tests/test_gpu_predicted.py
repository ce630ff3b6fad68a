"""Motion path on the GPU (BASELINE config 5): P frames through the C ABI
(fb200_create_predicted / fb200_encode_predicted) and through fiasco_coder(), against the oracle and the
golden streams of the reference coder.  Bit exact: states, edges, weights, vectors, stream bytes."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import fiasco_b200 as F
from fiasco_b200 import ffi, hostlib
import oracle_lib as O
import gen_frames
from test_emu_device_code import (_holes_mode_automata, assert_same_predicted_automaton, check_b_frame_sequence,
                                  check_video_param_against_reference_library)

pytestmark = pytest.mark.gpu

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("name", ["v160_q20_ippp", "v352_q30_ippip"])
def test_gpu_predicted_frames_match_oracle(name):
    m, frames, ws, rec = _holes_mode_automata(name)
    p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
    enc = F.TileEncoder(p, 1, motion=F.Motion(1, 6, 10, 16))
    golden = list(O.golden_video_frames(name))
    try:
        for f in range(1, len(frames)):
            od = O.struct_dict(ws[f]["_struct"])
            if od["frame_type"] != 1:
                continue
            g = enc.encode_predicted([O.planes_of(frames[f])[0]], [rec[f - 1]])[0]
            assert_same_predicted_automaton(g, od)
            # holes closed by the host: the reference's own automaton (golden dump of the reference binary)
            g["frame_type"], g["frame_number"] = 1, f
            g["mv_bx"] = np.zeros_like(g["mv_fx"])
            g["mv_by"] = np.zeros_like(g["mv_fx"])
            g["delta_state"] = np.zeros(g["states"], np.uint8)
            done = hostlib.finish_predicted_frame(g)
            number, ftype, states, root, body = golden[f]
            assert (done["states"], done["root_state"]) == (states, root)
    finally:
        enc.close()


def test_gpu_predicted_frames_of_several_sequences_in_one_launch():
    """One thread block per sequence: the same P frame as three tiles of one launch and alone."""
    m, frames, ws, rec = _holes_mode_automata("v160_q20_ippp")
    p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
    enc = F.TileEncoder(p, 3, motion=F.Motion(1, 6, 10, 16))
    try:
        planes = [O.planes_of(frames[f])[0] for f in (1, 2, 3)]
        gs = enc.encode_predicted(planes, [rec[0], rec[1], rec[2]])
        for f, g in zip((1, 2, 3), gs):
            assert_same_predicted_automaton(g, O.struct_dict(ws[f]["_struct"]))
    finally:
        enc.close()


@pytest.mark.parametrize("p_min,p_max,sr", [(7, 9, 8), (6, 10, 5)])
def test_gpu_predicted_frames_other_levels_and_search_ranges(p_min, p_max, sr):
    frames = list(gen_frames.video(3, 176, 144))
    L = O.lib()
    L.fo_set_holes_mode(1)
    try:
        ws, rec = O.encode_video(frames, quality=20.0, pattern="ipp", p_min_level=p_min, p_max_level=p_max,
                                 search_range=sr)
    finally:
        L.fo_set_holes_mode(0)
    p = ffi.make_params(176, 144, 1, 20.0, 0)
    enc = F.TileEncoder(p, 1, motion=F.Motion(1, p_min, p_max, sr))
    try:
        for f in (1, 2):
            g = enc.encode_predicted([O.planes_of(frames[f])[0]], [rec[f - 1]])[0]
            assert_same_predicted_automaton(g, O.struct_dict(ws[f]["_struct"]))
    finally:
        enc.close()


def test_gpu_video_param_against_reference_library(tmp_path):
    check_video_param_against_reference_library(tmp_path)


def test_gpu_b_frames_match_oracle():
    check_b_frame_sequence()


@pytest.mark.parametrize("name", ["v160_q20_ippp", "v352_q30_ippip", "v160_q20_ibbp"])
def test_gpu_fiasco_coder_sequence_stream_is_byte_identical(name, tmp_path):
    m = O.manifest()[name]
    names = []
    for i, f in enumerate(gen_frames.video(m["frames"], m["width"], m["height"])):
        names.append(str(tmp_path / ("f%02d.pgm" % i)))
        gen_frames.write_pnm(names[-1], f)
    o = hostlib.cli_options(0)
    hostlib.load().fiasco_c_options_set_frame_pattern(o, m["pattern"].encode())
    out = str(tmp_path / "v.fco")
    ok, msg = hostlib.coder(names, out, float(m["quality"]), options=o)
    assert ok, msg
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == m["fco_md5"]


def test_default_pattern_sequence_matches_reference_cli(tmp_path):
    """The CLI's default frame pattern (ippppppppp): a short sequence through fiasco_coder() gives the
    reference CLI's bytes when the reference binary is available on the box."""
    names = []
    for i, f in enumerate(gen_frames.video(4, 176, 144)):
        names.append(str(tmp_path / ("f%02d.pgm" % i)))
        gen_frames.write_pnm(names[-1], f)
    out = str(tmp_path / "seq.fco")
    ok, msg = hostlib.coder(names, out)
    assert ok, msg
    cf = os.path.join(REF, "cfiasco")
    if os.path.exists(cf):
        ref_out = str(tmp_path / "ref.fco")
        env = dict(os.environ, FIASCO_DATA=os.path.join(REF, "data"), FIASCO_IMAGES=str(tmp_path))
        subprocess.run([cf, "--progress-meter=0", "-V", "0", "-q", "20", "-o", ref_out] + names,
                       env=env, check=True, capture_output=True)
        assert md5(out) == md5(ref_out)


# ------------------------------------------------- nondeterministic prediction (`cfiasco --prediction')

from test_emu_device_code import check_nd_coder_stream, check_nd_frames_against_oracle  # noqa: E402


@pytest.mark.parametrize("name", ["nd160_q70_i", "nd512_q80", "g256_q20_nd"])
def test_gpu_nd_prediction_frames_match_oracle(name):
    """fiasco_tile_kernel<NT, true> with frame type FB200_FRAME_ND: the third alternative of subdivide() is
    nd_prediction (codec/prediction.c:371).  Every frame, state for state, against the oracle."""
    assert check_nd_frames_against_oracle(name) > 0


@pytest.mark.parametrize("name", ["nd160_q70_i", "nd160_q70_ippp", "nd512_q80", "g256_q20_nd", "c128_q30_nd",
                                  "nd222_q60_zero_dc"])
def test_gpu_fiasco_coder_nd_prediction_streams(name, tmp_path):
    check_nd_coder_stream(name, tmp_path)


def test_gpu_cli_prediction_flag(tmp_path):
    """The unchanged reference CLI linked against our libfiasco: `cfiasco --prediction' writes the bytes the
    reference binary writes."""
    from test_emu_device_code import nd_case_frames
    m, frames = nd_case_frames("nd512_q80")
    pnm = str(tmp_path / "nd.pgm")
    gen_frames.write_pnm(pnm, frames[0])
    out = str(tmp_path / "nd.fco")
    cli = os.path.join(os.path.dirname(REF), "..", "fiasco_b200", "lib", "cfiasco")
    data = tmp_path / "data"
    data.mkdir()
    (data / "small.fco").write_text("Fiasco\n")
    env = dict(os.environ, FIASCO_IMAGES=str(tmp_path), FIASCO_DATA=str(data))
    r = subprocess.run([cli, "--progress-meter=0", "-V", "0", "-q", str(m["quality"]), "--prediction", "-o", out, pnm],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert md5(out) == m["fco_md5"]


# --------------------------------------------------------- colour sequences with predicted frames

from test_emu_device_code import check_colour_coder_stream, check_colour_predicted_frames_against_oracle  # noqa: E402


@pytest.mark.parametrize("name", ["cv160_q20_ippp", "cv160_q20_ibbpbbp", "cv352_q35_ipp"])
def test_gpu_colour_predicted_frames_match_oracle(name):
    assert check_colour_predicted_frames_against_oracle(name) >= 2


@pytest.mark.parametrize("name", ["cv160_q20_ippp", "cv160_q20_ibbpbbp", "cv160_q20_i", "cv160_q25_ippibp",
                                  "cv160_q20_ippp_nd", "cv352_q35_ipp"])
def test_gpu_fiasco_coder_colour_sequences(name, tmp_path):
    check_colour_coder_stream(name, tmp_path)


def test_gpu_fiasco_coder_colour_sequence_at_config5_size(tmp_path):
    """BASELINE config 5's geometry in colour: 30 frames 720x576, IPPP, q = 20 -- every frame chained to the
    one before it (range levels, y_column history); the reference's bytes (c5cbb7cf...)."""
    check_colour_coder_stream("cv720_q20_ippp", tmp_path)
    assert O.manifest()["cv720_q20_ippp"]["fco_md5"] == "c5cbb7cfb29d9a6368b039e9d41716c0"


def test_gpu_tile_split_of_a_colour_sequence(tmp_path):
    """FIASCO_TILE_SPLIT on a colour sequence with predicted frames: four independent streams, each chained
    through its own frames (range levels, y_column history), several colour frames per launch -- every tile's
    file equals what the reference binary writes for the cropped frames."""
    cf = os.path.join(REF, "cfiasco")
    if not os.path.exists(cf):
        pytest.skip("reference binary not on this box")
    frames = gen_frames.colour_video(3, 320, 256)
    names = []
    for i, f in enumerate(frames):
        names.append(str(tmp_path / ("s%02d.ppm" % i)))
        gen_frames.write_pnm(names[-1], f)
    L = hostlib.load()
    o = hostlib.cli_options(0)
    L.fiasco_c_options_set_frame_pattern(o, b"ipp")
    saved = os.environ.get("FIASCO_TILE_SPLIT")
    os.environ["FIASCO_TILE_SPLIT"] = "2"
    try:
        ok, msg = hostlib.coder(names, str(tmp_path / "sp.fco"), 20.0, options=o)
    finally:
        if saved is None:
            os.environ.pop("FIASCO_TILE_SPLIT", None)
        else:
            os.environ["FIASCO_TILE_SPLIT"] = saved
    assert ok, msg
    env = dict(os.environ, FIASCO_DATA=os.path.join(REF, "data"), FIASCO_IMAGES=str(tmp_path))
    for t in range(4):
        ty, tx = divmod(t, 2)
        crops = []
        for i, f in enumerate(frames):
            crops.append(str(tmp_path / ("c%d_%02d.ppm" % (t, i))))
            gen_frames.write_pnm(crops[-1], np.ascontiguousarray(f[ty * 128:(ty + 1) * 128, tx * 160:(tx + 1) * 160]))
        ref = str(tmp_path / ("ref%d.fco" % t))
        subprocess.run([cf, "--progress-meter=0", "-V", "0", "-q", "20", "--pattern=ipp", "-o", ref] + crops, env=env,
                       check=True, capture_output=True)
        assert md5(str(tmp_path / ("sp.t%02d.fco" % t))) == md5(ref), "tile %d" % t
