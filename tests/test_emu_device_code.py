"""The device sources run on the CPU (tests/emu: one fibre per CUDA thread) against the oracle.

TEST INFRASTRUCTURE ONLY.  fiasco_b200/csrc/*.cu -- the files nvcc compiles for sm_100a -- are
compiled here as plain C++ against tests/emu/cuda_runtime.h and driven through the same C ABI,
so that the control flow of the kernels (DFS, pursuit, state appends, warp-level resolution) is
exercised in the CPU-only suite, where there is no GPU.  The product library is not involved
and keeps failing with FB200_ENODEVICE without a device (tests/test_abi.py); the parity tests
proper are the `-m gpu` ones.  What this cannot show: anything that depends on the real
hardware (memory model, cp.async ordering, divergence inside a warp, nvcc's code generation)."""
import os
import platform
import subprocess

import numpy as np
import pytest

import fiasco_b200 as F
from fiasco_b200 import ffi
import oracle_lib as O
import gen_frames
import test_gpu_parity as T

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libfiasco_b200_emu.so")

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's context switch is x86-64")


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
    saved = (ffi._LIB, ffi.lib_path, os.environ.get("FB200_NT"))
    ffi._LIB = None
    ffi.lib_path = lambda: EMU_LIB
    os.environ["FB200_NT"] = "128"
    try:
        ffi.load()
        yield ffi
    finally:
        ffi._LIB, ffi.lib_path = saved[0], saved[1]
        if saved[2] is None:
            os.environ.pop("FB200_NT", None)
        else:
            os.environ["FB200_NT"] = saved[2]


def test_emulated_device_code_grey_tile(emu):
    img = gen_frames.chan(256, 256, 1)
    gw, _, st = T.gpu_encode(img)
    T.assert_same_wfa(gw, O.encode(img))
    assert st["mp_calls"] > 1000


@pytest.mark.parametrize("nt", ["96", "512"])
def test_emulated_device_code_other_block_shapes(emu, nt):
    os.environ["FB200_NT"] = nt
    try:
        img = gen_frames.chan(128, 96, 5)
        T.assert_same_wfa(T.gpu_encode(img)[0], O.encode(img))
    finally:
        os.environ["FB200_NT"] = "128"


@pytest.mark.parametrize("cluster", ["2", "8"])
def test_emulated_cluster_per_stream(emu, cluster):
    """The clustered shape (one stream on several thread blocks: speculated pursuits of a range's label-0
    descendants, products and table levels shared out) under the emulator, which runs the blocks of a
    cluster together: grey, colour and -z 2 equal the oracle, work counters included."""
    saved = os.environ.get("FB200_CLUSTER")
    os.environ["FB200_NT"], os.environ["FB200_CLUSTER"] = "512", cluster
    try:
        img = gen_frames.chan(128, 96, 5)
        gw, _, st = T.gpu_encode(img)
        ow = O.encode(img, want_trace=True)
        T.assert_same_wfa(gw, ow)
        assert st["mp_calls"] == len(O.lc_lines(ow["trace"]))
        col = np.stack([gen_frames.chan(64, 64, s) for s in (11, 12, 13)], axis=-1)
        T.assert_same_wfa(T.gpu_encode(col, quality=30.0)[0], O.encode(col, quality=30.0), bands=3)
        grey = gen_frames.chan(96, 64, 2)
        T.assert_same_wfa(T.gpu_encode(grey, optimize=2)[0], O.encode(grey, optimize=2))
    finally:
        os.environ["FB200_NT"] = "128"
        if saved is None:
            os.environ.pop("FB200_CLUSTER", None)
        else:
            os.environ["FB200_CLUSTER"] = saved


def test_emulated_device_code_colour_and_optimisation_levels(emu):
    img = np.stack([gen_frames.chan(128, 128, s) for s in (11, 12, 13)], axis=-1)
    T.assert_same_wfa(T.gpu_encode(img, quality=30.0)[0], O.encode(img, quality=30.0), bands=3)
    grey = gen_frames.chan(96, 128, 2)
    for z in (1, 2):
        T.assert_same_wfa(T.gpu_encode(grey, optimize=z)[0], O.encode(grey, optimize=z))


def test_emulated_motion_norms(emu):
    rng = np.random.default_rng(5)
    orig = ((rng.integers(0, 256, (64, 96)).astype(np.int16) - 128) * 16).copy()
    past = (np.roll(orig, (2, -3), (0, 1)) + rng.integers(-40, 40, orig.shape).astype(np.int16)).copy()
    got, _ = ffi.motion_norms(orig, past, 6, 16)
    L = O.lib()
    ref = np.zeros(1024, np.float32)
    for by in range(got.shape[0]):
        for bx in range(got.shape[1]):
            L.fo_fill_norms_table(orig.ctypes.data, past.ctypes.data, 96, 64, bx * 8, by * 8, 6, 16, ref.ctypes.data)
            assert np.array_equal(got[by, bx].view(np.uint32), ref.view(np.uint32)), (bx, by)


# ------------------------------------------------------------------ predicted frames (motion path)

def _holes_mode_automata(name):
    m = O.manifest()[name]
    frames = list(gen_frames.video(m["frames"], m["width"], m["height"]))
    L = O.lib()
    L.fo_set_holes_mode(1)
    try:
        ws, rec = O.encode_video(frames, quality=m["quality"], pattern=m["pattern"])
    finally:
        L.fo_set_holes_mode(0)
    return m, frames, ws, rec


def assert_same_predicted_automaton(g, od):
    """Device automaton of a predicted frame (holes still open) against the oracle's in holes mode."""
    n = od["states"]
    assert g["status"] == 0 and g["states"] == n and g["root_state"] == od["root_state"]
    assert np.array_equal(g["level_of_state"][:n], od["level_of_state"][:n])
    live = od["level_of_state"][:n] != 255
    for k in ("tree", "x", "y", "mv_type", "mv_fx", "mv_fy", "domain_type"):
        assert np.array_equal(g[k][:n][live], od[k][:n][live]), k
    assert np.array_equal(g["final_distribution"][:n][live].view(np.uint32),
                          od["final_distribution"][:n][live].view(np.uint32))
    for s in np.nonzero(live)[0]:
        if s < 3:
            continue
        for label in range(2):
            for e in range(6):
                assert g["into"][s][label][e] == od["into"][s][label][e], (s, label, e)
                if od["into"][s][label][e] < 0:
                    break
                assert g["weight"][s][label][e].view(np.uint32) == od["weight"][s][label][e].view(np.uint32)


@pytest.mark.parametrize("name", ["v160_q20_ippp", "v352_q30_ippip"])
def test_emulated_device_code_predicted_frames(emu, name):
    """P frames: the kernel's third alternative (motion search over the norms tables, nested pass over
    the prediction error with the delta models, holes scheme) state for state against the oracle run in
    holes mode, each frame predicted from the oracle's regenerated previous frame."""
    m, frames, ws, rec = _holes_mode_automata(name)
    p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
    enc = F.TileEncoder(p, 1, motion=F.Motion(1, 6, 10, 16))
    checked = 0
    try:
        for f in range(1, len(frames)):
            od = O.struct_dict(ws[f]["_struct"])
            if od["frame_type"] != 1:
                continue
            g = enc.encode_predicted([O.planes_of(frames[f])[0]], [rec[f - 1]])[0]
            assert_same_predicted_automaton(g, od)
            checked += 1
    finally:
        enc.close()
    assert checked >= 3


def test_emulated_cluster_per_stream_predicted_frames(emu):
    """P frames on a cluster of thread blocks: the helper blocks follow rank 0 into the nested pass over a
    prediction error (the other product table, the delta models, the error block's pixels) -- every frame of
    the golden IPPP sequence state for state against the oracle in holes mode."""
    saved = os.environ.get("FB200_CLUSTER")
    os.environ["FB200_NT"], os.environ["FB200_CLUSTER"] = "512", "2"
    try:
        m, frames, ws, rec = _holes_mode_automata("v160_q20_ippp")
        p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
        enc = F.TileEncoder(p, 1, motion=F.Motion(1, 6, 10, 16))
        try:
            for f in range(1, len(frames)):
                od = O.struct_dict(ws[f]["_struct"])
                g = enc.encode_predicted([O.planes_of(frames[f])[0]], [rec[f - 1]])[0]
                assert_same_predicted_automaton(g, od)
        finally:
            enc.close()
    finally:
        os.environ["FB200_NT"] = "128"
        if saved is None:
            os.environ.pop("FB200_CLUSTER", None)
        else:
            os.environ["FB200_CLUSTER"] = saved


def predicted_frames_in_coding_order(name):
    """(frame type, display number, plane, past, future, oracle automaton in holes mode) of every predicted
    frame of a golden sequence, with the reference bookkeeping of video_coder() (codec/coder.c:571-627)
    and the oracle's regenerated frames as references."""
    m, frames, ws, rec = _holes_mode_automata(name)
    past = future = reconst = None
    future_frame, expected, seen, out = False, 0, set(), []
    for k, w in enumerate(ws):
        d = O.struct_dict(w["_struct"])
        if d["frame_type"] == 0:
            past = future = None
        elif d["frame_type"] == 1:
            past, future = reconst, None
        elif future_frame:
            future = reconst
        else:
            past = reconst
        seen.add(d["frame_number"])
        future_frame = d["frame_number"] > expected
        while expected in seen:
            expected += 1
        if d["frame_type"]:
            out.append((d["frame_type"], d["frame_number"], O.planes_of(frames[d["frame_number"]])[0], past, future, d))
        reconst = rec[k]
    return m, out


def check_b_frame_sequence(name="v160_q20_ibbp"):
    m, todo = predicted_frames_in_coding_order(name)
    p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
    encs = {t: F.TileEncoder(p, 1, motion=F.Motion(t, 6, 10, 16)) for t in (1, 2)}
    used = np.zeros(4, int)
    try:
        for t, number, plane, past, future, d in todo:
            g = encs[t].encode_predicted([plane], [past], [future] if t == 2 else None)[0]
            assert_same_predicted_automaton(g, d)
            n = d["states"]
            live = d["level_of_state"][:n] != 255
            assert np.array_equal(g["mv_bx"][:n][live], d["mv_bx"][:n][live])
            assert np.array_equal(g["mv_by"][:n][live], d["mv_by"][:n][live])
            if t == 2:
                used += np.bincount(d["mv_type"][:n][live].ravel(), minlength=4)
    finally:
        for e_ in encs.values():
            e_.close()
    assert used[1] and used[2] and used[3]          # forward, backward and interpolated ranges occurred


def test_emulated_device_code_b_frames(emu):
    """B frames (find_B_frame_mc: best forward vector, best backward vector, both together; backward
    norms tables; interpolated prediction error), coded out of display order, B frames as past
    references: every predicted frame of the golden IBBP sequence against the oracle in holes mode."""
    check_b_frame_sequence()


@pytest.mark.parametrize("name", ["v160_q20_ippp", "v352_q30_ippip", "v160_q20_ibbp"])
def test_emulated_fiasco_coder_writes_the_reference_stream_for_sequences(emu, name, tmp_path):
    """fiasco_coder() on a sequence with predicted frames -- I frames in one launch, the P frames group by
    group along their chains (sequences with B frames: frame by frame in coding order), holes closed and
    frames regenerated on the host between the steps -- writes the stream the reference cfiasco writes,
    byte for byte (md5 of the golden .fco)."""
    import hashlib
    from fiasco_b200 import hostlib
    saved = (hostlib._LIB, hostlib.lib_path)
    hostlib._LIB, hostlib.lib_path = None, (lambda: os.path.join(EMU_DIR, "_build", "libfiasco_emu.so"))
    try:
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
    finally:
        hostlib._LIB, hostlib.lib_path = saved


@pytest.mark.parametrize("order", ["1", "2"])
def test_emulated_results_do_not_depend_on_thread_scheduling(emu, order):
    """The emulator runs the threads of a block between two rendezvous points in ascending order by
    default; descending (1) and pseudo-random (2) orders must give the same automata -- a result that
    depended on the order would be a race between barriers on the GPU."""
    os.environ["FB200_EMU_ORDER"] = order
    try:
        img = gen_frames.chan(128, 128, 4)
        T.assert_same_wfa(T.gpu_encode(img)[0], O.encode(img))
        m, frames, ws, rec = _holes_mode_automata("v160_q20_ippp")
        p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
        enc = F.TileEncoder(p, 1, motion=F.Motion(1, 6, 10, 16))
        try:
            g = enc.encode_predicted([O.planes_of(frames[1])[0]], [rec[0]])[0]
            assert_same_predicted_automaton(g, O.struct_dict(ws[1]["_struct"]))
        finally:
            enc.close()
    finally:
        os.environ.pop("FB200_EMU_ORDER", None)


def test_emulated_python_group_loop_equals_fiasco_coder(emu, tmp_path):
    """fiasco_b200/video.py (groups of pictures for several GPUs) against the C frame loop of
    fiasco_coder() on groups of two frames each (pattern "ip": adjacent groups, each reference regenerated
    from an I frame)."""
    import hashlib
    from fiasco_b200 import hostlib, video
    saved = (hostlib._LIB, hostlib.lib_path)
    hostlib._LIB, hostlib.lib_path = None, (lambda: os.path.join(EMU_DIR, "_build", "libfiasco_emu.so"))
    try:
        frames = list(gen_frames.video(4, 176, 144))
        names = []
        for i, f in enumerate(frames):
            names.append(str(tmp_path / ("f%02d.pgm" % i)))
            gen_frames.write_pnm(names[-1], f)
        o = hostlib.cli_options(0)
        hostlib.load().fiasco_c_options_set_frame_pattern(o, b"ip")
        ok, msg = hostlib.coder(names, str(tmp_path / "c.fco"), 20.0, options=o)
        assert ok, msg
        p = ffi.make_params(176, 144, 1, 20.0, 0)
        seq, _ = video.encode_sequence([ffi.pixels_from_grey(f) for f in frames], "ip", p)
        hostlib.write_video_stream(str(tmp_path / "p.fco"), p, seq)
        md5 = lambda n: hashlib.md5(open(str(tmp_path / n), "rb").read()).hexdigest()
        assert md5("p.fco") == md5("c.fco")
    finally:
        hostlib._LIB, hostlib.lib_path = saved


def test_emulated_random_sequences_match_the_reference_binary(emu):
    """A few cases of tools/fuzz_video_emu.py: random short sequences with P and B frames through
    fiasco_coder() (device code under the emulator) against the unmodified reference binary of
    oracle/_ref, byte for byte.  Skipped where the reference binary was not built."""
    import subprocess
    import sys
    root = os.path.dirname(HERE)
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "cfiasco")):
        pytest.skip("oracle/_ref/cfiasco not built")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_video_emu.py"), "8", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "8 cases, 0 mismatches" in r.stdout


@pytest.mark.parametrize("p_min,p_max,sr", [(7, 9, 8), (6, 10, 5), (8, 8, 16)])
def test_emulated_predicted_frames_other_levels_and_search_ranges(emu, p_min, p_max, sr):
    """Prediction levels and search ranges other than the CLI's 6..10 / 16 (c_options_t.p_min_level,
    p_max_level, search_range): the tables hold 4 sr^2 vectors, levels outside [p_min, p_max] are not
    predicted."""
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


def check_video_param_against_reference_library(tmp_path):
    """fiasco_c_options_set_video_param (frames per second, B frames as past references or not): the
    reference CLI parses these flags but never passes them on, so the comparison goes through
    oracle/_ref/refcoder (our harness around the unmodified reference LIBRARY).  Byte for byte."""
    import hashlib
    import subprocess
    from fiasco_b200 import hostlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rc = os.path.join(root, "oracle", "_ref", "refcoder")
    if not os.path.exists(rc):
        pytest.skip("oracle/_ref/refcoder not built")
    L = hostlib.load()
    names = []
    for i, f in enumerate(gen_frames.video(7, 176, 144)):
        names.append(str(tmp_path / ("f%02d.pgm" % i)))
        gen_frames.write_pnm(names[-1], f)
    env = dict(os.environ, FIASCO_DATA=os.path.join(root, "oracle", "_ref", "data"), FIASCO_IMAGES=str(tmp_path))
    for pattern, fps, b_as_past in (("ibbp", 25, 0), ("ibbbp", 12, 0), ("ibbp", 30, 1)):
        o = hostlib.cli_options(0)
        L.fiasco_c_options_set_frame_pattern(o, pattern.encode())
        assert L.fiasco_c_options_set_video_param(o, fps, 0, 0, b_as_past)
        out, ref = str(tmp_path / "ours.fco"), str(tmp_path / "ref.fco")
        ok, msg = hostlib.coder(names, out, 20.0, options=o)
        L.fiasco_c_options_delete(o)
        assert ok, msg
        subprocess.run([rc, ref, "20", pattern, str(fps), "0", "0", str(b_as_past)] + names, env=env, check=True,
                       capture_output=True)
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == hashlib.md5(open(ref, "rb").read()).hexdigest(), \
            (pattern, fps, b_as_past)


def test_emulated_video_param_against_reference_library(emu, tmp_path):
    from fiasco_b200 import hostlib
    saved = (hostlib._LIB, hostlib.lib_path)
    hostlib._LIB, hostlib.lib_path = None, (lambda: os.path.join(EMU_DIR, "_build", "libfiasco_emu.so"))
    try:
        check_video_param_against_reference_library(tmp_path)
    finally:
        hostlib._LIB, hostlib.lib_path = saved


# ------------------------------------------------- nondeterministic prediction (`cfiasco --prediction')

def nd_case_frames(name):
    m = O.manifest()[name]
    if name.startswith("nd160"):
        frames = gen_frames.nd_sequence()
    elif name == "nd512_q80":
        frames = [gen_frames.nd_still()]
    elif name == "g256_q20_nd":
        frames = [gen_frames.frame("g256")]
    elif "fixture" in m:                           # a frame kept as a file (found by the fuzzer)
        import gzip
        raw = gzip.open(os.path.join(O.GOLDEN, m["fixture"])).read()
        frames = [np.frombuffer(raw[raw.index(b"255\n") + 4:], np.uint8).reshape(m["height"], m["width"]).copy()]
    else:
        frames = [gen_frames.colour_sequence(2, 128, 128)[1]]
    assert len(frames) == m["frames"]
    return m, frames


def check_nd_frames_against_oracle(name, which=None):
    """Intra frames with nondeterministic prediction through the C ABI (a context of frame type
    FB200_FRAME_ND): the device's automaton -- holes still open -- state for state against the oracle
    run in holes mode (restatement of nd_prediction, codec/prediction.c:371)."""
    m, frames = nd_case_frames(name)
    L = O.lib()
    L.fo_set_holes_mode(1)
    L.fo_set_nd_prediction(1)
    try:
        ws, _ = O.encode_video(frames, quality=m["quality"], pattern="i")
    finally:
        L.fo_set_holes_mode(0)
        L.fo_set_nd_prediction(0)
    p = ffi.make_params(m["width"], m["height"], 1, float(m["quality"]), 0)
    enc = F.TileEncoder(p, 1, motion=F.Motion(3, 6, 10, 16))
    predicted = 0
    try:
        for f in (range(len(frames)) if which is None else which):
            od = O.struct_dict(ws[f]["_struct"])
            g = enc.encode_predicted([O.planes_of(frames[f])[0]], None)[0]
            assert_same_predicted_automaton(g, od)
            n = od["states"]
            predicted += int(((od["tree"][:n] >= 0) & (od["into"][:n, :, 0] >= 0)).sum())
    finally:
        enc.close()
    return predicted


def check_nd_coder_stream(name, tmp_path):
    """fiasco_coder() with fiasco_c_options_set_prediction (intra_prediction = YES): the bytes the
    reference `cfiasco --prediction' writes (md5 of the golden stream)."""
    import hashlib
    from fiasco_b200 import hostlib
    m, frames = nd_case_frames(name)
    names = []
    for i, f in enumerate(frames):
        names.append(str(tmp_path / ("n%02d.%s" % (i, "pgm" if f.ndim == 2 else "ppm"))))
        gen_frames.write_pnm(names[-1], f)
    o = hostlib.cli_options(0)
    L = hostlib.load()
    L.fiasco_c_options_set_frame_pattern(o, m["pattern"].encode())
    L.fiasco_c_options_set_prediction(o, 1, 6, 10)
    out = str(tmp_path / "nd.fco")
    ok, msg = hostlib.coder(names, out, float(m["quality"]), options=o)
    assert ok, msg
    b = open(out, "rb").read()
    assert (len(b), hashlib.md5(b).hexdigest()) == (m["fco_bytes"], m["fco_md5"])


def test_emulated_device_code_nd_prediction(emu):
    assert check_nd_frames_against_oracle("nd160_q70_i", which=(0, 2)) > 5
    # DC weights that round to zero cost infinitely many bits in the reference (it reads in front of its
    # table of counts, codec/coeff.c:237): never taken
    assert check_nd_frames_against_oracle("nd222_q60_zero_dc") >= 0


def test_emulated_fiasco_coder_nd_prediction_streams(emu, tmp_path):
    """Under the emulator: an I-only sequence with ND prediction, and the IPPP form of the same frames,
    two of whose frames end off a byte boundary (no edges at all)."""
    from fiasco_b200 import hostlib
    saved = (hostlib._LIB, hostlib.lib_path)
    hostlib._LIB, hostlib.lib_path = None, (lambda: os.path.join(EMU_DIR, "_build", "libfiasco_emu.so"))
    try:
        for name in ("nd160_q70_ippp", "c128_q30_nd", "nd222_q60_zero_dc"):
            check_nd_coder_stream(name, tmp_path)
    finally:
        hostlib._LIB, hostlib.lib_path = saved


# --------------------------------------------------------- colour sequences with predicted frames

def colour_case_frames(name):
    m = O.manifest()[name]
    frames = gen_frames.colour_video(m["frames"], m["width"], m["height"])
    return m, frames


def check_colour_predicted_frames_against_oracle(name, which=None):
    """Colour P / B frames through the C ABI: luminance with prediction, the luminance tree's motion
    compensation taken off the chroma planes on the device (subtract_mc, codec/mwfa.c:156), chroma bands on
    the same workspace -- state for state against the oracle run in holes mode, every frame predicted from
    the oracle's regenerated reference frames, starting with the range level the frame before left."""
    m, frames = colour_case_frames(name)
    L = O.lib()
    L.fo_set_holes_mode(1)
    try:
        ws, rec = O.encode_video(frames, quality=m["quality"], pattern=m["pattern"])
    finally:
        L.fo_set_holes_mode(0)
    p = ffi.make_params(m["width"], m["height"], 3, float(m["quality"]), 0)
    coded = {w["_struct"].frame_number: k for k, w in enumerate(ws)}

    def lc_min_after(st):
        d = O.struct_dict(st)
        y_root = d["tree"][d["tree"][d["root_state"]][0]][0]
        levels = [int(d["level_of_state"][s]) - 1 for s in range(3, y_root + 1)
                  if d["level_of_state"][s] != 255 and (d["tree"][s] < 0).any()]
        return min(levels)

    encs, checked = {}, 0
    past = future = reconst = None
    future_frame, expected, seen = False, 0, set()
    try:
        for k, w in enumerate(ws):
            od = O.struct_dict(w["_struct"])
            t = od["frame_type"]
            if t == 0:
                past = future = reconst = None
            elif t == 1:
                past, future, reconst = reconst, None, None
            elif future_frame:
                future, reconst = reconst, None
            else:
                past, reconst = reconst, None
            seen.add(od["frame_number"])
            future_frame = od["frame_number"] > expected
            while expected in seen:
                expected += 1
            reconst = rec[k]
            if t == 0 or (which is not None and k not in which):
                continue
            if t not in encs:
                encs[t] = F.TileEncoder(p, 1, motion=F.Motion(t, 6, 10, 16))
            g = encs[t].encode_predicted(O.planes_of(frames[od["frame_number"]]), [past],
                                         [future] if t == 2 else None, lc_min=[lc_min_after(ws[k - 1]["_struct"])])[0]
            assert_same_predicted_automaton(g, od)
            checked += 1
    finally:
        for e in encs.values():
            e.close()
    assert coded and checked
    return checked


def check_colour_coder_stream(name, tmp_path):
    """fiasco_coder() on a colour sequence with predicted frames: the bytes of the reference cfiasco."""
    import hashlib
    from fiasco_b200 import hostlib
    m, frames = colour_case_frames(name)
    names = []
    for i, f in enumerate(frames):
        names.append(str(tmp_path / ("c%02d.ppm" % i)))
        gen_frames.write_pnm(names[-1], f)
    o = hostlib.cli_options(0)
    L = hostlib.load()
    L.fiasco_c_options_set_frame_pattern(o, m["pattern"].encode())
    L.fiasco_c_options_set_prediction(o, int(m["nd_prediction"]), 6, 10)
    out = str(tmp_path / "cv.fco")
    ok, msg = hostlib.coder(names, out, float(m["quality"]), options=o)
    assert ok, msg
    b = open(out, "rb").read()
    assert (len(b), hashlib.md5(b).hexdigest()) == (m["fco_bytes"], m["fco_md5"])


def test_emulated_device_code_colour_predicted_frame(emu):
    assert check_colour_predicted_frames_against_oracle("cv160_q20_ippp", which=(1,)) == 1


def test_emulated_fiasco_coder_colour_sequence_with_predicted_frames(emu, tmp_path):
    from fiasco_b200 import hostlib
    saved = (hostlib._LIB, hostlib.lib_path)
    hostlib._LIB, hostlib.lib_path = None, (lambda: os.path.join(EMU_DIR, "_build", "libfiasco_emu.so"))
    try:
        check_colour_coder_stream("cv160_q20_ippp", tmp_path)
    finally:
        hostlib._LIB, hostlib.lib_path = saved
