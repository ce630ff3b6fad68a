"""Test-side access to the oracle (CPU restatement, oracle/fiasco_oracle.c) and to the golden
fixtures generated from the reference (tests/golden/, see oracle/make_golden.py).

TEST INFRASTRUCTURE: nothing in fiasco_b200/ imports this.
"""
import ctypes as C
import gzip
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")
MAXSTATES = 6000

import sys
sys.path.insert(0, ORACLE_DIR)
import gen_frames  # noqa: E402


class FoParams(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("color", C.c_int), ("quality", C.c_float),
        ("lc_min_level", C.c_int), ("lc_max_level", C.c_int), ("images_level", C.c_int),
        ("max_elements", C.c_int), ("max_states", C.c_int), ("chroma_max_states", C.c_int),
        ("chroma_decrease", C.c_float), ("rpf_mantissa", C.c_int), ("rpf_range_e", C.c_int),
        ("dc_rpf_mantissa", C.c_int), ("dc_rpf_range_e", C.c_int),
        ("second_domain_block", C.c_int), ("check_for_underflow", C.c_int),
        ("check_for_overflow", C.c_int), ("full_search", C.c_int),
    ]


class FoWfa(C.Structure):
    _fields_ = [
        ("states", C.c_uint), ("basis_states", C.c_uint), ("root_state", C.c_uint), ("level", C.c_uint),
        ("final_distribution", C.c_float * MAXSTATES),
        ("level_of_state", C.c_uint8 * MAXSTATES),
        ("domain_type", C.c_uint8 * MAXSTATES),
        ("tree", C.c_int16 * 2 * MAXSTATES),
        ("x", C.c_uint16 * 2 * MAXSTATES),
        ("y", C.c_uint16 * 2 * MAXSTATES),
        ("into", C.c_int16 * 6 * 2 * MAXSTATES),
        ("weight", C.c_float * 6 * 2 * MAXSTATES),
        ("y_state", C.c_int16 * 2 * MAXSTATES),
        ("y_column", C.c_uint8 * 2 * MAXSTATES),
        ("costs", C.c_float * 3), ("err", C.c_float * 3), ("tree_bits", C.c_float * 3),
        ("matrix_bits", C.c_float * 3), ("weights_bits", C.c_float * 3),
        ("mv_type", C.c_int8 * 2 * MAXSTATES), ("mv_fx", C.c_int8 * 2 * MAXSTATES), ("mv_fy", C.c_int8 * 2 * MAXSTATES),
        ("mv_bx", C.c_int8 * 2 * MAXSTATES), ("mv_by", C.c_int8 * 2 * MAXSTATES),
        ("delta_state", C.c_uint8 * MAXSTATES), ("frame_type", C.c_int), ("frame_number", C.c_int),
    ]


class FoStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "subdivide_calls", "mp_calls", "pass1", "pass2", "ortho_steps", "accepted", "append_states",
        "leaf_dots", "ipss_lookups", "mp_domains", "ip_bytes", "blocks")]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
        src = os.path.join(ORACLE_DIR, "fiasco_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, capture_output=True)
        L = C.CDLL(so)
        L.fo_default_params.argtypes = [C.POINTER(FoParams), C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
        L.fo_default_params.restype = None
        L.fo_encode.argtypes = [C.POINTER(FoParams), C.POINTER(C.c_void_p), C.POINTER(FoWfa), C.POINTER(FoStats),
                                C.c_void_p, C.c_char_p, C.c_size_t]
        L.fo_rtob.argtypes = [C.c_float, C.c_uint, C.c_int]
        L.fo_btor.argtypes = [C.c_int, C.c_uint, C.c_int]
        L.fo_btor.restype = C.c_float
        L.fo_bits_bin_code.argtypes = [C.c_uint, C.c_uint]
        L.fo_bits_bin_code.restype = C.c_uint
        L.fo_neg_log2f.argtypes = [C.c_int, C.c_int]
        L.fo_neg_log2f.restype = C.c_float
        L.fo_image_level.argtypes = [C.c_uint, C.c_uint]
        L.fo_image_level.restype = C.c_uint
        L.fo_tree_model_kat.argtypes = [C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_float),
                                        C.POINTER(C.c_float)]
        L.fo_tree_model_kat.restype = None
        L.fo_decode_image.argtypes = [C.POINTER(FoWfa), C.c_int, C.c_uint, C.c_uint, C.POINTER(C.c_void_p)]
        L.fo_encode_video.argtypes = [C.POINTER(FoParams), C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_int, C.c_int,
                                      C.c_int, C.POINTER(FoWfa), C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t]
        L.fo_fill_norms_table.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint,
                                          C.c_uint, C.c_void_p]
        L.fo_fill_norms_table.restype = None
        L.fo_set_holes_mode.argtypes = [C.c_int]
        L.fo_set_holes_mode.restype = None
        L.fo_set_nd_prediction.argtypes = [C.c_int]
        L.fo_set_nd_prediction.restype = None
        L.fo_close_holes.argtypes = [C.POINTER(FoWfa)]
        L.fo_close_holes.restype = None
        L.fo_wfa_from_dump.argtypes = [C.c_char_p, C.c_uint, C.POINTER(FoWfa)]
        L.fo_restore_mc.argtypes = [C.POINTER(FoWfa), C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fo_restore_mc.restype = None
        L.fo_grey_to_plane.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.fo_rgb_to_planes.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p
        libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        L._libc = libc
        _LIB = L
    return _LIB


def planes_of(img):
    """u8 image (h,w) or (h,w,3) -> list of int16 planes in the coder's pixel format."""
    L = lib()
    img = np.ascontiguousarray(img, np.uint8)
    n = img.shape[0] * img.shape[1]
    if img.ndim == 2:
        p = np.zeros(n, np.int16)
        L.fo_grey_to_plane(img.ctypes.data, n, p.ctypes.data)
        return [p]
    y, cb, cr = (np.zeros(n, np.int16) for _ in range(3))
    L.fo_rgb_to_planes(img.ctypes.data, n, y.ctypes.data, cb.ctypes.data, cr.ctypes.data)
    return [y, cb, cr]


def default_params(width, height, color=0, quality=20.0, optimize=0):
    p = FoParams()
    lib().fo_default_params(C.byref(p), width, height, color, quality, optimize)
    return p


def encode(img, quality=20.0, optimize=0, want_trace=False, params=None):
    """Run the oracle on a u8 image.  Returns dict(wfa fields as numpy, stats, trace lines)."""
    L = lib()
    planes = planes_of(img)
    h, w = img.shape[:2]
    p = params or default_params(w, h, int(img.ndim == 3), quality, optimize)
    ptrs = (C.c_void_p * 3)(*([pl.ctypes.data for pl in planes] + [None] * (3 - len(planes))))
    wfa = FoWfa()
    st = FoStats()
    err = C.create_string_buffer(256)
    trace_lines = None
    fp = None
    path = None
    if want_trace:
        fd, path = tempfile.mkstemp(suffix=".trace")
        os.close(fd)
        fp = L._libc.fopen(path.encode(), b"w")
    rc = L.fo_encode(C.byref(p), ptrs, C.byref(wfa), C.byref(st), fp, err, 256)
    if fp:
        L._libc.fclose(fp)
        trace_lines = open(path).read().splitlines()
        os.unlink(path)
    if rc:
        raise RuntimeError("oracle: " + err.value.decode())
    n = wfa.states
    d = {
        "states": n, "basis_states": wfa.basis_states, "root_state": wfa.root_state, "level": wfa.level,
        "final_distribution": np.ctypeslib.as_array(wfa.final_distribution)[:n].copy(),
        "level_of_state": np.ctypeslib.as_array(wfa.level_of_state)[:n].copy(),
        "domain_type": np.ctypeslib.as_array(wfa.domain_type)[:n].copy(),
        "tree": np.ctypeslib.as_array(wfa.tree)[:n].copy(),
        "x": np.ctypeslib.as_array(wfa.x)[:n].copy(),
        "y": np.ctypeslib.as_array(wfa.y)[:n].copy(),
        "into": np.ctypeslib.as_array(wfa.into)[:n].copy(),
        "weight": np.ctypeslib.as_array(wfa.weight)[:n].copy(),
        "y_state": np.ctypeslib.as_array(wfa.y_state)[:n].copy(),
        "y_column": np.ctypeslib.as_array(wfa.y_column)[:n].copy(),
        "costs": list(wfa.costs), "err": list(wfa.err), "tree_bits": list(wfa.tree_bits),
        "matrix_bits": list(wfa.matrix_bits), "weights_bits": list(wfa.weights_bits),
        "stats": {k: getattr(st, k) for k, _ in FoStats._fields_},
        "trace": trace_lines,
        "_struct": wfa,                            # for decode()
        "_shape": (h, w, 3 if img.ndim == 3 else 1),
    }
    return d


def golden_video_frames(name):
    """[(frame number, frame type, states, root state, text of the frame's lines)] of a golden stream."""
    txt = gzip.open(os.path.join(GOLDEN, name + ".wfa.gz"), "rt").read()
    out = []
    for part in txt.split("frame ")[1:]:
        head, _, body = part.partition("\n")
        number, ftype, states, root = (int(v) for v in head.split())
        out.append((number, ftype, states, root, body))
    return out


def wfa_from_dump(text, root_state, width, height):
    """Automaton (dict with the ctypes struct, for decode()) from the canonical text of one frame."""
    w = FoWfa()
    if lib().fo_wfa_from_dump(text.encode(), root_state, C.byref(w)):
        raise RuntimeError("bad dump")
    return {"_struct": w, "_shape": (height, width, 1), "states": w.states}


def restore_mc(w, image, past, half_pixel=0, future=None):
    """restore_mc (codec/motion.c:37) on a regenerated grey frame, in place."""
    h, wd, _ = w["_shape"]
    past = np.ascontiguousarray(past, np.int16)
    future = np.ascontiguousarray(future, np.int16) if future is not None else None
    lib().fo_restore_mc(C.byref(w["_struct"]), wd, h, half_pixel, image.ctypes.data, past.ctypes.data,
                        future.ctypes.data if future is not None else None)
    return image


def decode(w):
    """Regenerate the frame of an oracle automaton (fo_decode_image): list of int16 (h, w) planes in
    the coder's pixel format."""
    h, wd, bands = w["_shape"]
    planes = [np.zeros((h, wd), np.int16) for _ in range(bands)]
    ptrs = (C.c_void_p * 3)(*([p.ctypes.data for p in planes] + [None] * (3 - bands)))
    rc = lib().fo_decode_image(C.byref(w["_struct"]), int(bands == 3), wd, h, ptrs)
    if rc:
        raise RuntimeError("oracle decode failed")
    return planes


def encode_video(frames, quality=20.0, pattern="ippp", p_min_level=6, p_max_level=10, search_range=16, trace_path=None):
    """Run the oracle on a sequence (list of u8 (h, w) grey or (h, w, 3) RGB frames; CLI defaults of cfiasco).
    Returns (list of automata as ctypes structs wrapped like wfa_from_dump(), regenerated frames int16 [n][h][w],
    colour: [n][3][h][w])."""
    L = lib()
    h, w = frames[0].shape[:2]
    colour = frames[0].ndim == 3
    planes = [pl for f in frames for pl in planes_of(f)]
    ptrs = (C.c_void_p * len(planes))(*[pl.ctypes.data for pl in planes])
    p = default_params(w, h, int(colour), quality, 0)
    out = (FoWfa * len(frames))()
    rec = np.zeros((len(frames), 3, h, w) if colour else (len(frames), h, w), np.int16)
    err = C.create_string_buffer(256)
    fp = L._libc.fopen(trace_path.encode(), b"w") if trace_path else None
    rc = L.fo_encode_video(C.byref(p), len(frames), ptrs, pattern.encode(), p_min_level, p_max_level, search_range,
                           out, rec.ctypes.data, fp, err, 256)
    if fp:
        L._libc.fclose(fp)
    if rc:
        raise RuntimeError("oracle: " + err.value.decode())
    return [{"_struct": out[i], "_shape": (h, w, 3 if colour else 1), "states": out[i].states}
            for i in range(len(frames))], rec


def struct_dict(st):
    """numpy view of a ctypes automaton (the fields the stream writer takes)."""
    n = st.states
    d = {"states": n, "basis_states": st.basis_states, "root_state": st.root_state, "frame_type": st.frame_type,
         "frame_number": st.frame_number}
    for name in ("final_distribution", "level_of_state", "domain_type", "tree", "x", "y", "into", "weight",
                 "y_state", "y_column", "mv_type", "mv_fx", "mv_fy", "mv_bx", "mv_by", "delta_state"):
        d[name] = np.ctypeslib.as_array(getattr(st, name))[:n].copy()
    return d


def struct_lines(st):
    """Canonical lines ('s', 'm', 'd', 'e' in the order of oracle/wfadump.c) of a ctypes automaton."""
    out = []
    for s in range(st.basis_states, st.states):
        out.append("s %d %d %d %d %d %d %d %d 0 0" % (s, st.level_of_state[s], st.tree[s][0], st.tree[s][1],
                                                   st.x[s][0], st.y[s][0], st.x[s][1], st.y[s][1]))
        for label in range(2):
            if st.mv_type[s][label]:
                out.append("m %d %d %d %d %d %d %d" % (s, label, st.mv_type[s][label], st.mv_fx[s][label],
                                                     st.mv_fy[s][label], st.mv_bx[s][label], st.mv_by[s][label]))
        if st.delta_state[s]:
            out.append("d %d" % s)
        for label in range(2):
            for e in range(6):
                t = st.into[s][label][e]
                if t < 0:
                    break
                wgt = np.float32(st.weight[s][label][e])
                out.append("e %d %d %d %08x %.9g" % (s, label, t, int(wgt.view(np.uint32)), float(wgt)))
    return out


def wfa_lines(w):
    """Canonical 's' / 'e' lines (grammar of oracle/wfadump.c) of an oracle automaton."""
    out = []
    fb = w["weight"].view(np.uint32)
    for s in range(w["basis_states"], w["states"]):
        out.append("s %d %d %d %d %d %d %d %d 0 0" % (
            s, int(w["level_of_state"][s]), int(w["tree"][s][0]), int(w["tree"][s][1]),
            int(w["x"][s][0]), int(w["y"][s][0]), int(w["x"][s][1]), int(w["y"][s][1])))
        for label in range(2):
            for e in range(6):
                t = int(w["into"][s][label][e])
                if t < 0:
                    break
                out.append("e %d %d %d %08x %.9g" % (s, label, t, int(fb[s][label][e]), float(w["weight"][s][label][e])))
    return out


# ---------------------------------------------------------------- golden fixtures

def manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def golden_wfa_lines(name, image_level=None):
    """'s'/'e' lines of the reference's own WFA dump.  For states above the image level (the
    virtual colour states) the reader invents coordinates; mask them the way we emit them."""
    txt = gzip.open(os.path.join(GOLDEN, name + ".wfa.gz"), "rt").read().splitlines()
    info = [l for l in txt if l.startswith("info ")][0].split()
    level = int(info[3])
    out = []
    for l in txt:
        if l.startswith("s "):
            f = l.split()
            if int(f[2]) > level:
                f[5:9] = ["0", "0", "0", "0"]
            out.append(" ".join(f))
        elif l.startswith("e "):
            out.append(l)
    return out


def mask_virtual(lines, level):
    out = []
    for l in lines:
        if l.startswith("s "):
            f = l.split()
            if int(f[2]) > level:
                f[5:9] = ["0", "0", "0", "0"]
            l = " ".join(f)
        out.append(l)
    return out


def golden_trace(name):
    return gzip.open(os.path.join(GOLDEN, name + ".trace.gz"), "rt").read().splitlines()


def golden_kat():
    return gzip.open(os.path.join(GOLDEN, "kat.txt.gz"), "rt").read().splitlines()


_FRAMES = {}


def case_image(name):
    """The u8 input image of a golden case (regenerated from the seeded generator)."""
    m = manifest()[name]
    fr = m["frame"]
    if fr not in _FRAMES:
        _FRAMES[fr] = gen_frames.frame(fr)
    img = _FRAMES[fr]
    crop = m["crop"]
    if crop is not None:
        if crop[0] == "rect":
            _, x0, y0, w, h = crop
            img = img[y0:y0 + h, x0:x0 + w].copy()
        else:
            img = gen_frames.crops(img, crop[0])[crop[1]]
    import hashlib
    assert hashlib.md5(gen_frames.pnm_bytes(img)).hexdigest() == m["pnm_md5"], "frame generator drifted: " + name
    return img


def lc_lines(trace_lines):
    return [l for l in trace_lines if l.startswith("lc ")]
