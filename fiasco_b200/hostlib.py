"""ctypes binding of the host library libfiasco.so (include/fiasco.h, include/fiasco_host.h)."""
import ctypes as C
import os

import numpy as np

from . import ffi

_LIB = None


class StreamInfo(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("color", C.c_int),
        ("max_states", C.c_uint), ("chroma_max_states", C.c_uint),
        ("p_min_level", C.c_uint), ("p_max_level", C.c_uint), ("smoothing", C.c_uint), ("fps", C.c_uint),
        ("rpf_mantissa", C.c_int), ("rpf_range_e", C.c_int), ("dc_rpf_mantissa", C.c_int), ("dc_rpf_range_e", C.c_int),
        ("title", C.c_char_p), ("comment", C.c_char_p), ("nd_prediction", C.c_int),
    ]


class FrameMotion(C.Structure):
    _fields_ = [("frame_type", C.c_int), ("frame_number", C.c_int), ("mv_type", C.c_void_p), ("mv_fx", C.c_void_p),
                ("mv_fy", C.c_void_p), ("mv_bx", C.c_void_p), ("mv_by", C.c_void_p), ("delta_state", C.c_void_p)]


_MOTION_ARRAYS = (("mv_type", np.int8), ("mv_fx", np.int8), ("mv_fy", np.int8), ("mv_bx", np.int8), ("mv_by", np.int8),
                  ("delta_state", np.uint8))


def lib_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libfiasco.so")


def load():
    global _LIB
    if _LIB is None:
        ffi.load()                                 # libfiasco.so depends on libfiasco_b200.so
        L = C.CDLL(lib_path())
        L.fiasco_get_error_message.restype = C.c_char_p
        L.fiasco_stream_info_init.argtypes = [C.POINTER(StreamInfo), C.POINTER(ffi.Params)]
        L.fiasco_stream_info_init.restype = None
        L.fiasco_write_stream.argtypes = [C.c_char_p, C.POINTER(StreamInfo), C.POINTER(ffi._Wfa), C.c_int]
        L.fiasco_write_video_stream.argtypes = [C.c_char_p, C.POINTER(StreamInfo), C.POINTER(ffi._Wfa),
                                                C.POINTER(FrameMotion), C.c_int, C.c_uint]
        L.fiasco_regenerate_frame.argtypes = [C.POINTER(ffi._Wfa), C.POINTER(FrameMotion), C.c_int, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p]
        L.fiasco_regenerate_colour_frame.argtypes = L.fiasco_regenerate_frame.argtypes
        L.fiasco_finish_predicted_frame.argtypes = [C.POINTER(ffi._Wfa)] + [C.c_void_p] * 6
        L.fiasco_coder.argtypes = [C.POINTER(C.c_char_p), C.c_char_p, C.c_float, C.c_void_p]
        L.fiasco_c_options_new.restype = C.c_void_p
        L.fiasco_c_options_delete.argtypes = [C.c_void_p]
        L.fiasco_c_options_set_optimizations.argtypes = [C.c_void_p] + [C.c_uint] * 5
        L.fiasco_c_options_set_quantization.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_uint, C.c_int]
        L.fiasco_c_options_set_frame_pattern.argtypes = [C.c_void_p, C.c_char_p]
        L.fiasco_c_options_set_title.argtypes = [C.c_void_p, C.c_char_p]
        L.fiasco_c_options_set_video_param.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int]
        L.fiasco_c_options_set_chroma_quality.argtypes = [C.c_void_p, C.c_float, C.c_uint]
        L.fiasco_c_options_set_prediction.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint]
        L.fiasco_c_options_set_smoothing.argtypes = [C.c_void_p, C.c_int]
        L.fiasco_c_options_set_tiling.argtypes = [C.c_void_p, C.c_int, C.c_uint]
        L.fiasco_c_options_set_progress_meter.argtypes = [C.c_void_p, C.c_int]
        L.fiasco_c_options_set_basisfile.argtypes = [C.c_void_p, C.c_char_p]
        L.fiasco_set_verbosity.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def error_message():
    return load().fiasco_get_error_message().decode(errors="replace")


def wfa_struct(w):
    """Build a fb200_wfa_t from a dict of numpy arrays (as returned by TileEncoder / the test oracle)."""
    s = ffi._Wfa()
    keep = {}
    n = int(w["states"])
    s.capacity, s.status, s.states = n, 0, n
    s.basis_states, s.root_state = int(w["basis_states"]), int(w["root_state"])
    for name, dt in (("final_distribution", np.float32), ("level_of_state", np.uint8), ("domain_type", np.uint8),
                     ("tree", np.int16), ("x", np.uint16), ("y", np.uint16), ("into", np.int16),
                     ("weight", np.float32), ("y_state", np.int16), ("y_column", np.uint8)):
        a = np.ascontiguousarray(w[name], dtype=dt)
        keep[name] = a
        setattr(s, name, a.ctypes.data)
    return s, keep


def write_stream(path, params, wfas, title=None, comment=None):
    """Serialise automata (one per intra frame) into a .fco file."""
    L = load()
    info = StreamInfo()
    L.fiasco_stream_info_init(C.byref(info), C.byref(params))
    info.title = title.encode() if title else None
    info.comment = comment.encode() if comment else None
    arr = (ffi._Wfa * len(wfas))()
    keep = []
    for i, w in enumerate(wfas):
        arr[i], k = wfa_struct(w)
        keep.append(k)
    if not L.fiasco_write_stream(path.encode(), C.byref(info), arr, len(wfas)):
        raise RuntimeError("fiasco_write_stream: " + error_message())


def write_video_stream(path, params, wfas, p_min_level=6, p_max_level=10, search_range=16, nd_prediction=False):
    """Serialise a sequence with predicted frames: every automaton dict also holds "frame_type" and, for
    predicted frames, "mv_type", "mv_fx", "mv_fy" ([states][2] int8) and "delta_state" ([states] uint8)."""
    L = load()
    info = StreamInfo()
    L.fiasco_stream_info_init(C.byref(info), C.byref(params))
    info.p_min_level, info.p_max_level = p_min_level, p_max_level
    info.nd_prediction = int(nd_prediction)
    arr = (ffi._Wfa * len(wfas))()
    mot = (FrameMotion * len(wfas))()
    keep = []
    for i, w in enumerate(wfas):
        arr[i], k = wfa_struct(w)
        keep.append(k)
        mot[i].frame_type = int(w.get("frame_type", 0))
        mot[i].frame_number = int(w.get("frame_number", i))
        if mot[i].frame_type:
            for name, dt in _MOTION_ARRAYS:
                a = np.ascontiguousarray(w[name], dtype=dt)
                k[name] = a
                setattr(mot[i], name, a.ctypes.data)
        elif nd_prediction:                        # the delta states of an intra frame with ND prediction
            a = np.ascontiguousarray(w["delta_state"], dtype=np.uint8)
            k["delta_state"] = a
            mot[i].delta_state = a.ctypes.data
    if not L.fiasco_write_video_stream(path.encode(), C.byref(info), arr, mot, len(wfas), search_range):
        raise RuntimeError("fiasco_write_video_stream: " + error_message())


def regenerate_frame(w, width, height, past=None, future=None, colour=False):
    """The frame an automaton describes, int16 (h, w) -- colour: (3, h, w) -- in the coder's pixel format
    (fiasco_regenerate_frame / fiasco_regenerate_colour_frame); predicted frames ("frame_type" 1, 2) need the
    regenerated reference frame(s)."""
    L = load()
    s, keep = wfa_struct(w)
    mot = FrameMotion()
    mot.frame_type = int(w.get("frame_type", 0))
    if mot.frame_type:
        for name, dt in _MOTION_ARRAYS:
            a = np.ascontiguousarray(w[name], dtype=dt)
            keep[name] = a
            setattr(mot, name, a.ctypes.data)
        past = np.ascontiguousarray(past, np.int16)
        future = np.ascontiguousarray(future, np.int16) if future is not None else None
    out = np.zeros((3, height, width) if colour else (height, width), np.int16)
    fn = L.fiasco_regenerate_colour_frame if colour else L.fiasco_regenerate_frame
    if not fn(C.byref(s), C.byref(mot), width, height,
                                     past.ctypes.data if past is not None else None,
                                     future.ctypes.data if future is not None else None, out.ctypes.data):
        raise RuntimeError("fiasco_regenerate_frame: " + error_message())
    return out


def finish_predicted_frame(w):
    """fiasco_finish_predicted_frame() on an automaton dict with holes (level_of_state 255): returns a new
    dict with the holes closed and the delta flags derived from the structure."""
    L = load()
    w = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in w.items()}
    s, keep = wfa_struct(w)
    extra = {name: np.ascontiguousarray(w[name], dtype=dt) for name, dt in _MOTION_ARRAYS}
    n = L.fiasco_finish_predicted_frame(C.byref(s), *[extra[name].ctypes.data for name, _ in _MOTION_ARRAYS])
    if not n:
        raise RuntimeError("fiasco_finish_predicted_frame: " + error_message())
    out = {"states": n, "basis_states": s.basis_states, "root_state": s.root_state, "frame_type": w.get("frame_type", 0),
           "frame_number": w.get("frame_number", 0)}
    for name, a in list(keep.items()) + list(extra.items()):
        out[name] = a[:n]
    return out


def cli_options(optimize=0):
    """The option object the reference CLI builds for its default parameters (bin/cwfa.c:253-388)."""
    L = load()
    o = L.fiasco_c_options_new()
    L.fiasco_c_options_set_frame_pattern(o, b"ippppppppp")
    L.fiasco_c_options_set_chroma_quality(o, 2.0, 40)
    L.fiasco_c_options_set_smoothing(o, 70)
    L.fiasco_c_options_set_progress_meter(o, 0)
    L.fiasco_c_options_set_tiling(o, 3, 4)
    if optimize <= 0:
        L.fiasco_c_options_set_optimizations(o, 6, 10, 3, 10000, 0)
    else:
        L.fiasco_c_options_set_optimizations(o, 4, 12, 5, 10000, optimize - 1)
    L.fiasco_c_options_set_prediction(o, 0, 6, 10)
    L.fiasco_c_options_set_quantization(o, 3, 2, 5, 1)
    return o


def coder(inputs, output, quality=20.0, optimize=0, options=None):
    """fiasco_coder() with the CLI's default options; returns (ok, error message)."""
    L = load()
    own = options is None
    o = options or cli_options(optimize)
    names = (C.c_char_p * (len(inputs) + 1))(*([s.encode() for s in inputs] + [None]))
    ok = L.fiasco_coder(names, output.encode(), quality, o)
    if own:
        L.fiasco_c_options_delete(o)
    return bool(ok), error_message()
