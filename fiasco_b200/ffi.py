"""ctypes binding of include/fiasco_b200.h (tests / bench only; no computation here)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAXEDGES = 5
OK, EINVAL, ENODEVICE, ECUDA, ECAPACITY, EMAXSTATES, ENOROOT, EUNSUPPORTED = range(8)


class FB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fiasco_b200 error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("bands", C.c_int), ("level", C.c_int),
        ("lc_min_level", C.c_int), ("lc_max_level", C.c_int), ("images_level", C.c_int),
        ("max_elements", C.c_int), ("max_states", C.c_int), ("chroma_max_states", C.c_int),
        ("price", C.c_float), ("chroma_decrease", C.c_float),
        ("rpf_mantissa", C.c_int), ("rpf_range", C.c_float),
        ("dc_rpf_mantissa", C.c_int), ("dc_rpf_range", C.c_float),
        ("second_domain_block", C.c_int), ("state_capacity", C.c_int),
    ]


class _Wfa(C.Structure):
    _fields_ = [
        ("capacity", C.c_int), ("status", C.c_int),
        ("states", C.c_uint), ("basis_states", C.c_uint), ("root_state", C.c_uint),
        ("costs", C.c_float * 3), ("err", C.c_float * 3), ("tree_bits", C.c_float * 3),
        ("matrix_bits", C.c_float * 3), ("weights_bits", C.c_float * 3),
        ("final_distribution", C.c_void_p), ("level_of_state", C.c_void_p),
        ("domain_type", C.c_void_p), ("tree", C.c_void_p), ("x", C.c_void_p), ("y", C.c_void_p),
        ("into", C.c_void_p), ("weight", C.c_void_p), ("y_state", C.c_void_p), ("y_column", C.c_void_p),
        ("mv_type", C.c_void_p), ("mv_fx", C.c_void_p), ("mv_fy", C.c_void_p),
        ("mv_bx", C.c_void_p), ("mv_by", C.c_void_p),
        ("lc_min_level", C.c_int), ("progress", (C.c_uint32 * 4) * 3), ("y_column_history", C.c_void_p),
    ]


class Motion(C.Structure):
    _fields_ = [("frame_type", C.c_int), ("p_min_level", C.c_int), ("p_max_level", C.c_int),
                ("search_range", C.c_int)]


class TraceRec(C.Structure):
    _fields_ = [
        ("level", C.c_uint16), ("image", C.c_uint16), ("address", C.c_uint16),
        ("x", C.c_uint16), ("y", C.c_uint16), ("y_state", C.c_int16),
        ("states", C.c_uint16), ("n_edges", C.c_int16),
        ("max_costs", C.c_float), ("price", C.c_float), ("costs", C.c_float), ("err", C.c_float),
        ("matrix_bits", C.c_float), ("weights_bits", C.c_float),
        ("into", C.c_int16 * 6), ("weight", C.c_float * 6),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("ip_bytes", C.c_uint64),
        ("mp_calls", C.c_uint64), ("mp_steps", C.c_uint64), ("pass2", C.c_uint64),
        ("blocks", C.c_uint64), ("states", C.c_uint64), ("mp_bytes", C.c_uint64), ("ss_bytes", C.c_uint64),
        ("cyc_total", C.c_uint64), ("cyc_T", C.c_uint64), ("cyc_mp", C.c_uint64), ("cyc_append", C.c_uint64),
        ("lap", C.c_uint64 * 16),
        ("kernel_launches", C.c_int),
    ]


class Counters(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64)]


def lib_path():
    if os.environ.get("FB200_LIB"):         # experiment builds (make laps): never the default
        return os.path.abspath(os.environ["FB200_LIB"])
    return os.path.join(_HERE, "lib", "libfiasco_b200.so")


def load():
    """Load the CUDA C-ABI library; fail loudly if it was not built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise FB200Error(ENODEVICE, "%s not built -- run `make product` (there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    vp, ip, cp = C.c_void_p, C.c_int, C.c_char_p
    lib.fb200_version.restype = cp
    lib.fb200_device_count.restype = ip
    lib.fb200_params_init.argtypes = [C.POINTER(Params), ip, ip, ip, C.c_float, ip, cp, C.c_size_t]
    lib.fb200_create.argtypes = [C.POINTER(vp), C.POINTER(Params), ip, ip, cp, C.c_size_t]
    if hasattr(lib, "fb200_create_predicted"):      # absent in A/B builds of older revisions (tools/build_prev.sh)
        lib.fb200_create_predicted.argtypes = [C.POINTER(vp), C.POINTER(Params), C.POINTER(Motion), ip, ip, cp,
                                               C.c_size_t]
        lib.fb200_encode_predicted.argtypes = [vp, ip, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(_Wfa),
                                               cp, C.c_size_t]
    lib.fb200_destroy.argtypes = [vp]
    lib.fb200_destroy.restype = None
    lib.fb200_encode_tiles.argtypes = [vp, ip, C.POINTER(vp), C.POINTER(_Wfa), C.POINTER(TraceRec), ip,
                                       C.POINTER(ip), cp, C.c_size_t]
    lib.fb200_upload.argtypes = [vp, ip, C.POINTER(vp), cp, C.c_size_t]
    lib.fb200_launch.argtypes = [vp, ip, vp, cp, C.c_size_t]
    lib.fb200_download.argtypes = [vp, ip, C.POINTER(_Wfa), C.POINTER(TraceRec), ip, C.POINTER(ip), cp, C.c_size_t]
    lib.fb200_sync.argtypes = [vp, cp, C.c_size_t]
    lib.fb200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.fb200_get_stats.restype = None
    lib.fb200_resident_tiles.argtypes = [vp]
    lib.fb200_state_capacity.argtypes = [vp]
    if hasattr(lib, "fb200_motion_norms"):
        lib.fb200_motion_norms.argtypes = [ip, vp, vp, ip, ip, ip, ip, vp, C.POINTER(C.c_float), cp, C.c_size_t]
    lib.fb200_probe.argtypes = [ip, ip, vp, vp, vp, vp, vp, vp, cp, C.c_size_t]
    if hasattr(lib, "fb200_counters_get"):
        lib.fb200_counters_get.argtypes = [C.POINTER(Counters)]
        lib.fb200_counters_get.restype = None
        lib.fb200_counters_reset.restype = None
    _LIB = lib
    return lib


def counters(reset=False):
    """Process-wide work counters of the library (whoever launched: TileEncoder or fiasco_coder())."""
    c = Counters()
    lib = load()
    lib.fb200_counters_get(C.byref(c))
    if reset:
        lib.fb200_counters_reset()
    return {"kernel_ms": c.kernel_ms, "launches": int(c.launches), "h2d_bytes": int(c.h2d_bytes),
            "d2h_bytes": int(c.d2h_bytes)}


def device_count():
    return load().fb200_device_count()


def _check(rc, err):
    if rc != 0:
        raise FB200Error(rc, err.value.decode(errors="replace"))


def make_params(width, height, bands=1, quality=20.0, optimize=0, state_capacity=0):
    p = Params()
    err = C.create_string_buffer(512)
    _check(load().fb200_params_init(C.byref(p), width, height, bands, quality, optimize, err, 512), err)
    p.state_capacity = state_capacity
    return p


def pixels_from_grey(img_u8):
    """lib/image.c:362: 8-bit grey sample -> the coder's int16 12.4 fixed point."""
    return ((img_u8.astype(np.int32) - 128) * 16).astype(np.int16)


class _WfaArrays:
    def __init__(self, cap):
        self.cap = cap
        self.final_distribution = np.zeros(cap, np.float32)
        self.level_of_state = np.zeros(cap, np.uint8)
        self.domain_type = np.zeros(cap, np.uint8)
        self.tree = np.zeros((cap, 2), np.int16)
        self.x = np.zeros((cap, 2), np.uint16)
        self.y = np.zeros((cap, 2), np.uint16)
        self.into = np.zeros((cap, 2, 6), np.int16)
        self.weight = np.zeros((cap, 2, 6), np.float32)
        self.y_state = np.zeros((cap, 2), np.int16)
        self.y_column = np.zeros((cap, 2), np.uint8)
        self.mv_type = np.zeros((cap, 2), np.int8)
        self.mv_fx = np.zeros((cap, 2), np.int8)
        self.mv_fy = np.zeros((cap, 2), np.int8)
        self.mv_bx = np.zeros((cap, 2), np.int8)
        self.mv_by = np.zeros((cap, 2), np.int8)
        self.y_column_history = np.zeros((6000, 2), np.uint8)

    def fill(self, w):
        w.capacity = self.cap
        for name in ("final_distribution", "level_of_state", "domain_type", "tree", "x", "y", "into", "weight",
                     "y_state", "y_column", "mv_type", "mv_fx", "mv_fy", "mv_bx", "mv_by", "y_column_history"):
            setattr(w, name, getattr(self, name).ctypes.data)


class TileEncoder:
    """Device workspace for up to `max_tiles` tiles of one geometry (fb200_ctx_t)."""

    def __init__(self, params, max_tiles=1, device=0, out_capacity=None, motion=None):
        """motion = Motion(...): a workspace for predicted frames (fb200_create_predicted)."""
        self.lib = load()
        self.params = params
        self.max_tiles = max_tiles
        self.motion = motion
        self.ctx = C.c_void_p()
        err = C.create_string_buffer(512)
        if motion is None:
            _check(self.lib.fb200_create(C.byref(self.ctx), C.byref(params), max_tiles, device, err, 512), err)
        else:
            _check(self.lib.fb200_create_predicted(C.byref(self.ctx), C.byref(params), C.byref(motion), max_tiles,
                                                   device, err, 512), err)
        self.out_capacity = out_capacity or 6000
        self._arrays = [_WfaArrays(self.out_capacity) for _ in range(max_tiles)]
        self._wfas = (_Wfa * max_tiles)()
        for a, w in zip(self._arrays, self._wfas):
            a.fill(w)
        self._keep = None

    def close(self):
        if self.ctx:
            self.lib.fb200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _plane_ptrs(self, planes):
        planes = [np.ascontiguousarray(p, dtype=np.int16) for p in planes]
        n = self.params.width * self.params.height
        for p in planes:
            assert p.size == n, (p.shape, self.params.width, self.params.height)
        arr = (C.c_void_p * len(planes))(*[p.ctypes.data for p in planes])
        self._keep = planes
        return arr

    def _collect(self, n_tiles, trace, trace_len):
        out = []
        for t in range(n_tiles):
            w, a = self._wfas[t], self._arrays[t]
            n = w.states
            d = {
                "status": w.status, "states": n, "basis_states": w.basis_states, "root_state": w.root_state,
                "costs": list(w.costs), "err": list(w.err), "tree_bits": list(w.tree_bits),
                "matrix_bits": list(w.matrix_bits), "weights_bits": list(w.weights_bits),
                "lc_min_level": w.lc_min_level,
                "progress": [[p for p in range(128) if w.progress[b][p >> 5] >> (p & 31) & 1] for b in range(3)],
            }
            for name in ("final_distribution", "level_of_state", "domain_type", "tree", "x", "y", "into", "weight",
                         "y_state", "y_column") + (("mv_type", "mv_fx", "mv_fy", "mv_bx", "mv_by") if self.motion is not None else ()):
                d[name] = getattr(a, name)[:n].copy()
            out.append(d)
        tr = None
        if trace is not None:
            tr = [trace[i] for i in range(min(trace_len.value, len(trace)))]
        return out, tr

    def encode(self, planes, trace_cap=0):
        """Host-buffer path: planes = [tile0 band0, (tile0 band1, ...), tile1 band0, ...]."""
        n_tiles = len(planes) // self.params.bands
        ptrs = self._plane_ptrs(planes)
        err = C.create_string_buffer(512)
        trace = (TraceRec * trace_cap)() if trace_cap else None
        tl = C.c_int(0)
        rc = self.lib.fb200_encode_tiles(self.ctx, n_tiles, ptrs, self._wfas, trace, trace_cap, C.byref(tl), err, 512)
        _check(rc, err)
        return self._collect(n_tiles, trace, tl)

    def encode_predicted(self, planes, past, future=None, lc_min=None):
        """One predicted frame per tile: planes the frames' bands (1 or 3 per tile), past[t] the regenerated
        previous frame (all bands behind each other; future[t]: the regenerated next reference, B frames).
        past=None: an intra frame with nondeterministic prediction (a context of frame type FB200_FRAME_ND
        = 3).  lc_min[t]: the range level the frame starts with (frames of a colour sequence are chained)."""
        bands = self.params.bands
        n_tiles = len(planes) // bands
        ptrs = self._plane_ptrs(planes)
        keep = [self._keep]

        def whole(frames):
            frames = [np.ascontiguousarray(f, dtype=np.int16) for f in frames]
            for f in frames:
                assert f.size == bands * self.params.width * self.params.height
            keep.append(frames)
            return (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])

        pptrs = whole(past) if past is not None else None
        fptrs = whole(future) if future is not None else None
        self._keep = keep
        for t in range(n_tiles):
            self._wfas[t].lc_min_level = int(lc_min[t]) if lc_min is not None else 0
        err = C.create_string_buffer(512)
        _check(self.lib.fb200_encode_predicted(self.ctx, n_tiles, ptrs, pptrs, fptrs, self._wfas, err, 512), err)
        return self._collect(n_tiles, None, None)[0]

    def upload(self, planes):
        n_tiles = len(planes) // self.params.bands
        err = C.create_string_buffer(512)
        _check(self.lib.fb200_upload(self.ctx, n_tiles, self._plane_ptrs(planes), err, 512), err)
        return n_tiles

    def launch(self, n_tiles, stream=None):
        err = C.create_string_buffer(512)
        _check(self.lib.fb200_launch(self.ctx, n_tiles, C.c_void_p(stream or 0), err, 512), err)

    def sync(self):
        err = C.create_string_buffer(512)
        _check(self.lib.fb200_sync(self.ctx, err, 512), err)

    def download(self, n_tiles):
        err = C.create_string_buffer(512)
        tl = C.c_int(0)
        _check(self.lib.fb200_download(self.ctx, n_tiles, self._wfas, None, 0, C.byref(tl), err, 512), err)
        return self._collect(n_tiles, None, tl)[0]

    def resident_tiles(self):
        return self.lib.fb200_resident_tiles(self.ctx)

    def state_capacity(self):
        return self.lib.fb200_state_capacity(self.ctx)

    def stats(self):
        s = Stats()
        self.lib.fb200_get_stats(self.ctx, C.byref(s))
        d = {k: getattr(s, k) for k, _ in Stats._fields_}
        d["lap"] = list(s.lap)
        return d


def motion_norms(orig, past, level=6, search_range=16, device=0):
    """Norms tables of the motion search for all blocks of `level` (fb200_motion_norms): orig / past int16
    (h, w) in the coder's pixel format.  Returns (float32 [block rows][block columns][(2 sr)^2], kernel ms)."""
    orig = np.ascontiguousarray(orig, np.int16)
    past = np.ascontiguousarray(past, np.int16)
    h, w = orig.shape
    bw, bh = 1 << (level >> 1), 1 << ((level + 1) >> 1)
    out = np.zeros(((h + bh - 1) // bh, (w + bw - 1) // bw, 4 * search_range * search_range), np.float32)
    ms = C.c_float(0)
    err = C.create_string_buffer(512)
    _check(load().fb200_motion_norms(device, orig.ctypes.data, past.ctypes.data, w, h, level, search_range,
                                     out.ctypes.data, C.byref(ms), err, 512), err)
    return out, ms.value


def probe(kind, f=None, a=None, b=None, c=None):
    """Evaluate the device's rtob / btor / bits_bin_code / -log2 on arrays (KAT helper)."""
    lib = load()
    arrs = [x for x in (f, a, b, c) if x is not None]
    n = len(arrs[0])
    f_ = np.ascontiguousarray(f, np.float32) if f is not None else None
    a_, b_, c_ = [np.ascontiguousarray(x, np.int32) if x is not None else None for x in (a, b, c)]
    oi = np.zeros(n, np.int32)
    of = np.zeros(n, np.float32)
    err = C.create_string_buffer(512)

    def ptr(x):
        return C.c_void_p(x.ctypes.data) if x is not None else C.c_void_p(0)

    _check(lib.fb200_probe(kind, n, ptr(f_), ptr(a_), ptr(b_), ptr(c_), ptr(oi), ptr(of), err, 512), err)
    return oi, of


def wfa_lines(w, image_level=None):
    """Canonical text form of one automaton (dict from TileEncoder): "s" state lines, "e" edge lines."""
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
