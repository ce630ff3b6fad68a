"""Sequences with predicted frames over one or several GPUs (BASELINE config 5; SURVEY.md 8e: "video:
split by GOP").  A group of pictures -- an I frame and the P frames that follow it -- is a chain (every P
frame needs the regenerated frame before it, codec/coder.c:642-651); the groups are independent.  Group g
goes to rank g mod world; a rank codes the I frames of its groups in one launch and then, step by step, the
k-th P frame of all its groups in one launch (one thread block per group).  One gather at the end brings
the finished automata to rank 0, which writes the single FIASCO stream.

Orchestration only: every computation is a C entry point (fb200_encode_tiles, fb200_encode_predicted,
fiasco_finish_predicted_frame, fiasco_regenerate_frame, fiasco_write_video_stream); the same loop in C is
the frame loop of fiasco_coder() (fiasco_b200/host/coder_api.c)."""
import io

import numpy as np

from . import ffi, hostlib

_ARRAYS = ("final_distribution", "level_of_state", "domain_type", "tree", "x", "y", "into", "weight",
           "y_state", "y_column", "mv_type", "mv_fx", "mv_fy", "mv_bx", "mv_by", "delta_state")
_SCALARS = ("states", "basis_states", "root_state", "frame_type", "frame_number")


def frame_types(n_frames, pattern):
    """0 = I, 1 = P per display frame (frame 0 is always intra, coder.c:540-560)."""
    out = []
    for f in range(n_frames):
        t = pattern[f % len(pattern)].upper()
        if t not in "IP":
            raise ValueError("only I and P frames are coded on the device (pattern %r)" % pattern)
        out.append(0 if f == 0 or t == "I" else 1)
    return out


def groups(n_frames, pattern):
    """[(first, last + 1)] of the groups of pictures."""
    types = frame_types(n_frames, pattern)
    starts = [f for f in range(n_frames) if types[f] == 0]
    return [(s, e) for s, e in zip(starts, starts[1:] + [n_frames])]


def _with_motion_fields(w, frame_type, number):
    n = w["states"]
    w = dict(w)
    w["frame_type"], w["frame_number"] = frame_type, number
    for k in ("mv_type", "mv_fx", "mv_fy", "mv_bx", "mv_by"):
        w.setdefault(k, np.zeros((n, 2), np.int8))
    w.setdefault("delta_state", np.zeros(n, np.uint8))
    return w


def encode_groups(planes, my_groups, params, p_min_level=6, p_max_level=10, search_range=16, device=0):
    """planes: {frame number: int16 (h, w) plane} for the frames of `my_groups`.  Returns
    {frame number: finished automaton dict} and the device milliseconds spent in the kernels."""
    out, kernel_ms = {}, 0.0
    if not my_groups:
        return out, kernel_ms
    W, H = params.width, params.height
    enc = ffi.TileEncoder(params, len(my_groups), device)
    try:
        ws, _ = enc.encode([planes[s] for s, _ in my_groups])
        kernel_ms += enc.stats()["kernel_ms"]
    finally:
        enc.close()
    for (s, _), w in zip(my_groups, ws):
        out[s] = _with_motion_fields(w, 0, s)
    chains = [g for g in my_groups if g[1] - g[0] > 1]
    if not chains:
        return out, kernel_ms
    recon = {}
    penc = ffi.TileEncoder(params, len(chains), device, motion=ffi.Motion(1, p_min_level, p_max_level, search_range))
    try:
        k = 1
        while True:
            todo = [s + k for s, e in chains if s + k < e]
            if not todo:
                break
            for f in todo:                         # the references: regenerate the frames before
                prev = out[f - 1]
                recon[f - 1] = hostlib.regenerate_frame(prev, W, H, recon.get(f - 2) if prev["frame_type"] else None)
            gs = penc.encode_predicted([planes[f] for f in todo], [recon[f - 1] for f in todo])
            kernel_ms += penc.stats()["kernel_ms"]
            for f, g in zip(todo, gs):
                out[f] = hostlib.finish_predicted_frame(_with_motion_fields(g, 1, f))
                if k >= 2:
                    recon.pop(f - 2, None)         # two frames back in the same chain: no longer a reference
            k += 1
    finally:
        penc.close()
    return out, kernel_ms


def pack(w):
    """One automaton as bytes (for the gather)."""
    buf = io.BytesIO()
    n = int(w["states"])
    np.savez(buf, scalars=np.array([int(w[k]) for k in _SCALARS], np.int64),
             **{k: np.asarray(w[k])[:n] for k in _ARRAYS})
    return buf.getvalue()


def unpack(b):
    z = np.load(io.BytesIO(b))
    w = {k: z[k] for k in _ARRAYS}
    w.update({k: int(v) for k, v in zip(_SCALARS, z["scalars"])})
    return w


def encode_sequence(planes, pattern, params, rank=0, world=1, gather_device="cpu", **kw):
    """All ranks call this with the same arguments (planes: list of int16 (h, w) planes in display order;
    a rank only touches the frames of its own groups).  Returns on rank 0 the list of finished automata in
    display order (None elsewhere) and this rank's kernel milliseconds."""
    from . import distributed as D
    gl = groups(len(planes), pattern)
    mine = [gl[i] for i in D.shard(len(gl), rank, world)]
    done, ms = encode_groups({f: planes[f] for s, e in mine for f in range(s, e)}, mine, params, **kw)
    local = {i: b"".join(len(p).to_bytes(8, "little") + p for p in (pack(done[f]) for f in range(*gl[i])))
             for i in D.shard(len(gl), rank, world)}
    allb = D.gather_streams(local, len(gl), rank, world, device=gather_device)
    if allb is None:
        return None, ms
    seq = []
    for b in allb:
        o = 0
        while o < len(b):
            n = int.from_bytes(b[o:o + 8], "little")
            seq.append(unpack(b[o + 8:o + 8 + n]))
            o += 8 + n
    return seq, ms
