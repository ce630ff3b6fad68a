"""Parity tests proper: the CUDA path (through the C ABI, include/fiasco_b200.h) against the
oracle on the same seeded inputs, and against the golden vectors taken from the reference
binary.  Integer / index work must be bit exact; the fp32 weights are quantised codes, so they
are compared bit-for-bit as well (tolerance 0, stricter than the 1e-4 the north star allows)."""
import os
import struct

import numpy as np
import pytest

import fiasco_b200 as F
from fiasco_b200 import ffi
import oracle_lib as O
import gen_frames

pytestmark = pytest.mark.gpu


def gpu_encode(img, quality=20.0, optimize=0, cap=0, trace=False):
    h, w = img.shape[:2]
    p = ffi.make_params(w, h, 3 if img.ndim == 3 else 1, quality, optimize, cap)
    enc = F.TileEncoder(p, 1)
    try:
        ws, tr = enc.encode(O.planes_of(img), trace_cap=300000 if trace else 0)
        return ws[0], tr, enc.stats()
    finally:
        enc.close()


def fb(x):
    return int(np.float32(x).view(np.uint32))


def trace_line(i, r):
    s = "lc %d %d %d %d %d %d %d %d %08x %08x %08x" % (i, r.level, r.image, r.address, r.x, r.y, r.y_state, r.states,
                                                      fb(r.max_costs), fb(r.price), fb(r.costs))
    if r.n_edges >= 0 and r.into[0] >= 0:
        s += " %08x %08x %08x :" % (fb(r.err), fb(r.matrix_bits), fb(r.weights_bits))
        for e in range(r.n_edges):
            s += " %d:%08x" % (r.into[e], fb(r.weight[e]))
    return s


def assert_same_wfa(gw, ow, bands=1):
    assert gw["status"] == 0
    assert gw["states"] == ow["states"]
    assert gw["root_state"] == ow["root_state"]
    assert F.wfa_lines(gw) == O.wfa_lines(ow)
    n = ow["states"]
    assert np.array_equal(gw["final_distribution"].view(np.uint32), ow["final_distribution"].view(np.uint32))
    assert np.array_equal(gw["domain_type"][:n], ow["domain_type"][:n])
    assert np.array_equal(gw["y_state"][:n], ow["y_state"][:n])
    assert np.array_equal(gw["y_column"][:n], ow["y_column"][:n])
    for band in range(bands):
        assert fb(gw["costs"][band]) == fb(ow["costs"][band])
        assert fb(gw["err"][band]) == fb(ow["err"][band])
        for k in ("tree_bits", "matrix_bits", "weights_bits"):
            assert fb(gw[k][band]) == fb(ow[k][band]), (k, band)


# ------------------------------------------------------------------ device arithmetic

def test_device_pure_functions_known_answers():
    """rtob / btor / bits_bin_code on the device against the reference's own outputs."""
    kat = O.golden_kat()
    rt = [l.split() for l in kat if l.startswith("rtob ")]
    f = np.array([int(x[3], 16) for x in rt], np.uint32).view(np.float32)
    oi, _ = F.probe(0, f=f, a=[int(x[1]) for x in rt], b=[int(x[2]) for x in rt])
    assert np.array_equal(oi, np.array([int(x[4]) for x in rt], np.int32))
    bt = [l.split() for l in kat if l.startswith("btor ")]
    _, of = F.probe(1, a=[int(x[3]) for x in bt], b=[int(x[1]) for x in bt], c=[int(x[2]) for x in bt])
    assert np.array_equal(of.view(np.uint32), np.array([int(x[4], 16) for x in bt], np.uint32))
    bb = [l.split() for l in kat if l.startswith("bbc ")]
    oi, _ = F.probe(2, a=[int(x[1]) for x in bb], b=[int(x[2]) for x in bb])
    assert np.array_equal(oi, np.array([int(x[3]) for x in bb], np.int32))


def test_device_log2_rate_terms_match_glibc():
    """-log2(count/(float)total) as fp32: CUDA's log2(double) vs glibc's, over every (count,total)
    with total <= 1500 plus a seeded sample up to the int16 range the models can reach."""
    L = O.lib()
    a, b = [], []
    for t in range(1, 1501):
        a.extend(range(1, t + 1))
        b.extend([t] * t)
    rng = np.random.default_rng(5)
    tt = rng.integers(1501, 32767, 200000)
    cc = (rng.random(200000) * tt).astype(np.int64) + 1
    a = np.concatenate([np.array(a), cc]).astype(np.int32)
    b = np.concatenate([np.array(b), tt]).astype(np.int32)
    _, of = F.probe(3, a=a, b=b)
    # glibc via the oracle's C helper (vectorised through ctypes would be slow: use numpy's
    # float64 log2 and confirm a sample against the C helper)
    want = (-np.log2((a.astype(np.float32) / b.astype(np.float32)).astype(np.float64))).astype(np.float32)
    idx = rng.integers(0, len(a), 20000)
    for i in idx:
        assert fb(L.fo_neg_log2f(int(a[i]), int(b[i]))) == fb(want[i])
    bad = np.nonzero(of.view(np.uint32) != want.view(np.uint32))[0]
    assert len(bad) == 0, (len(bad), a[bad[:5]], b[bad[:5]])


# ------------------------------------------------------------------ whole-path parity

SMALL = [("g1024", 0, 0, 64, 64, 20, 0), ("g1024", 256, 512, 64, 64, 20, 0), ("g1024", 640, 128, 128, 128, 20, 0),
         ("g1024", 0, 0, 200, 136, 20, 0), ("g512", 100, 60, 96, 64, 40, 0), ("g512", 300, 200, 64, 96, 30, 0),
         ("g512", 17 * 2, 33 * 2, 34, 38, 20, 0), ("g256", 0, 0, 256, 256, 20, 0), ("g256", 0, 0, 256, 256, 60, 0),
         ("g256", 0, 0, 128, 256, 8, 0), ("g256", 64, 64, 128, 128, 20, 1), ("g256", 0, 0, 128, 128, 20, 2)]


@pytest.mark.parametrize("frame,x0,y0,w,h,q,z", SMALL)
def test_gpu_matches_oracle_trace_and_wfa(frame, x0, y0, w, h, q, z):
    img = np.ascontiguousarray(gen_frames.frame(frame)[y0:y0 + h, x0:x0 + w])
    ow = O.encode(img, quality=q, optimize=z, want_trace=True)
    gw, tr, _ = gpu_encode(img, q, z, trace=True)
    olc = O.lc_lines(ow["trace"])
    glc = [trace_line(i, r) for i, r in enumerate(tr)]
    for i, (a, b) in enumerate(zip(olc, glc)):
        assert a == b, "first divergence at approximate_range call %d" % i
    assert len(olc) == len(glc)
    assert_same_wfa(gw, ow)


COLOUR = [("c256", 0, 0, 64, 64, 20, 0), ("c256", 64, 128, 128, 128, 20, 0), ("c256", 0, 0, 256, 256, 20, 0),
          ("c256", 0, 0, 256, 256, 30, 0), ("c256", 32, 16, 96, 64, 45, 0), ("c256", 0, 0, 128, 128, 20, 1)]


@pytest.mark.parametrize("frame,x0,y0,w,h,q,z", COLOUR)
def test_gpu_colour_matches_oracle(frame, x0, y0, w, h, q, z):
    """Y, Cb, Cr bands (4:4:4) with the chroma dictionary and the y-state threading."""
    img = np.ascontiguousarray(gen_frames.frame(frame)[y0:y0 + h, x0:x0 + w])
    ow = O.encode(img, quality=q, optimize=z, want_trace=True)
    gw, tr, _ = gpu_encode(img, q, z, trace=True)
    olc = O.lc_lines(ow["trace"])
    glc = [trace_line(i, r) for i, r in enumerate(tr)]
    for i, (a, b) in enumerate(zip(olc, glc)):
        assert a == b, "first divergence at approximate_range call %d" % i
    assert len(olc) == len(glc)
    assert_same_wfa(gw, ow, bands=3)


@pytest.mark.parametrize("name", ["c256_q20_z0", "c256_q30_z0", "c2048t0_q30_z0"])
def test_gpu_colour_matches_reference_golden(name):
    m = O.manifest()[name]
    gw, _, _ = gpu_encode(O.case_image(name), m["quality"], m["optimize"])
    level = O.lib().fo_image_level(m["width"], m["height"])
    assert O.mask_virtual(F.wfa_lines(gw), level) == O.golden_wfa_lines(name)


def test_gpu_flat_and_noise_edge_cases():
    """Constant image (every range is pure DC), saturated black/white, and white noise."""
    rng = np.random.default_rng(11)
    imgs = [np.full((64, 64), 128, np.uint8), np.zeros((64, 64), np.uint8), np.full((32, 32), 255, np.uint8),
            rng.integers(0, 256, (64, 64)).astype(np.uint8), rng.integers(0, 256, (128, 64)).astype(np.uint8)]
    for img in imgs:
        ow = O.encode(img, quality=20, optimize=0)
        gw, _, _ = gpu_encode(img, 20, 0)
        assert_same_wfa(gw, ow)


GOLD = ["g256_q20_z0", "g1024t0_q20_z0", "g1024t15_q20_z0", "g1024r_q20_z0", "g1024s_q40_z0", "g512_q20_z0",
        "g4096t0_q20_z0", "g256_q20_z1", "g256_q20_z2"]


@pytest.mark.parametrize("name", GOLD)
def test_gpu_matches_reference_golden(name):
    """Device output vs the WFA the reference binary itself produced (tests/golden)."""
    m = O.manifest()[name]
    img = O.case_image(name)
    gw, _, _ = gpu_encode(img, m["quality"], m["optimize"])
    assert F.wfa_lines(gw) == O.golden_wfa_lines(name)


def test_gpu_full_frame_1024_matches_reference():
    """BASELINE.json config[1]: 1024x1024 grey, q=20, monolithic -- bit-identical automaton."""
    name = "g1024_q20_z0"
    img = O.case_image(name)
    gw, _, st = gpu_encode(img, 20, 0)
    assert gw["states"] == 1487
    assert F.wfa_lines(gw) == O.golden_wfa_lines(name)
    assert st["mp_calls"] == 20385


def test_gpu_batch_of_tiles_equals_individual_streams():
    """16 independent 256^2 streams of the 1024^2 frame in ONE launch (one thread block per tile)
    give exactly the per-tile oracle automata; order inside the batch does not matter."""
    crops = gen_frames.crops(gen_frames.frame("g1024"), 256)
    p = ffi.make_params(256, 256, 1, 20.0, 0)
    enc = F.TileEncoder(p, 16)
    try:
        planes = [O.planes_of(c)[0] for c in crops]
        ws, _ = enc.encode(planes)
        perm = np.random.default_rng(3).permutation(16)
        ws2, _ = enc.encode([planes[i] for i in perm])
    finally:
        enc.close()
    for k in (0, 5, 15):
        assert_same_wfa(ws[k], O.encode(crops[k], quality=20, optimize=0))
    for j, i in enumerate(perm):
        assert F.wfa_lines(ws2[j]) == F.wfa_lines(ws[i])
    assert F.wfa_lines(ws[0]) == O.golden_wfa_lines("g1024t0_q20_z0")
    assert F.wfa_lines(ws[15]) == O.golden_wfa_lines("g1024t15_q20_z0")


def test_gpu_more_tiles_than_workspaces():
    """A launch with more tiles than can be resident (the big tables exist once per RESIDENT tile;
    blocks take a workspace when they start and return it when done): every tile of three waves
    equals the automaton of the same crop encoded alone, and the rtob o btor identity the kernel's
    quantiser tables rely on holds on the device."""
    g = gen_frames.frame("g1024")
    crops = [np.ascontiguousarray(g[y:y + 64, x:x + 64]) for y in range(0, 512, 64) for x in range(0, 512, 64)]
    p = ffi.make_params(64, 64, 1, 20.0, 0)
    probe = F.TileEncoder(p, 1)
    resident = probe.resident_tiles()
    probe.close()
    assert resident >= 148
    n = 2 * resident + 37
    enc = F.TileEncoder(p, n)
    try:
        planes = [O.planes_of(crops[i % len(crops)])[0] for i in range(n)]
        ws, _ = enc.encode(planes)
        ws_again, _ = enc.encode(planes)             # the flags are all free again after a launch
    finally:
        enc.close()
    single = {}
    one = F.TileEncoder(p, 1)
    try:
        for k in range(len(crops)):
            single[k] = F.wfa_lines(one.encode(O.planes_of(crops[k]))[0][0])
    finally:
        one.close()
    for i in range(n):
        assert F.wfa_lines(ws[i]) == single[i % len(crops)], "tile %d of %d (resident %d)" % (i, n, resident)
    for i in (0, resident, n - 1):
        assert F.wfa_lines(ws_again[i]) == single[i % len(crops)]
    for k in (0, 17, 63):
        assert_same_wfa(ws[k], O.encode(crops[k], quality=20, optimize=0))


def test_gpu_deterministic_and_reusable_context():
    img = gen_frames.frame("g256")
    p = ffi.make_params(256, 256, 1, 20.0, 0)
    enc = F.TileEncoder(p, 1)
    try:
        a, _ = enc.encode(O.planes_of(img))
        b, _ = enc.encode(O.planes_of(img[::-1].copy()))
        c, _ = enc.encode(O.planes_of(img))
    finally:
        enc.close()
    assert F.wfa_lines(a[0]) == F.wfa_lines(c[0])
    assert F.wfa_lines(a[0]) != F.wfa_lines(b[0])


def test_gpu_structural_invariants_full_size():
    """Size-independent properties on the 1024^2 frame: every state's edges are sorted by target
    and point to earlier, usable states; weights are fixed points of the quantiser; the tree is
    a proper bintree covering the frame."""
    img = gen_frames.frame("g1024")
    gw, _, _ = gpu_encode(img, 20, 0)
    n, basis = gw["states"], gw["basis_states"]
    L = O.lib()
    covered = 0
    for s in range(basis, n):
        lvl = int(gw["level_of_state"][s])
        for label in range(2):
            tgt = [int(t) for t in gw["into"][s][label] if t >= 0][: 6]
            into = []
            for t in gw["into"][s][label]:
                if t < 0:
                    break
                into.append(int(t))
            assert into == sorted(into) and len(set(into)) == len(into)
            assert all(t < s and gw["domain_type"][t] == 2 for t in into)
            for e, t in enumerate(into):
                wgt = float(gw["weight"][s][label][e])
                m, r = (5, 1) if t == 0 else (3, 2)
                assert wgt != 0 and fb(L.fo_btor(L.fo_rtob(wgt, m, r), m, r)) == fb(wgt)
            child = int(gw["tree"][s][label])
            assert child < s
            if child >= 0:
                assert int(gw["level_of_state"][child]) == lvl - 1
            else:
                covered += 1 << (lvl - 1)
    assert covered == 1024 * 1024
    assert int(gw["level_of_state"][gw["root_state"]]) == 20


def test_gpu_capacity_error_is_reported():
    img = gen_frames.frame("g256")
    p = ffi.make_params(256, 256, 1, 20.0, 0, 64)
    enc = F.TileEncoder(p, 1)
    try:
        with pytest.raises(F.FB200Error) as e:
            enc.encode(O.planes_of(img))
        assert e.value.code == ffi.ECAPACITY
    finally:
        enc.close()


def test_gpu_random_crops_match_oracle():
    """Seeded random campaign (tools/fuzz_gpu.py): 60 random crops, sizes, qualities, optimisation
    levels, grey and colour, flat blocks -- device vs oracle, bit for bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_gpu.py"), "60", "11"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "60 cases, 0 mismatches" in r.stdout


def test_gpu_config3_colour_2048_all_tiles():
    """BASELINE config 3 at full size: 2048^2 colour (4:4:4), q = 30, 64 independent 256^2 tiles in ONE
    launch: tile 0 is the reference's golden automaton, sampled tiles equal the oracle, every tile
    is a complete automaton (three bands under two virtual roots)."""
    img = gen_frames.frame("c2048")
    crops = gen_frames.crops(img, 256)
    assert len(crops) == 64
    p = ffi.make_params(256, 256, 3, 30.0, 0)
    enc = F.TileEncoder(p, 64)
    try:
        planes = [pl for c in crops for pl in O.planes_of(c)]
        ws, _ = enc.encode(planes)
    finally:
        enc.close()
    level = O.lib().fo_image_level(256, 256)
    assert O.mask_virtual(F.wfa_lines(ws[0]), level) == O.golden_wfa_lines("c2048t0_q30_z0")
    for k in (7, 36, 63):
        ow = O.encode(crops[k], quality=30, optimize=0)
        assert ws[k]["states"] == ow["states"]
        assert O.mask_virtual(F.wfa_lines(ws[k]), level) == O.mask_virtual(O.wfa_lines(ow), level), "tile %d" % k
    for w in ws:
        assert int(w["level_of_state"][w["root_state"]]) == level + 2


def test_gpu_config4_grey_4096_all_tiles():
    """BASELINE config 4 at full size: 4096^2 grey, q = 20, 64 independent 512^2 tiles in one launch."""
    img = gen_frames.frame("g4096")
    crops = gen_frames.crops(img, 512)
    assert len(crops) == 64
    p = ffi.make_params(512, 512, 1, 20.0, 0)
    enc = F.TileEncoder(p, 64)
    try:
        ws, _ = enc.encode([O.planes_of(c)[0] for c in crops])
    finally:
        enc.close()
    assert F.wfa_lines(ws[0]) == O.golden_wfa_lines("g4096t0_q20_z0")
    for k in (21, 63):
        assert_same_wfa(ws[k], O.encode(crops[k], quality=20, optimize=0))
    for w in ws:
        assert int(w["level_of_state"][w["root_state"]]) == 18


@pytest.mark.parametrize("level,sr", [(6, 16), (8, 16), (9, 8), (7, 5)])
def test_gpu_motion_norms_match_reference_loops(level, sr):
    """The norms tables of the motion search (fb200_motion_norms: all blocks of a frame in one launch)
    against the oracle's restatement of the reference's per-block fill_norms_table (codec/mwfa.c:544),
    bit for bit: frame 1 of the golden sequence against the REGENERATED frame 0, blocks in the
    interior, on every border and partly outside."""
    import gzip
    m = O.manifest()["v160_q20_ippp"]
    frames = list(gen_frames.video(2, m["width"], m["height"]))
    orig = O.planes_of(frames[1])[0].reshape(m["height"], m["width"])
    raw = np.frombuffer(gzip.open(os.path.join(O.GOLDEN, "v160_q20_ippp.decoded.raw.gz")).read(), np.int16)
    past = raw.reshape(m["frames"], m["height"], m["width"])[0].copy()
    got, ms = ffi.motion_norms(orig, past, level, sr)
    bw, bh = 1 << (level >> 1), 1 << ((level + 1) >> 1)
    L = O.lib()
    ref = np.zeros(4 * sr * sr, np.float32)
    nby, nbx = got.shape[:2]
    assert (nby, nbx) == ((m["height"] + bh - 1) // bh, (m["width"] + bw - 1) // bw)
    checked = 0
    for by in range(nby):
        for bx in range(nbx):
            if bx * bw + bw > m["width"] or by * bh + bh > m["height"]:
                assert not got[by, bx].any()
                continue
            if (bx + 3 * by) % 3 and 0 < bx < nbx - 2 and 0 < by < nby - 2:
                continue                              # sample the interior, take every border block
            L.fo_fill_norms_table(orig.ctypes.data, past.ctypes.data, m["width"], m["height"], bx * bw, by * bh,
                                  level, sr, ref.ctypes.data)
            assert np.array_equal(got[by, bx].view(np.uint32), ref.view(np.uint32)), (bx, by)
            checked += 1
    assert checked > 20 and ms > 0


# ------------------------------------------------------------------ cluster per stream (round 2)

def _with_cluster(c, fn):
    saved = os.environ.get("FB200_CLUSTER")
    os.environ["FB200_CLUSTER"] = str(c)
    try:
        return fn()
    finally:
        if saved is None:
            os.environ.pop("FB200_CLUSTER", None)
        else:
            os.environ["FB200_CLUSTER"] = saved


CLUSTER_CASES = [("g256", 0, 0, 256, 256, 20, 0), ("g1024", 0, 0, 200, 136, 20, 0), ("g256", 64, 64, 128, 128, 20, 1),
                 ("g256", 0, 0, 128, 128, 20, 2), ("c256", 64, 128, 128, 128, 20, 0), ("c256", 0, 0, 256, 256, 30, 0)]


@pytest.mark.parametrize("cluster", [2, 4, 8])
@pytest.mark.parametrize("frame,x0,y0,w,h,q,z", CLUSTER_CASES)
def test_gpu_cluster_per_stream_matches_oracle_trace_and_wfa(cluster, frame, x0, y0, w, h, q, z):
    """One stream on a thread-block cluster: the helper blocks run the pursuits of a range's label-0
    descendants ahead (same models, same states: subdivide.c:188-237), their share of a block's products
    and of a new state's table levels.  Every approximate_range result, in the reference's order, and the
    automaton: bit for bit, for every cluster size (grey, colour, -z 1, -z 2, a ragged picture)."""
    img = np.ascontiguousarray(gen_frames.frame(frame)[y0:y0 + h, x0:x0 + w])
    ow = O.encode(img, quality=q, optimize=z, want_trace=True)
    gw, tr, st = _with_cluster(cluster, lambda: gpu_encode(img, q, z, trace=True))
    olc = O.lc_lines(ow["trace"])
    glc = [trace_line(i, r) for i, r in enumerate(tr)]
    for i, (a, b) in enumerate(zip(olc, glc)):
        assert a == b, "first divergence at approximate_range call %d" % i
    assert len(olc) == len(glc)
    assert_same_wfa(gw, ow, bands=3 if img.ndim == 3 else 1)
    # the work counters count the pursuits that were used, not the ones that ran ahead in vain
    # (-z 2: two pursuits per range, the second without the first one's first domain, approx.c:103-127)
    assert st["mp_calls"] == len(olc) * (2 if z == 2 else 1)


def test_gpu_cluster_full_frame_1024_matches_reference():
    """BASELINE config[1], the monolithic 1024^2 frame, on the cluster the launcher picks for one stream
    (8 blocks): the reference's golden automaton and work counters."""
    gw, _, st = gpu_encode(O.case_image("g1024_q20_z0"), 20.0, 0)
    assert F.wfa_lines(gw) == O.golden_wfa_lines("g1024_q20_z0")
    assert st["mp_calls"] == 20385 and st["mp_steps"] == 27768


def test_gpu_cluster_tiles_share_the_device():
    """Fewer streams than SMs: every stream gets a cluster (16 tiles -> clusters of 8 on a B200); same
    automata as the tiles coded alone on one block each."""
    img = gen_frames.frame("g1024")
    crops = gen_frames.crops(img, 256)
    p = ffi.make_params(256, 256, 1, 20.0, 0)
    planes = [ffi.pixels_from_grey(c).reshape(-1) for c in crops]

    def run():
        enc = F.TileEncoder(p, len(crops))
        try:
            return enc.encode(planes)[0]
        finally:
            enc.close()

    auto = run()
    alone = _with_cluster(1, run)
    for a, b in zip(auto, alone):
        assert F.wfa_lines(a) == F.wfa_lines(b)
