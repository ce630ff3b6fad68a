"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/fiasco_b200.h declares, clamps parameters like the reference's alloc_coder(), and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

import fiasco_b200 as F
from fiasco_b200 import ffi
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fb200_[a-z0-9_]+|fiasco_[a-z0-9_]+|open_file)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = F.load()
    syms = declared_symbols("fiasco_b200.h")
    assert len(syms) >= 14, syms
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert b"sm_100a" in lib.fb200_version()


def test_sass_is_sm100a():
    """The shipped library must carry sm_100a code (not PTX for another arch)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", F.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.parametrize("w,h", [(256, 256), (512, 512), (1024, 1024), (64, 64), (200, 136), (720, 576), (32, 32),
                                 (2048, 2048), (96, 64)])
@pytest.mark.parametrize("z", [0, 1, 2])
def test_params_clamping_matches_alloc_coder(w, h, z):
    """fb200_params_init follows codec/coder.c:249-296 (checked against the oracle's restatement,
    which is pinned to the reference)."""
    p = ffi.make_params(w, h, 1, 20.0, z)
    L = O.lib()
    level = L.fo_image_level(w, h)
    assert p.level == level
    lc_min, lc_max, edges = ((6, 10, 3) if z == 0 else (4, 12, 5))
    exp_max = min(lc_max, level - 1)
    exp_min = min(max(lc_min, 3), exp_max)
    assert (p.lc_min_level, p.lc_max_level) == (exp_min, exp_max)
    assert p.images_level == min(5, exp_max - 1)
    assert p.max_elements == edges
    assert p.second_domain_block == (1 if z == 2 else 0)
    import numpy as np
    assert np.float32(p.price) == np.float32(128 * 64) / np.float32(20.0)


def test_params_errors():
    with pytest.raises(F.FB200Error) as e:
        ffi.make_params(255, 256)
    assert "even" in str(e.value)
    with pytest.raises(F.FB200Error) as e:
        ffi.make_params(256, 256, 1, 0.0)
    assert "positive" in str(e.value)
    with pytest.raises(F.FB200Error) as e:
        ffi.make_params(256, 256, 1, 20.0, 3)
    assert e.value.code == ffi.EUNSUPPORTED


def test_no_cpu_fallback():
    """Without a CUDA device the compute entry points must refuse, not silently compute."""
    if F.device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = ffi.make_params(64, 64)
    with pytest.raises(F.FB200Error) as e:
        F.TileEncoder(p, 1)
    assert e.value.code == ffi.ENODEVICE
    with pytest.raises(F.FB200Error) as e:
        F.probe(2, a=[1, 2], b=[3, 4])
    assert e.value.code == ffi.ENODEVICE


def test_product_does_not_reference_oracle():
    """Nothing under fiasco_b200/ or include/ may include, link or import the oracle."""
    bad = []
    for base in ("fiasco_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".c", ".h", ".cu", ".cuh", ".py", ".mk")) or f == "Makefile":
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"fiasco_oracle|oracle_lib|liboracle|oracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's own CPU coder, all host cores, no GPU): one JSON line
    with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "cfiasco")):
        pytest.skip("reference binary not built here")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpixels/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["metric"].startswith("encoder Mpixels/s")


def test_product_does_not_know_the_emulator():
    """tests/emu (the device sources compiled for a CPU thread emulator) is test infrastructure: nothing
    the product builds, imports or runs refers to it -- the Python binding loads the CUDA library only,
    the Makefile, bench.py and the driver entry points do not mention it, and the device sources carry
    nothing but the launch / shared-memory spelling (#ifdef FB200_EMU) that lets them compile as C++."""
    import re
    from fiasco_b200 import ffi, hostlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert ffi.lib_path().endswith(os.path.join("fiasco_b200", "lib", "libfiasco_b200.so"))
    assert hostlib.lib_path().endswith(os.path.join("fiasco_b200", "lib", "libfiasco.so"))
    for rel in ("Makefile", "bench.py", "__graft_entry__.py", "fiasco_b200/ffi.py", "fiasco_b200/hostlib.py",
                "fiasco_b200/video.py", "fiasco_b200/distributed.py", "fiasco_b200/__init__.py",
                "fiasco_b200/host/coder_api.c", "fiasco_b200/csrc/ffi.cu"):
        text = open(os.path.join(root, rel)).read()
        assert not re.search(r"\bemu\b|_emu|emu_", text), rel
