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
