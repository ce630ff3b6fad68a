import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# tests of the -m gpu suite that only mean something on real hardware (device timings, the CLI binary or a
# subprocess that loads the real library, workspace counts taken from the SM count)
_NEEDS_HARDWARE = ("test_gpu_more_tiles_than_workspaces", "test_gpu_random_crops_match_oracle",
                   "test_gpu_motion_norms_match_reference_loops", "test_unchanged_reference_cli_on_our_library",
                   "test_progress_meter_output", "test_gpu_cli_prediction_flag",
                   "test_gpu_fiasco_coder_colour_sequence_at_config5_size")


def pytest_addoption(parser):
    parser.addoption("--emu", action="store_true", default=False,
                     help="development aid for sessions without a GPU: run the -m gpu tests with the device "
                          "sources compiled for the CPU thread emulator of tests/emu (TEST INFRASTRUCTURE; "
                          "proves control flow and arithmetic order, not hardware behaviour)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if config.getoption("--emu"):
        emu = os.path.join(ROOT, "tests", "emu")
        subprocess.run(["make", "-s", "-C", emu], check=True)
        from fiasco_b200 import ffi, hostlib
        # FB200_EMU_ASAN=1: the sanitizer build (make -C tests/emu asan; LD_PRELOAD libasan.so)
        sub = "_asan" if os.environ.get("FB200_EMU_ASAN") else "_build"
        ffi.lib_path = lambda: os.path.join(emu, sub, "libfiasco_b200_emu.so")
        hostlib.lib_path = lambda: os.path.join(emu, sub, "libfiasco_emu.so")
        os.environ.setdefault("FB200_NT", "128")


def pytest_collection_modifyitems(config, items):
    """GPU tests are selected explicitly with -m gpu; without a device they fail loudly rather
    than skip (a silent skip would hide a missing CUDA path)."""
    if config.getoption("--emu"):
        skip = pytest.mark.skip(reason="needs real hardware (not meaningful under the emulator)")
        for item in items:
            if any(item.name.startswith(n) for n in _NEEDS_HARDWARE):
                item.add_marker(skip)
