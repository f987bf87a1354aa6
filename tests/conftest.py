import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="module")
def emu():
    """TEST INFRASTRUCTURE: binds the CPU kernel-emulator build of csrc/*.cu (tests/emu) for this module, so the
    real kernel sources are exercised without a GPU. The product never loads it."""
    from multiview_motion_capture_b200 import _lib
    from emu.build_emu import build_emulator
    path = build_emulator()
    old = (_lib._lib, _lib._lib_path, _lib._device)
    lib = _lib.use_library(path, device="cpu")
    yield lib
    _lib._lib, _lib._lib_path, _lib._device = old


@pytest.fixture(scope="module")
def cuda():
    """The real library on a real device (GPU tier)."""
    import torch
    from multiview_motion_capture_b200 import _lib
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    assert os.path.exists(_lib.LIB_PATH), "libmvmc.so is missing: run __graft_entry__.build() (no CPU fallback)"
    old = (_lib._lib, _lib._lib_path, _lib._device)
    lib = _lib.use_library(_lib.LIB_PATH)
    yield lib
    _lib._lib, _lib._lib_path, _lib._device = old
