"""The C-ABI boundary without a GPU: libmvmc.so loads, exports every symbol include/mvmc.h declares, and its
host-only entry points and argument validation behave. No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import ROOT

from multiview_motion_capture_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mvmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvmc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mvmc.h but not exported by libmvmc.so"


def test_binding_covers_header():
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()


def test_record_layout_matches_library(lib):
    lib.mvmc_sizeof_step_out.restype = ctypes.c_size_t
    assert lib.mvmc_sizeof_step_out() == _lib.STEP_OUT_DTYPE.itemsize
    assert _lib.TRACK_OUT_DTYPE.fields["param"][1] % 8 == 0


def test_rand_stream_is_numpy_randomstate0(lib):
    n = 288 * 64
    out = np.empty(n)
    assert lib.mvmc_rand_stream_host(ctypes.c_void_p(out.ctypes.data), n) == 0
    assert np.array_equal(out, np.random.RandomState(0).rand(n))


def test_error_strings_and_version(lib):
    lib.mvmc_error_string.restype = ctypes.c_char_p
    assert lib.mvmc_version() >= 100
    assert lib.mvmc_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4):
        assert lib.mvmc_error_string(code) not in (b"ok", b"unknown error")


def test_invalid_arguments_are_rejected_before_any_cuda_call(lib):
    null = ctypes.c_void_p(None)
    assert lib.mvmc_fundamental(null, null, 1, 5, null) == _lib.ERR_INVALID
    assert lib.mvmc_rand_stream_host(null, 4) == _lib.ERR_INVALID
    assert lib.mvmc_fk(null, 3, null, null) == _lib.ERR_INVALID
    buf = (ctypes.c_double * 16)()
    assert lib.mvmc_fundamental(buf, buf, 1, _lib.MAX_VIEWS + 1, null) == _lib.ERR_INVALID
    assert lib.mvmc_triangulate(buf, buf, buf, 1, _lib.MAX_SEL + 1, 18, ctypes.c_double(0.01), 0, buf, null) == _lib.ERR_INVALID


def test_default_config_is_the_reference_constants(lib):
    cfg = _lib.Config()
    lib.mvmc_default_config(ctypes.byref(cfg))
    # MvTracklet: n_inits=3, max_age=0 (motion_capture.py:319-320); PoseSolver max_nfev 5/50 (inverse_kinematics.py:397,400)
    assert (cfg.n_inits, cfg.max_age, cfg.nfev_update, cfg.nfev_birth) == (3, 0, 5, 50)


def test_no_device_is_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cfg = _lib.Config()
    lib.mvmc_default_config(ctypes.byref(cfg))
    h = ctypes.c_void_p()
    rc = lib.mvmc_clips_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == _lib.ERR_NO_DEVICE and not h.value


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libmvmc.so"))
    with pytest.raises(_lib.MvmcError, match="no CPU fallback"):
        _lib.get_lib()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multiview_motion_capture_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "mvmc_oracle" not in txt and "import oracle" not in txt and "cuda_emu" not in txt.replace(
                    '#include "cuda_emu.h"', ""), f
