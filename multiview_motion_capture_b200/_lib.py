"""ctypes binding of libmvmc.so (include/mvmc.h). No fallback: a missing library is an error."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_size_t, c_uint8, c_uint32, c_ulonglong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmvmc.so")

MAX_VIEWS, MAX_POSES, MAX_TRACKS, N_COCO, N_B18, N_PARAM, MAX_SEL = 8, 32, 64, 17, 18, 68, 16
MAX_GROUP, MAX_BIG = 256, 8
OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_NO_DEVICE = 0, -1, -2, -3, -4


class Config(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("n_clips", "n_views", "max_poses", "max_tracks", "max_new", "n_inits", "max_age",
                                     "nfev_update", "nfev_birth", "keep_matrices")]


TRACK_OUT_DTYPE = np.dtype([
    ("track_id", np.int32), ("state", np.int32), ("hits", np.int32), ("time_since_update", np.int32),
    ("length", np.int32), ("updated", np.int32), ("n_sel", np.int32), ("sel", np.int32, (MAX_SEL, 2)),
    ("nfev", np.int32, (2,)), ("njev", np.int32, (2,)), ("status", np.int32, (2,)), ("pad_", np.int32),
    ("cost", np.float64, (2,)), ("param", np.float64, (N_PARAM,)), ("joints", np.float64, (N_B18 * 3,)),
], align=True)

STEP_OUT_DTYPE = np.dtype([
    ("frame_idx", np.int32), ("n_alive", np.int32), ("n_died", np.int32), ("died_ids", np.int32, (MAX_TRACKS,)),
    ("n_total", np.int32), ("als_iters", np.int32), ("n_dup_view", np.int32), ("error", np.int32),
    ("n_truncated", np.int32),
    ("tracks", TRACK_OUT_DTYPE, (MAX_TRACKS,)),
], align=True)

_P = c_void_p  # every array argument is passed as a raw address

_SIGNATURES = {
    "mvmc_version": (c_int, []),
    "mvmc_error_string": (c_char_p, [c_int]),
    "mvmc_last_cuda_error": (c_char_p, []),
    "mvmc_rand_stream_host": (c_int, [_P, c_int]),
    "mvmc_fundamental": (c_int, [_P, _P, c_int, c_int, _P]),
    "mvmc_fundamental_krt": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "mvmc_prepare": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "mvmc_affinity": (c_int, [_P] * 9 + [c_int] * 4 + [_P, _P, _P]),
    "mvmc_match_als_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mvmc_match_als": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "mvmc_assign": (c_int, [_P] * 5 + [c_int] * 5 + [_P] * 7 + [_P]),
    "mvmc_assign_listed": (c_int, [_P] * 5 + [c_int] * 5 + [_P] * 9 + [_P]),
    "mvmc_distances": (c_int, [_P] * 8 + [c_int] * 4 + [_P, _P]),
    "mvmc_linear_sum_assignment": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "mvmc_match_views_workspace_bytes": (c_size_t, [c_int]),
    "mvmc_match_views_hungarian": (c_int, [_P, _P, c_int, c_int, c_int, c_double, _P, _P, _P, _P, _P]),
    "mvmc_tracklet_pose_association": (c_int, [_P] * 6 + [c_int] * 4 + [c_double, _P, _P, _P, _P]),
    "mvmc_assign_groups": (c_int, [_P] * 5 + [c_int] * 5 + [_P] * 13 + [_P]),
    "mvmc_transform_closure": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "mvmc_triangulate": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_double, c_int, _P, _P]),
    "mvmc_fk": (c_int, [_P, c_int, _P, _P]),
    "mvmc_fk_chain": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P]),
    "mvmc_ik_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mvmc_ik_solve": (c_int, [_P] * 7 + [c_int, c_int] + [_P] * 6),
    "mvmc_ik_birth_big_workspace_bytes": (c_size_t, []),
    "mvmc_ik_birth_big": (c_int, [_P] * 6 + [c_int] * 7 + [_P] * 5 + [_P]),
    "mvmc_ik_solve_targets": (c_int, [_P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P]),
    "mvmc_default_config": (None, [POINTER(Config)]),
    "mvmc_clips_create": (c_int, [POINTER(Config), POINTER(c_void_p)]),
    "mvmc_clips_destroy": (None, [c_void_p]),
    "mvmc_clips_device_bytes": (c_size_t, [c_void_p]),
    "mvmc_clips_set_calib": (c_int, [c_void_p, _P, _P, _P, _P]),
    "mvmc_clips_reset": (c_int, [c_void_p, _P]),
    "mvmc_clips_step": (c_int, [c_void_p, _P, _P, c_int, _P]),
    "mvmc_sizeof_step_out": (c_size_t, []),
    "mvmc_clips_last_out": (c_void_p, [c_void_p]),
    "mvmc_clips_pack_records": (c_int, [c_void_p, c_int, c_int, _P, _P, _P]),
    "mvmc_clips_step_host": (c_int, [c_void_p, _P, _P, c_int, _P, _P]),
    "mvmc_clips_step_host_async": (c_int, [c_void_p, _P, _P, c_int, _P, _P]),
    "mvmc_parse_openpose_host": (c_int, [c_char_p, c_size_t, c_int, _P, _P]),
    "mvmc_parse_openpose_files_host": (c_int, [_P, c_int, c_int, _P, _P, c_int]),
    "mvmc_ingest_body25": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "mvmc_clips_step_body25_host": (c_int, [c_void_p, _P, _P, c_int, _P, _P]),
    "mvmc_clips_set_tracks_host": (c_int, [c_void_p] + [_P] * 9 + [_P]),
    "mvmc_clips_read_big_groups_host": (c_int, [c_void_p, c_int, _P, _P, _P, _P, _P]),
    "mvmc_clips_read_matrices_host": (c_int, [c_void_p, c_int, _P, _P, _P, _P, _P, _P]),
    "mvmc_clips_stats_host": (c_int, [c_void_p, _P, c_int, _P]),
    "mvmc_clips_profile": (c_int, [c_void_p, c_int, _P, _P, _P]),
    "mvmc_fp64_probe": (c_int, [c_int, c_int, _P, _P]),
    "mvmc_fp64_tensor_probe": (c_int, [c_int, c_int, _P, _P]),
    "mvmc_als_phase_profile": (c_int, [c_int, _P]),
    "mvmc_als_force_variant": (c_int, [c_int]),
    "mvmc_launch_count": (c_ulonglong, []),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lib_path = None
_device = None


class MvmcError(RuntimeError):
    pass


def _bind(lib):
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mvmc_sizeof_step_out() != STEP_OUT_DTYPE.itemsize:
        raise MvmcError(f"mvmc_step_out layout mismatch: C {lib.mvmc_sizeof_step_out()} vs numpy {STEP_OUT_DTYPE.itemsize}")
    return lib


def use_library(path, device=None):
    """Bind an explicit build of the C-ABI (same header, same symbols) and, optionally, the torch device its buffers
    live on. The product never calls this; get_lib() binds the in-tree CUDA build and the device is CUDA."""
    global _lib, _lib_path, _device
    _lib = _bind(ctypes.CDLL(path))
    _lib_path = path
    _device = device
    return _lib


def get_lib():
    """The CUDA library. Raises if it has not been built (run `python __graft_entry__.py` / build())."""
    global _lib, _lib_path
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MvmcError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                            f"(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        _lib = _bind(ctypes.CDLL(LIB_PATH))
        _lib_path = LIB_PATH
    return _lib


def lib_path():
    return _lib_path


TORCH_EXT_PATH = os.path.join(_HERE, "lib", "libmvmc_torch.so")
_ops = None


def torch_ops():
    """`torch.ops.mvmc`: the PyTorch C++ extension over the C-ABI (csrc/torch_ext.cpp), loaded next to the in-tree CUDA
    library. None when another build of the C-ABI was bound explicitly (use_library). Raises if the extension is missing:
    build() produces both files."""
    global _ops
    if _ops is None:
        get_lib()
        if _lib_path != LIB_PATH:
            return None
        if not os.path.exists(TORCH_EXT_PATH):
            raise MvmcError(f"{TORCH_EXT_PATH} is missing: run __graft_entry__.build()")
        import torch
        torch.ops.load_library(TORCH_EXT_PATH)
        _ops = torch.ops.mvmc
    return _ops


def default_device():
    """torch device of the library's buffers: the current CUDA device (raises without one - no CPU fallback)."""
    import torch
    get_lib()
    if _device is not None:
        return torch.device(_device)
    if not torch.cuda.is_available():
        raise MvmcError("no CUDA device: the capture path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def check(rc, what=""):
    if rc != OK:
        lib = get_lib()
        msg = lib.mvmc_error_string(rc).decode()
        if rc == ERR_CUDA:
            msg += " — " + lib.mvmc_last_cuda_error().decode()
        raise MvmcError(f"{what}: {msg} ({rc})")


def ptr(t):
    """Raw address of a torch tensor / numpy array (must be contiguous), or None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return t.ctypes.data
    assert t.is_contiguous()
    return t.data_ptr()
