"""Ingest: OpenPose JSON trees -> packed BODY_25 clips, without the per-frame Python json.load + pickle
round trip of the reference's `--mode prepare` (src/motion_capture.py:974-1005). The JSON text is scanned by the native
parser of libmvmc.so (mvmc_parse_openpose_files_host, one thread per file); BODY_25 -> COCO and the pose filter then run on
the device (mvmc_ingest_body25 / mvmc_prepare) when the clip is tracked.

Packed clip (`clip.npz`, the side format `--mode run` reads directly):
    kps25 [F,C,Pmax,25,3] float64, n_pose [F,C] int32, K [C,3,3], RT [C,3,4], img_wh [C,2], cams [C] (directory names)
"""
import ctypes
import json
import os
import pickle
from pathlib import Path

import numpy as np

from . import _lib
from ._lib import check


def _frame_key(path: Path):
    """The reference's sort key (src/motion_capture.py:995: int(stem.split('_')[1])), falling back to the stem."""
    parts = path.stem.split("_")
    try:
        return (0, int(parts[1]), path.stem)
    except (IndexError, ValueError):
        return (1, 0, path.stem)


def load_calib_arrays(cpath: Path):
    """(K, Rt, img_wh) of src/motion_capture.py:250-272 load_calib: .json (K, RT, imgSize) or .pkl (K, R, t; 1920x1080)."""
    cpath = Path(cpath)
    if "pkl" in cpath.suffix:
        with open(cpath, "rb") as f:
            d = pickle.load(f)
        K = np.array(d["K"], dtype=np.float64).reshape(3, 3)
        Rt = np.concatenate([np.array(d["R"], dtype=np.float64).reshape(3, 3), np.array(d["t"], dtype=np.float64).reshape(3, 1)], 1)
        return K, Rt, [1920, 1080]
    if "js" in cpath.suffix:
        with open(cpath) as f:
            js = json.load(f)
        return (np.array(js["K"], dtype=np.float64).reshape(3, 3), np.array(js["RT"], dtype=np.float64).reshape(3, 4),
                list(js["imgSize"]))
    raise ValueError(f"unsupported calibration format. {cpath.name}")


def parse_openpose_files(paths, max_people, threads=None):
    """paths: list of OpenPose JSON files -> (kps25 [n,max_people,25,3] float64, n_people [n] int32)."""
    lib = _lib.get_lib()
    n = len(paths)
    out = np.zeros((n, max_people, 25, 3))
    cnt = np.zeros(n, dtype=np.int32)
    if n == 0:
        return out, cnt
    arr = (ctypes.c_char_p * n)(*[os.fsencode(str(p)) for p in paths])
    check(lib.mvmc_parse_openpose_files_host(ctypes.cast(arr, ctypes.c_void_p), n, max_people, out.ctypes.data, cnt.ctypes.data,
                                             int(threads or min(32, os.cpu_count() or 1))), "mvmc_parse_openpose_files_host")
    return out, cnt


def parse_openpose_text(text: bytes, max_people):
    lib = _lib.get_lib()
    out = np.zeros((max_people, 25, 3))
    n = ctypes.c_int(0)
    check(lib.mvmc_parse_openpose_host(text, len(text), max_people, out.ctypes.data, ctypes.addressof(n)), "mvmc_parse_openpose_host")
    return out, n.value


def pack_openpose_clip(opn_kps_dir, calib_dir, max_people=None, threads=None):
    """The reference's prepare-mode inputs (one sub-directory of `*_keypoints.json` per camera, sorted by stem; calibration
    files matched by stem, any extension) -> packed clip dict."""
    from ._lib import MAX_POSES
    cam_dirs = sorted([d for d in Path(opn_kps_dir).glob("*") if d.is_dir()], key=lambda p: p.stem)
    calib_paths = {c.stem: c for c in Path(calib_dir).glob("*.*")}
    cal = [load_calib_arrays(calib_paths[d.stem]) for d in cam_dirs]
    per_cam = [sorted(d.glob("*.json"), key=_frame_key) for d in cam_dirs]
    F = min(len(p) for p in per_cam)
    C = len(cam_dirs)
    Pm = max_people or MAX_POSES
    flat = [per_cam[c][f] for f in range(F) for c in range(C)]
    kps25, cnt = parse_openpose_files(flat, Pm, threads)
    if (cnt > Pm).any():
        raise ValueError(f"a frame holds {int(cnt.max())} people, more than max_people={Pm}")
    pmax = max(int(cnt.max()), 1) if max_people is None else Pm
    return dict(kps25=np.ascontiguousarray(kps25.reshape(F, C, Pm, 25, 3)[:, :, :pmax]), n_pose=cnt.reshape(F, C),
                K=np.stack([c[0] for c in cal]), RT=np.stack([c[1] for c in cal]),
                img_wh=np.array([c[2] for c in cal], dtype=np.int32), cams=np.array([d.stem for d in cam_dirs]))


def save_clip_npz(path, clip):
    np.savez_compressed(path, **clip)


def load_clip_npz(path):
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}
