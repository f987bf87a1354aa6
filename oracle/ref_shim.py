"""TEST INFRASTRUCTURE — import shim for the real reference (container only).

This module puts ``/root/reference/src`` on ``sys.path`` and imports the
reference's own modules *unmodified* so that ``oracle/make_golden.py`` can run
them and dump golden vectors into ``tests/golden/``.  It cannot travel to the GPU
box (``/root/reference`` does not exist there); nothing in ``-m gpu`` tests,
``smoke()`` or ``bench.py`` imports it.

Why a shim is needed: ``motion_capture.py`` imports GUI /
robotics packages that are absent here (matplotlib, tensorflow, imageio, pulp,
pinocchio, qpsolvers, easydict, tensorlayer) at module level
(/root/reference/src/motion_capture.py:12-25,37) and uses ``np.float`` /
``np.int`` (:672,:857).  We install permissive stub modules for those names,
re-add the two numpy aliases and swap the SciPy solver
(/root/reference/src/inverse_kinematics.py:351-433) in for the un-runnable
Pinocchio one (/root/reference/src/inverse_kinematics_pino.py).
"""
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("MVMC_REFERENCE_ROOT", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")

_STUBBED = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.gridspec",
    "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.axes3d",
    "tensorflow", "tensorlayer", "easydict", "imageio", "pulp",
    "pinocchio", "pinocchio.robot_wrapper", "pinocchio.utils", "qpsolvers",
]


class _Permissive(types.ModuleType):
    """A module whose every attribute is another permissive stub; calling it is a no-op."""

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        child = _Permissive(self.__name__ + "." + key)
        setattr(self, key, child)
        return child

    def __call__(self, *args, **kwargs):
        return None


def available() -> bool:
    return os.path.isdir(REFERENCE_SRC)


_loaded = None


def load():
    """Import the reference. Returns a namespace with ``mc`` (motion_capture), ``ik``
    (inverse_kinematics), ``mvu`` (mv_math_util), ``mva`` (mv_association), ``pose_def``,
    ``common``."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    for name in _STUBBED:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            sub = ".".join(parts[:i])
            if sub not in sys.modules:
                sys.modules[sub] = _Permissive(sub)
    sys.modules["pinocchio.utils"].__all__ = []

    # our own drop-in modules share names (common, pose_def, motion_capture) with the
    # reference's; make sure the reference's win inside this process
    for name in ("common", "pose_def", "motion_capture", "mv_math_util", "mv_association",
                 "inverse_kinematics", "Quaternions", "kinematics"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_SRC):
            del sys.modules[name]

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import motion_capture as mc
        import inverse_kinematics as ik
        import mv_math_util as mvu
        import mv_association as mva
        import pose_def
        import common
        import kinematics
        import Quaternions

    class PoseSolverAdapter(ik.PoseSolver):
        # MvTracklet passes cam_calibs= (motion_capture.py:326-331); the SciPy solver has no such arg
        def __init__(self, skeleton, init_pose, cam_poses_2d, cam_projs, cam_calibs=None, obs_kps_format=None):
            super().__init__(skeleton, init_pose, cam_poses_2d, cam_projs, obs_kps_format)

    mc.PoseSolver = PoseSolverAdapter
    ns = types.SimpleNamespace(mc=mc, ik=ik, mvu=mvu, mva=mva, pose_def=pose_def, common=common,
                               kinematics=kinematics, Quaternions=Quaternions)
    _loaded = ns
    return ns
