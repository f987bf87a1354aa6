"""TEST INFRASTRUCTURE: runs the drop-in CLI (dropin/motion_capture.py) as `__main__` with the kernel-emulator build of
the C-ABI bound instead of libmvmc.so, so the CPU test tier can exercise the CLI and its file formats without a GPU.
The product has no such switch: `python dropin/motion_capture.py` binds the CUDA library or fails."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from multiview_motion_capture_b200 import _lib  # noqa: E402

_lib.use_library(os.path.join(ROOT, "tests", "emu", "libmvmc_emu.so"), device="cpu")
script = os.path.join(ROOT, "multiview_motion_capture_b200", "dropin", "motion_capture.py")
sys.argv = [script] + sys.argv[1:]
runpy.run_path(script, run_name="__main__")
