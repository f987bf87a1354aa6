"""B200-native capture hot path of khanhha/multiview_motion_capture.

Cross-view association -> DLT triangulation -> temporal IK, as hand-written sm_100a CUDA kernels behind
the C-ABI in include/mvmc.h (libmvmc.so). This package is the Python host side: ctypes bindings
(`_lib`), stage wrappers (`stages`), the clip-batch pipeline (`clips`), the synthetic scene generator
(`synthetic`) and drop-in modules mirroring the reference's Python interface (`dropin/`).

There is no CPU fallback: importing `_lib.get_lib()` fails loudly when libmvmc.so has not been built.
"""
__version__ = "0.1.0"
