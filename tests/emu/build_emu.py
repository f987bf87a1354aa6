"""TEST INFRASTRUCTURE: (re)builds tests/emu/libmvmc_emu.so — the real csrc/*.cu kernel sources compiled with g++
against the CUDA execution-model emulator in cuda_emu.h — when it is stale."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "multiview_motion_capture_b200", "csrc")
OUT = os.path.join(HERE, "libmvmc_emu.so")


def build_emulator(verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps += [os.path.join(ROOT, "include", "mvmc.h"), os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "build_emu.sh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.run(["bash", os.path.join(HERE, "build_emu.sh")], check=True,
                       stdout=None if verbose else subprocess.DEVNULL)
    return OUT


if __name__ == "__main__":
    print(build_emulator(verbose=True))
