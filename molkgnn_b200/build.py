"""In-tree build of the CUDA extension (sm_100a only): nvcc -> molkgnn_b200/libmolkgnn_b200.so."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmolkgnn_b200.so")
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--shared"]


def sources():
    return sorted(glob.glob(os.path.join(SRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(SRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "molkgnn_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False, phase_clocks=False):
    """phase_clocks=True builds the profiling variant libmolkgnn_b200_prof.so (-DMK_PHASE_CLOCKS: in-kernel phase timers
    of the tile kernels, tools/phase_clocks.py); it is never loaded unless MOLKGNN_B200_LIB points at it."""
    out = LIB.replace(".so", "_prof.so") if phase_clocks else LIB
    if not force and not phase_clocks and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = ([nvcc] + NVCC_FLAGS + (["-DMK_PHASE_CLOCKS"] if phase_clocks else []) + (["-Xptxas", "-v"] if verbose else [])
           + sources() + ["-o", out])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libmolkgnn_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, phase_clocks="--phase-clocks" in sys.argv))
