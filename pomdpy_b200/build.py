"""Compile the CUDA extension in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpomcp_b200.so")
SOURCES = ["pomcp_kernels.cu", "pomcp_refexact.cu", "vi_kernels.cu", "vi_tcgen05.cu"]
HEADERS = ["pomcp_device.cuh", "pomcp_state.cuh", "pomcp_models.cuh", "pomcp_plan.cuh",
           os.path.join("..", "..", "include", "pomcp_b200.h"),
           os.path.join("..", "..", "include", "vi_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


CHECKS_LIB = os.path.join(HERE, "libpomcp_b200_checks.so")


def build(force=False, verbose=False, checks=False, defines=(), out=None):
    """Build libpomcp_b200.so next to this package; returns its path.  checks=True builds the variant with
    device-side bounds checks (-DPOMCP_CHECKS) as libpomcp_b200_checks.so; `defines` + `out` build an experiment
    variant (e.g. defines=("POMCP_TSEC",) for the expand-phase section timers) to load through POMCP_B200_LIB."""
    variant = checks or bool(defines)
    out = out or (CHECKS_LIB if checks else LIB)
    if not force and not variant and not stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DPOMCP_CHECKS"] if checks else []) + ["-D" + d for d in defines] + \
          (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build(force=True, checks=True))
