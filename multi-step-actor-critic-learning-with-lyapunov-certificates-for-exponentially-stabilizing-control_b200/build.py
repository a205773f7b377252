"""Build recipe for csrc/libmsacl_b200.so (sm_100a only; nvcc cross-compiles without a GPU).

    python <package dir>/build.py [--force] [--verbose]

The library is built in-tree next to its sources so that it travels with the repo snapshot
to the GPU box; it is git-ignored (*.so).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libmsacl_b200.so")
STAMP = LIB + ".stamp"
SOURCES = ["env_step.cu", "rollout_fused.cu", "windows.cu", "targets.cu", "tc_selftest.cu", "rollout_tc.cu"]
HEADERS = ["common.cuh", "env_dynamics.cuh", "philox.cuh", "tcgen05.cuh", os.path.join(ROOT, "include", "msacl_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # dynamics must round like NumPy scalar math; GEMMs call __fmaf_rn explicitly
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-shared",
] + (["-DMSACL_TC_NS(ID)=" + os.environ["MSACL_TC_NS"]] if os.environ.get("MSACL_TC_NS") else []) + os.environ.get("MSACL_NVCC_EXTRA", "").split() + (["-DMSACL_TC_TIMING"] if os.environ.get("MSACL_TC_TIMING") else []) + (["-DMSACL_TC_WATCHDOG"] if os.environ.get("MSACL_TC_WATCHDOG") else [])   # role timers for tools/tc_timing.py


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []),
           *[os.path.join(CSRC, s) for s in SOURCES], "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
