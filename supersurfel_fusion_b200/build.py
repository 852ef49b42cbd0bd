"""Builds libssf.so (the sm_100a CUDA kernels + C-ABI) in-tree with nvcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libssf.so")
SOURCES = ["ssf_icp.cu", "ssf_surfels.cu", "ssf_tps.cu", "ssf_ingest.cu", "ssf_engine.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # decision arithmetic must round as written (see csrc/ssf_math.cuh)
    "-Xcompiler", "-fPIC",
    "-cudart", "static",
    "--threads", "4",
]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ssf.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into supersurfel_fusion_b200/libssf.so."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB_PATH] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
