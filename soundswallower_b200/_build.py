"""Build libssb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libssb200.so")
SOURCES = ["model.cpp", "lexicon.cpp", "gmm_topn.cu", "gmm_topn_tc.cu", "gmm_topn_tc2.cu", "senone_mix.cu", "cont_score.cu", "chain_viterbi.cu", "fsg_search.cu", "frontend.cu", "api.cu"]
HEADERS = ["model.h", "device.cuh", "tc_common.cuh", "hmm_step.cuh", os.path.join("..", "..", "include", "ssb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libssb200.so cannot be built")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    """Compile every CUDA source of the package into soundswallower_b200/libssb200.so."""
    if not force and not stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
