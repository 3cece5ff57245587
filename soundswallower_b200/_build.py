"""Build libssb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libssb200.so")
SOURCES = ["model.cpp", "lexicon.cpp", "gmm_topn.cu", "gmm_topn_tc2.cu", "gmm_scan_ft.cu", "senone_mix.cu", "cont_score.cu", "chain_viterbi.cu", "fsg_search.cu", "topn_fixup.cu", "frontend.cu", "api.cu", "search.cpp", "align_texts.cpp"]
HEADERS = ["model.h", "device.cuh", "tc_common.cuh", "hmm_step.cuh", os.path.join("..", "..", "include", "ssb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libssb200.so cannot be built")


OBJDIR = os.path.join(CSRC, "build")


def _deps():
    return [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(CSRC, "host_util.cuh")]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + _deps()
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    """Compile every CUDA source of the package into soundswallower_b200/libssb200.so:
    one object per source under csrc/build/ (only stale ones are recompiled, in parallel),
    then one link."""
    if not force and not stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = max(os.path.getmtime(d) for d in _deps() if os.path.exists(d))
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    jobs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        src_t = max(os.path.getmtime(os.path.join(CSRC, src)), hdr_t)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_t:
            jobs.append((src, subprocess.Popen([nvcc] + flags + ["-c", "-o", obj, src], cwd=CSRC)))
    failed = [src for src, p in jobs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed on " + ", ".join(failed))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                           "-Xcompiler", "-fPIC", "-o", LIB] + objs, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
