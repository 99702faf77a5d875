"""Build recipe for csrc/libsrb200.so: explicit nvcc for sm_100a only, in-tree (the .so travels with the
repo snapshot to the GPU box; a JIT cache would not).  Used by __graft_entry__.build() and `python -m
speaker_recognition_b200.build`."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libsrb200.so")
SOURCES = ["api.cu", "gmm_score.cu", "gmm_score_tc2.cu", "gmm_em.cu", "gmm_em_tc.cu", "mfcc.cu", "lpc.cu", "silence.cu", "kmeans.cu"]
HEADERS = ["common.cuh", "gmm_model.h", "kernels.h", "fft_core.h", "gmm_score_tc2.cuh", "gmm_score_tc_layout.h", os.path.join("..", "..", "include", "srb200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link libsrb200.so.  Returns the path."""
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    nvcc = _nvcc()
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper that links libstdc++ statically); use the system
    # compiler as nvcc's host so the library shares libstdc++.so.6 with the Python process
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []

    def one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(bdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + ccbin + NVCC_FLAGS + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(o.replace(".o", ".log"), "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
            if verbose:
                sys.stderr.write(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(one, SOURCES))
    if force or _stale(OUT, objs):
        cmd = [nvcc] + ccbin + ["-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return OUT


BENCH_OUT = os.path.join(CSRC, "bench", "libsrb200_bench.so")


def build_bench(force=False):
    """measurement helper for bench.py (fp32 FMA-pipe peak); separate from the product library"""
    src = os.path.join(CSRC, "bench", "fma_peak.cu")
    if force or _stale(BENCH_OUT, [src]):
        ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
        cmd = [_nvcc()] + ccbin + ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                                   "-Xcompiler", "-fPIC", "-shared", "-o", BENCH_OUT, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for fma_peak.cu:\n" + r.stderr[-4000:])
    return BENCH_OUT


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_bench(force="--force" in sys.argv))
