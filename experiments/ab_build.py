"""A/B builds of libsrb200.so with extra -D flags (experiments only): python experiments/ab_build.py NAME -DSRB_T2_STAGGER=1 ...
writes experiments/ab/libsrb200_NAME.so; select it at run time with SRB200_LIB=<path> (speaker-recognition_b200/_lib.py)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaker_recognition_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, "experiments", "ab")
    bdir = os.path.join(out_dir, "build_" + name)
    os.makedirs(bdir, exist_ok=True)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []

    def one(src):
        o = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [B._nvcc()] + ccbin + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + flags + ["-c", os.path.join(B.CSRC, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-4000:])
        return o

    with ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(one, B.SOURCES))
    out = os.path.join(out_dir, "libsrb200_%s.so" % name)
    subprocess.check_call([B._nvcc()] + ccbin + ["-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    print(out)


if __name__ == "__main__":
    main()
