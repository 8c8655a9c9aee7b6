"""Build tests/_build/libhostcheck.so: the product's HOST sources + the C oracle as stage backend (CPU tests only)."""
import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "miniprot_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build", "libhostcheck.so")
HOST_SRCS = ["tables.cpp", "ntdb.cpp", "index.cpp", "hits.cpp", "align.cpp", "paf.cpp", "annot.cpp", "pipeline.cpp"]


def build(force=False):
    if os.environ.get("MPB_HOSTCHECK_SO"):  # a build made elsewhere, e.g. with -fsanitize=address,undefined (run python under LD_PRELOAD=libasan)
        return os.environ["MPB_HOSTCHECK_SO"]
    srcs = [os.path.join(CSRC, s) for s in HOST_SRCS] + [os.path.join(ROOT, "tests", "hostcheck", f) for f in ("hostcheck.cpp", "emu_nasw.cpp", "emu_chain.cpp")]
    ora = sorted(glob.glob(os.path.join(ROOT, "oracle", "*.c")))
    deps = srcs + ora + glob.glob(os.path.join(CSRC, "*.hpp")) + glob.glob(os.path.join(CSRC, "cuda", "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + \
        glob.glob(os.path.join(ROOT, "oracle", "*.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    for c in ora:
        o = os.path.join(os.path.dirname(OUT), "ora_" + os.path.basename(c) + ".o")
        subprocess.run(["gcc", "-std=c11", "-O2", "-g", "-fPIC", "-c", c, "-o", o], check=True)
        objs.append(o)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-I" + os.path.join(ROOT, "include"),
           "-I" + CSRC, "-I" + os.path.join(ROOT, "oracle"), "-o", OUT] + srcs + objs + ["-lz", "-lpthread", "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
