"""In-tree build of the native pieces (sm_100a only; nvcc cross-compiles without a GPU).

  libds2i_gpu.so   CUDA kernels + C ABI (include/ds2i_gpu.h)              <- csrc/ds2i_gpu.cu
  ds2i_build       format-compatible index builder / synthetic generator  <- csrc/builder.cpp
  queries_gpu      the `queries` front end (same argv/stdin/stdout)       <- csrc/queries_main.cpp
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "lib")
LIB = os.path.join(LIBDIR, "libds2i_gpu.so")
BUILDER = os.path.join(LIBDIR, "ds2i_build")
QUERIES = os.path.join(LIBDIR, "queries_gpu")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # BM25 must not be contracted (bit-exact vs -ffp-contract=off reference)
    "-Xcompiler", "-fPIC",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(ext):
    out = []
    for d in (CSRC, os.path.join(os.path.dirname(ROOT), "include")):
        for f in os.listdir(d):
            if f.endswith(ext):
                out.append(os.path.join(d, f))
    return out


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    deps = _sources((".cu", ".cuh", ".hpp", ".h", ".cpp"))
    nvcc = os.environ.get("NVCC", "nvcc")
    if force or _newer(LIB, deps):
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("DS2I_NVCC_EXTRA", "").split() + ["-shared", "-o", LIB, os.path.join(CSRC, "ds2i_gpu.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        out = _run(cmd)
        if verbose:
            print(out)
    bsrc = os.path.join(CSRC, "builder.cpp")
    if os.path.exists(bsrc) and (force or _newer(BUILDER, deps)):
        _run(["g++", "-O3", "-std=c++17", "-march=x86-64-v3", "-pthread", "-o", BUILDER, bsrc])
    qsrc = os.path.join(CSRC, "queries_main.cpp")
    if os.path.exists(qsrc) and (force or _newer(QUERIES, deps)):
        _run(["g++", "-O2", "-std=c++17", "-I", os.path.join(os.path.dirname(ROOT), "include"), "-o", QUERIES, qsrc,
              "-L", LIBDIR, "-lds2i_gpu", "-Wl,-rpath,$ORIGIN"])
    return LIB


def builder_writes(index_type):
    """Whether ds2i_build can write `index_type` (ds2i_build types)."""
    try:
        r = subprocess.run([BUILDER, "types"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10)
        return r.returncode == 0 and index_type in r.stdout.split()
    except Exception:
        return False


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
