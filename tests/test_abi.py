"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/ds2i_gpu.h declares; without a CUDA device it fails loudly (no fallback)."""
import ctypes as C
import os
import re

import pytest

from util import GOLDEN, ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "ds2i_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ds2i_gpu_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(native_lib):
    from ds2i_b200 import _native
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(native_lib, n), "libds2i_gpu.so does not export " + n
    assert sorted(_native.SYMBOLS) == names


def test_op_names(native_lib):
    ops = ["and", "and_freq", "or", "or_freq", "ranked_and", "wand", "maxscore", "ranked_or"]
    assert [native_lib.ds2i_gpu_op_from_name(o.encode()) for o in ops] == list(range(8))
    assert native_lib.ds2i_gpu_op_from_name(b"nonsense") < 0


def test_no_cpu_fallback(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = native_lib.ds2i_gpu_index_open_file(os.path.join(GOLDEN, "mini.block_optpfor.idx").encode(), b"block_optpfor", 0, C.byref(h))
    assert rc == -3 and h.value is None        # DS2I_E_CUDA: fails loudly, nothing is computed on the CPU
    assert b"CUDA" in native_lib.ds2i_gpu_last_error() or b"cuda" in native_lib.ds2i_gpu_last_error()


def test_cpp_adapters_compile_as_cpp11(native_lib, tmp_path):
    """include/ds2i_gpu.hpp is what a ds2i maintainer includes (the reference builds with -std=c++11, CMakeLists.txt:9)."""
    import subprocess
    from ds2i_b200 import build
    exe = tmp_path / "enumerator_walk"
    r = subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                        os.path.join(ROOT, "tests", "cpp", "enumerator_walk.cpp"), "-L", build.LIBDIR, "-lds2i_gpu", "-Wl,-rpath," + build.LIBDIR],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
