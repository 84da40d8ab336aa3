"""CPU suite: the C-ABI library loads, exports every symbol include/bpe_cuda.h declares, and fails
loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include/bpe_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bpe_cuda_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from llmtokenizer_b200 import _lib
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bpe_cuda.h but not exported"


def test_header_is_plain_c_and_cxx(tmp_path):
    """The boundary is a C ABI: include/bpe_cuda.h must compile as strict C99 (what the reference's bpe.c would include)
    and as C++ (extern "C"), with no torch / CUDA types in any signature, and link against the library."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "use.c"
    src.write_text('#include "bpe_cuda.h"\n'
                   "int main(void) { bpe_pair_t p = {1u, 2u}; bpe_cuda_stats_t st; (void)p; (void)st;\n"
                   "  return bpe_cuda_device_count() < 0 || bpe_cuda_last_error() == 0; }\n")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", str(src)])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", inc, "-x", "c++", "-fsyntax-only", str(src)])
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, "bpe_cuda.h")).read(), flags=re.S)
    assert not re.search(r"torch|at::|cudaStream_t|__global__|#include\s*<cuda", text)
    exe = tmp_path / "use"
    pkg = os.path.join(ROOT, "llmtokenizer_b200")
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, "-o", str(exe), str(src), "-L", pkg, "-lbpe_cuda", f"-Wl,-rpath,{pkg}"])
    r = subprocess.run([str(exe)], capture_output=True)
    assert r.returncode in (0, 1)          # runs and returns (no device here: count 0, error string present)


def test_no_cpu_fallback_without_device():
    import llmtokenizer_b200 as L
    from llmtokenizer_b200 import _lib
    lib = _lib.load()
    if lib.bpe_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(L.BpeCudaError) as e:
        L.train(b"abababab")
    assert e.value.rc == -4
    with pytest.raises(L.BpeCudaError):
        L.encode(b"abababab", np.array([[97, 98]], dtype=np.uint32))
    with pytest.raises(L.BpeCudaError):
        L.Context(0)


def test_argument_errors_mirror_reference():
    import llmtokenizer_b200 as L
    # compress(NULL, ..) -> NULL (bpe.c:548); unreadable file -> NULL (bpe.c:133-137)
    assert L.compress(None) is None
    assert L.compress("/nonexistent/file") is None


def test_pair_file_format_round_trip(tmp_path):
    import llmtokenizer_b200 as L
    pair_arr = np.zeros((256 + 3, 2), dtype=np.uint32)
    pair_arr[:256, 0] = np.arange(256)
    pair_arr[256:] = [[97, 98], [256, 99], [257, 257]]
    p = tmp_path / "pairs.bin"
    L.dump_pairs(str(p), pair_arr)
    assert p.stat().st_size == 3 * 8          # LE {u32 a,u32 b} records from id 256 (bpe.c:268)
    assert np.array_equal(L.read_pairs(str(p)), pair_arr)
