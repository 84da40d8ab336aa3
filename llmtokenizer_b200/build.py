"""Build every native artefact in-tree (nothing goes to a JIT cache):

  llmtokenizer_b200/libbpe_cuda.so   sm_100a engine + C ABI (include/bpe_cuda.h)          [nvcc]
  llmtokenizer_b200/libbpe_corpus.so synthetic corpus generator (tools/gen_corpus.c)      [gcc]
  llmtokenizer_b200/dropin/libbpe.so the reference's C API (bpe.h) on top of the engine   [gcc]
  oracle/_build/*                    CPU oracle (test infrastructure)                     [gcc]
  oracle/_ref/*                      the unmodified reference, only where /root/reference exists
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, **kw)


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_engine(force=False):
    src = os.path.join(PKG, "csrc", "bpe_engine.cu")
    csrc = os.path.join(PKG, "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "bpe_cuda.h"))
    out = os.path.join(PKG, "libbpe_cuda.so")
    if force or _stale(out, deps):
        _run([NVCC, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-o", out, src, "-ldl"])
    return out


def build_corpus(force=False):
    src = os.path.join(ROOT, "tools", "gen_corpus.c")
    out = os.path.join(PKG, "libbpe_corpus.so")
    if force or _stale(out, [src]):
        _run(["gcc", "-O2", "-fPIC", "-shared", "-DGEN_CORPUS_LIB", "-o", out, src, "-lm"])
    return out


def build_dropin(force=False):
    d = os.path.join(PKG, "dropin")
    mk = os.path.join(d, "Makefile")
    if os.path.exists(mk):
        _run(["make", "-s", "-C", d] + (["-B"] if force else []))


def build_oracle():
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    if os.path.isdir("/root/reference/bpe"):
        _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])


def build_all(force=False):
    build_engine(force)
    build_corpus(force)
    build_dropin(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
