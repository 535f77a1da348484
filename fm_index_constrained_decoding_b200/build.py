"""Builds libfmidx.so (CUDA kernels + C ABI) IN-TREE for sm_100a with nvcc.

    python -m fm_index_constrained_decoding_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels to the GPU box
with the repo snapshot.  There is deliberately no other backend: if this library cannot be built
or loaded, the package raises.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libfmidx.so")
SOURCES = ["fmidx_build.cu", "fmidx_kernels.cu", "fmidx_capi.cu", "fmidx_shard.cu"]
HEADERS = ["fmidx_layout.cuh", "fmidx_internal.h", os.path.join("..", "..", "include", "fmidx.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    todo = []
    for s in SOURCES:
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, s)] + hdrs):
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=len(todo)) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or todo or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
