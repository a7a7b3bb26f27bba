"""In-tree build of the C-ABI CUDA library (sm_100a only).  `python -m pnode_b200.build [--force] [-v]`.

Every .cu is compiled to an object under csrc/_obj/ (in parallel, only when stale) and the objects are linked into
csrc/libpnode_b200.so, which travels with the tree (git-ignored, not gpurun-ignored)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SOURCES = ["vecops.cu", "mlp_rk.cu", "cnf_rk.cu", "bn_relu.cu", "conv_block.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "pnode_b200.h")]
LIB = os.path.join(CSRC, "libpnode_b200.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def _mtime(path):
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _compile(nvcc, src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr))
    if verbose:
        print(proc.stderr)
    return obj


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)
    objs, todo = [], []
    for src in SOURCES:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _mtime(obj) < max(_mtime(os.path.join(CSRC, src)), hdr_t):
            todo.append(src)
    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            list(ex.map(lambda s: _compile(nvcc, s, verbose), todo))
    if todo or _mtime(LIB) < max(_mtime(o) for o in objs):
        cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), proc.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
