"""Build the sm_100a CUDA library in-tree (stemseg_b200/libstemseg_b200.so).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
Usage: python -m stemseg_b200.build [--force]
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libstemseg_b200.so")
STAMP_PATH = os.path.join(PKG_DIR, ".libstemseg_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    inc = os.path.join(os.path.dirname(PKG_DIR), "include", "stemseg_b200.h")
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    root = os.path.dirname(PKG_DIR)
    for path in files + [inc]:
        with open(path, "rb") as f:
            h.update(os.path.relpath(path, root).encode())        # checkout-location independent
            h.update(f.read())
    return h.hexdigest()


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libstemseg_b200.so")
    return nvcc


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. Returns the library path."""
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH):
        with open(STAMP_PATH) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    nvcc = find_nvcc()
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (os.path.basename(src), out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(objdir, "nvcc.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(log)[-8000:])
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout)
    with open(STAMP_PATH, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
