"""Build libd2s_b200.so in-tree with nvcc for sm_100a (the only supported target).

    python dense2sparse-vit_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_NAME = "libd2s_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
STAMP = os.path.join(PKG_DIR, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return cand


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _extra_flags():
    """Extra nvcc flags for profiling builds, e.g. D2S_NVCC_EXTRA=-DD2S_GEMM_TRACE_BUILD (clock64 phase traces)."""
    return os.environ.get("D2S_NVCC_EXTRA", "").split()


def _fingerprint():
    h = hashlib.sha256()
    h.update(" ".join(_extra_flags()).encode())
    files = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        [os.path.join(PKG_DIR, "..", "include", "d2s.h"), os.path.abspath(__file__)]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())   # not the absolute path: the stamp travels with the snapshot
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == fp:
                return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    obj_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + _extra_flags() + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(src)} (rc={p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed; see output above")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as fh:
        fh.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
