"""Build the sm_100a shared library `libsparseconv_b200.so` in-tree with nvcc.

Run as `python -m nerf_downstream_b200.build` or through `__graft_entry__.build()`.
The library is the C-ABI boundary declared in `include/sparseconv_b200.h`.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libsparseconv_b200.so"
STAMP = PKG_DIR / "build" / "stamp.txt"

SOURCES = ["coords.cu", "elementwise.cu", "conv_simt.cu", "conv_umma.cu", "conv_umma_pair.cu", "conv_wgrad_umma.cu", "conv_wgrad_umma_pair.cu", "conv_api.cu", "pipeline.cu", "head.cu", "interp.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "umma_common.cuh", "../../include/sparseconv_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> Path:
    """`experiments=True`: the timing-experiment variant (-DSPC_EXPERIMENTS: debug knobs of the convolution kernels,
    multi-chunk stages) as libsparseconv_b200_exp.so — used by scripts/ only (SPARSECONV_B200_LIB), never shipped as
    the product library."""
    if experiments:
        return _build_to(PKG_DIR / "libsparseconv_b200_exp.so", PKG_DIR / "build_exp", ["-DSPC_EXPERIMENTS"], verbose)
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    _build_to(LIB_PATH, PKG_DIR / "build", [], verbose)
    STAMP.write_text(digest)
    return LIB_PATH


def _build_to(lib_path: Path, obj_dir: Path, extra_flags, verbose: bool) -> Path:
    nvcc = _nvcc()
    obj_dir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = obj_dir / (src + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [nvcc, "-shared", "-o", str(lib_path), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-cudart", "static"]
    subprocess.run(cmd, check=True)
    return lib_path


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv)
    print(path)
