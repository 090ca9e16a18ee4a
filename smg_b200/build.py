"""Builds the smgx shared library (hand-written sm_100a CUDA + the C ABI) in-tree: smg_b200/libsmgx.so.

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsmgx.so")
SOURCES = ["smgx.cu", "event_index.cu", "event_kernels.cu", "tokenizer.cu", "token_tree.cu", "string_tree.cu", "blake3.cu", "prefix_hash.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # f32 imbalance / match-rate compares must not be contracted (cache_aware.rs:670, :851)
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-shared", "-cudart", "static",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "smgx.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libsmgx.so")
    if verbose:
        print(res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
