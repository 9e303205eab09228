"""In-tree build of libst_b200.so (hand-written sm_100a CUDA, plain C ABI, no torch headers).

    python speech-tranformer-pytorch_b200/build.py [--force]

nvcc cross-compiles for sm_100a without a GPU; the resulting .so sits next to this file so it
travels with the working tree to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libst_b200.so")

SOURCES = ["st_host.cu", "st_gemm.cu", "st_gemm_h.cu", "st_gemm_bf.cu", "st_ln.cu", "st_lsce.cu", "st_attn.cu", "st_attn_bwd.cu", "st_attn16.cu", "st_optim.cu", "st_embed.cu", "st_beam.cu", "st_ctc.cu", "st_selftest.cu", "st_nccl.cu", "st_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _deps_mtime() -> float:
    t = os.path.getmtime(os.path.join(INCLUDE, "st_b200.h"))
    for f in os.listdir(CSRC):
        if f.endswith((".h", ".cuh")):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def _compile(src: str, hdr_mtime: float, force: bool) -> str:
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(path), hdr_mtime)):
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link libst_b200.so. Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdr = _deps_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(lambda s: _compile(s, hdr, force), sources))
    if (force or not os.path.exists(LIB_PATH)
            or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(o) for o in objs)):
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"linked {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
