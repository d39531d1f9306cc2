"""Build lib3dspa_b200.so (sm_100a only) in-tree with nvcc.  No GPU needed to compile."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
TAG = os.environ.get("SPA3D_BUILD_TAG", "")   # development: a second library next to the product one (A/B with SPA3D_LIB_PATH)
BUILD = os.path.join(HERE, "_build" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "lib3dspa_b200" + ("_" + TAG if TAG else "") + ".so")

SOURCES = [
    "api.cu",
    "elementwise.cu",
    "lift.cu",
    "gemm_simt.cu",
    "gemm_tcgen05.cu",
    "embed_tcgen05.cu",
    "mlp_fused_tcgen05.cu",
    "attention_simt.cu",
    "attention_tc.cu",
    "attention_tc_bwd.cu",
    "attention_q1.cu",
    "attention_api.cu",
    "host_pack.cc",      # host-only (g++): float32 -> bfloat16 staging of the end-to-end path
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--fmad=true",
]
NVCC_FLAGS += os.environ.get("SPA3D_NVCC_EXTRA", "").split()   # development switches, e.g. -DSPA3D_ATTN_TRACE


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(path, extra=""):
    h = hashlib.sha256()
    h.update(extra.encode())
    for dep in [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "..", "include", "spa3d_b200.h")
    ]:
        with open(dep, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".sha"
    dig = _digest(os.path.join(CSRC, src), " ".join(NVCC_FLAGS))
    if os.path.isfile(obj) and os.path.isfile(stamp) and open(stamp).read() == dig:
        return obj, False
    if src.endswith(".cc"):
        cmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-Wall", "-pthread", "-c", os.path.join(CSRC, src), "-o", obj]
    else:
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"compile failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, True


def build(verbose=False, force=False):
    """Compile every CUDA source for sm_100a and link the C-ABI shared library."""
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in results]
    if any(changed for _, changed in results) or not os.path.isfile(LIB):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
