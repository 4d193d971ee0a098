"""Builds rustcv_b200/librcv_imgproc.so (the C-ABI library of include/rcv_imgproc.h)
in-tree with nvcc for sm_100a.  No JIT cache: the .so travels with the repo snapshot.

    python -m rustcv_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "librcv_imgproc.so")
SOURCES = ["context.cu", "hostmem.cu", "multi.cu", "tma.cu", "cvt.cu", "strip_gauss5.cu", "strip_gauss3.cu", "strip_gaussq8_k3.cu", "strip_gaussq8_k5.cu", "strip_gaussq8_k7.cu", "strip_gaussq8_k9.cu", "strip_gaussq8_k11.cu", "strip_gaussq8_k13.cu", "strip_gaussq8_k15.cu", "strip_sobel.cu", "strip_f32.cu", "strip_f32cn.cu", "strip_f32wide.cu", "strip_f2d_u8.cu", "strip_yuyv_sobel.cu", "strip_yuyv_gauss5.cu", "filter.cu", "geom.cu", "mjpeg.cu", "abi.cu"]
HEADERS = [os.path.join(CSRC, "rcv_internal.cuh"), os.path.join(CSRC, "fastdiv.h"), os.path.join(CSRC, "tma_ptx.cuh"), os.path.join(CSRC, "strip_pipeline.cuh"),
           os.path.join(CSRC, "strip_gaussq8.cuh"), os.path.join(CSRC, "strip_gaussq8_wide.cuh"), os.path.join(CSRC, "strip_f32_gather.cuh"), os.path.join(CSRC, "cvt_math.cuh"), os.path.join(HERE, "..", "include", "rcv_imgproc.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: the f32 specs fix where fused multiply-adds happen (explicit fmaf only).
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if _stale(obj, [path] + HEADERS):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
    return obj


def _source_hash() -> str:
    import hashlib

    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "rcv_imgproc.h")]:
        path = os.path.join(CSRC, f)
        if os.path.isfile(path):
            h.update(f.encode())
            h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


STAMP = LIB + ".srchash"


def build(force: bool = False, verbose: bool = False) -> str:
    # The .so travels to the GPU box without the object files (and file times may not survive the
    # copy): an up-to-date library is recognised by the hash of its sources, not by mtimes.
    want = _source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == want:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with cf.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 4, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if force or _stale(LIB, objs):
        tmp = LIB + ".tmp"
        cmd = [NVCC, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lnvjpeg_static", "-lculibos", "-lpthread", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)  # never truncate a library another process has mapped
    open(STAMP, "w").write(want + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
