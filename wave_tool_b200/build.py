"""In-tree build of libwavetool_b200.so (nvcc, sm_100a only).  No JIT cache: the .so sits next to
this file so it travels to the GPU box with the repository snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libwavetool_b200.so")
# the production kernel's instantiations are split over four translation units so that they compile in parallel
SOURCES = ["wt_device.cu", "wt_launch_exact.cu", "wt_launch_exact_shade.cu", "wt_launch_recur.cu", "wt_launch_recur_shade.cu", "wt_compose.cu", "wt_sky.cu", "wt_host.cpp"]
HEADERS = ["wt_kernels.cuh", "wt_aux_kernels.cuh", "wt_launch.cuh", "wt_ctx.h", "wt_internal.h", "wt_host_math.h",
           os.path.join("..", "..", "include", "wavetool_b200.h")]
OBJ_DIR = os.path.join(PKG_DIR, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # host code: position independent, and NO fp contraction — the CPU prologue reproduces the
    # reference's fp32 operation order (wt_host_math.h)
    "-Xcompiler", "-fPIC,-ffp-contract=off",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def compile_all(out_path: str, extra_flags=(), verbose: bool = False, obj_dir: str = OBJ_DIR) -> str:
    """nvcc -c every translation unit (in parallel), then link them into one shared library.  Returns ptxas' log."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); nvcc should use the distro g++
    common = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-ccbin", "/usr/bin/g++"]

    def one(src):
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        res = subprocess.run(common + ["-c", "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        done = list(pool.map(one, SOURCES))
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", out_path] + [o for o, _ in done],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return "".join(log for _, log in done)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA translation unit for sm_100a into wave_tool_b200/libwavetool_b200.so."""
    if not force and not needs_build():
        return LIB_PATH
    log = compile_all(LIB_PATH, verbose=verbose)
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
