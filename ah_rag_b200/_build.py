"""In-tree build of libahrag_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build() and on first import
when the library is missing or older than its sources.  The .so is git-ignored but travels to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
DEBUG = os.environ.get("AHR_LIB_DEBUG", "0") == "1"  # the build with device-side invariant checks (common.cuh)
LIB_PATH = os.path.join(LIB_DIR, "libahrag_b200_dbg.so" if DEBUG else "libahrag_b200.so")
SOURCES = ["api.cu", "topk_tc.cu", "topk_tc2.cu", "topk_exact.cu", "select.cu", "graph.cu", "env.cu", "exchange.cu", "lca.cu", "graph_shard.cu"]
HEADERS = ["common.cuh", "kernels.h", "tc_shared.cuh", "graph_shared.cuh", os.path.join("..", "..", "include", "ahrag_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libahrag_b200.so must be built where the CUDA toolkit is installed")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile and link under an inter-process lock (every rank of a torchrun job may find the library stale at the
    same moment); the library is linked to a temporary name and renamed into place, so no process ever maps a
    half-written file."""
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while we waited
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    obj_dir = os.path.join(HERE, "build_dbg" if DEBUG else "build")
    os.makedirs(obj_dir, exist_ok=True)
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        ccbin = ["-ccbin", "/usr/bin/g++"]
    else:
        ccbin = []

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *(["-DAHR_DEBUG_BOUNDS"] if DEBUG else []), *ccbin, "-Xptxas", "-v", "-c",
               os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        log = os.path.join(obj_dir, src + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [nvcc, "-shared", *ccbin, "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
