"""In-tree build of libspecdec_b200.so (nvcc, sm_100a only; cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SD_LIB_VARIANT=prof: a second copy of the library with the ring kernel's in-kernel timeline probes compiled in
# (-DSD_RING_PROF; tools/ring_prof.py) — the production library does not carry them
VARIANT = os.environ.get("SD_LIB_VARIANT", "")
LIB = os.path.join(HERE, "libspecdec_b200.so" if not VARIANT else f"libspecdec_b200_{VARIANT}.so")
SOURCES = ["norm.cu", "norm_pipe.cu", "norm_pipe_f32.cu", "norm_pipe_bf16.cu", "norm_pipe_f16.cu", "norm_ring.cu", "norm_ring_f32.cu",
           "norm_ring_bf16.cu", "norm_ring_f16.cu", "verify.cu", "misc.cu", "api.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "specdec_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into llmspeculativesampling_b200/libspecdec_b200.so.  Returns the path."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build" if not VARIANT else f"build_{VARIANT}")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *(["-DSD_RING_PROF"] if VARIANT == "prof" else []), *os.environ.get("SD_EXTRA_NVCC_FLAGS", "").split(),
               "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
