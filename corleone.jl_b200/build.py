"""In-tree build of libcorleone_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python corleone.jl_b200/build.py [--force] [--ptxas-v]
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, os.environ.get("CB200_OBJ_DIR", "build"))
LIB = os.path.join(HERE, os.environ.get("CB200_LIB_NAME", "libcorleone_b200.so"))
EXTRA = os.environ.get("CB200_NVCC_EXTRA", "").split()   # tuning builds, e.g. -DCB200_MINB=4 (with CB200_OBJ_DIR)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -Xfatbin -compress-all: the device code is stored compressed (the driver inflates it at module load): 3x smaller
# objects -- with the per-model instantiation lists trimmed to G in {0, default} the library is ~35 MB instead of 198 MB
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-Xfatbin", "-compress-all",
]
# -lineinfo (ncu's source page) only for the translation units whose kernels are profiled (profiles/): the BASELINE
# configs' models and the host-side kernels; it is a fifth of the size of the others
LINEINFO = {"api.cu", "comm.cu", "shoot_dense20.cu", "shoot_lqr.cu", "shoot_lotka2_oed.cu", "shoot_lotka3.cu"}
if os.environ.get("CB200_LINEINFO_ALL"):
    LINEINFO = None


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, ptxas_v: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "corleone_b200.h")]
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            li = ["-lineinfo"] if (LINEINFO is None or os.path.basename(src) in LINEINFO) else []
            cmd = [NVCC] + FLAGS + li + EXTRA + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", src, "-o", obj]
            jobs.append((src, cmd))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = {ex.submit(subprocess.run, cmd, capture_output=True, text=True): src for src, cmd in jobs}
            for fut in cf.as_completed(futs):
                res = fut.result()
                if ptxas_v:
                    sys.stderr.write(res.stderr)
                if res.returncode != 0:
                    raise RuntimeError(f"nvcc failed for {futs[fut]}:\n{res.stdout}\n{res.stderr}")
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, ptxas_v="--ptxas-v" in sys.argv))
