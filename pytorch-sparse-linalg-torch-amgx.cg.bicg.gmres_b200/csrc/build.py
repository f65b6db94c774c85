#!/usr/bin/env python3
"""Build libbk_krylov.so (sm_100a only) in-tree with plain nvcc: one object per .cu in parallel, then link.

    python build.py [--force] [--verbose]

No torch headers are involved (the library is a plain C ABI), so this takes seconds per file and
cross-compiles on a machine without a GPU.
"""
import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE.parent / "pytorch_sparse_solver" / "_lib"
LIB = OUT / "libbk_krylov.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v" if "--verbose" in sys.argv else "-O3",
]


def newest(paths):
    return max(p.stat().st_mtime for p in paths)


def main():
    force = "--force" in sys.argv
    srcs = sorted(HERE.glob("*.cu"))
    hdrs = sorted(HERE.glob("*.cuh")) + [HERE.parent.parent / "include" / "bk_krylov.h"]
    OUT.mkdir(parents=True, exist_ok=True)
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    dep_time = newest(hdrs + [Path(__file__)])

    def compile_one(src):
        obj = objdir / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, dep_time):
            return obj, ""
        cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [str(o) for o, _ in results]
    if "--verbose" in sys.argv:
        for _, log in results:
            sys.stderr.write(log)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest([Path(o) for o in objs]):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *objs,
               "-Xcompiler", "-fPIC", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    print(str(LIB))


if __name__ == "__main__":
    main()
