"""In-tree build of libsemcom_sm100.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m multimodal_sc_b200.build [--force] [--verbose]

Objects go to multimodal_sc_b200/csrc/_obj/, the library to
multimodal_sc_b200/libsemcom_sm100.so (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = CSRC / "_obj"
LIB = PKG / "libsemcom_sm100.so"
INCLUDE = PKG.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    f"-I{INCLUDE}",
]


if os.environ.get("SC_MBAR_LIMIT", "") != "":                    # bring-up builds: mbarrier waits trap after 2^N tries
    NVCC_FLAGS.append(f"-DSC_MBAR_SPIN_LIMIT=(1u<<{int(os.environ['SC_MBAR_LIMIT'])})")
if os.environ.get("SC_MLP_ACCT", "0") not in ("", "0"):        # per-role wait accounting in the fused MLP kernel (debug builds)
    NVCC_FLAGS.append("-DSC_MLP_ACCT")
if os.environ.get("SC_GEMM_ACCT", "0") not in ("", "0"):       # per-role wait accounting in the GEMM kernel (debug builds)
    NVCC_FLAGS.append("-DSC_GEMM_ACCT")


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsemcom_sm100.so cannot be built")


def sources():
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    hs = list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return max(h.stat().st_mtime for h in hs)


def _compile(src: Path, verbose: bool) -> str:
    obj = OBJ / (src.stem + ".o")
    cmd = [nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
    (OBJ / (src.stem + ".ptxas.log")).write_text(r.stderr)
    if verbose:
        print(r.stderr)
    return str(obj)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sources()
    dep_t = _deps_mtime()
    todo = []
    for s in srcs:
        o = OBJ / (s.stem + ".o")
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, dep_t):
            todo.append(s)
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    objs = [str(OBJ / (s.stem + ".o")) for s in srcs]
    newest = max(os.path.getmtime(o) for o in objs)
    if force or todo or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs,
               "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
