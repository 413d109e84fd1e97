"""Build helpers: compile the CUDA library (sm_100a) and the CPU oracle in-tree.

The product is ``libarrowspace_b200.so`` (hand-written CUDA behind the C ABI declared in
``include/arrowspace_b200.h``).  The oracle (``oracle/libarrowspace_oracle.so``) is test
infrastructure only; building it here is not using it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libarrowspace_b200.so"
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "libarrowspace_oracle.so"

CUDA_SOURCES = ["api.cu", "taumode.cu", "search.cu", "laplacian.cu", "cluster.cu", "cluster_replay.cu", "extras.cu", "comm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (no CPU fallback exists)")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in CUDA_SOURCES]
    deps = srcs + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "arrowspace_b200.h"]
    if force or _stale(LIB_PATH, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, srcs)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.run(cmd, check=True, cwd=str(CSRC), env=env)
    return LIB_PATH


def build_oracle(force: bool = False) -> Path:
    src = ORACLE_DIR / "arrowspace_oracle.c"
    if force or _stale(ORACLE_LIB, [src, ORACLE_DIR / "arrowspace_oracle.h"]):
        gcc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else (shutil.which("gcc") or "gcc")
        base = [gcc, "-O2", "-ffp-contract=off", "-fPIC", "-std=c11", "-shared", "-o", str(ORACLE_LIB), str(src), "-lm"]
        try:
            subprocess.run(base[:1] + ["-fopenmp"] + base[1:], check=True, cwd=str(ORACLE_DIR),
                           stderr=subprocess.DEVNULL)
        except subprocess.CalledProcessError:
            subprocess.run(base, check=True, cwd=str(ORACLE_DIR))  # no libgomp: single-threaded oracle
    return ORACLE_LIB


CPP_EXAMPLE = ROOT / "tools" / "cpp_example"


def build_cpp_example(force: bool = False) -> Path:
    """The C++ host mirror (include/arrowspace_b200.hpp) compiled against the C ABI."""
    src = ROOT / "tools" / "cpp_example.cpp"
    if force or _stale(CPP_EXAMPLE, [src, ROOT / "include" / "arrowspace_b200.hpp", ROOT / "include" / "arrowspace_b200.h"]):
        gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else (shutil.which("g++") or "g++")
        subprocess.run([gxx, "-std=c++17", "-O2", f"-I{ROOT / 'include'}", str(src), f"-L{PKG_DIR}", "-larrowspace_b200",
                        f"-Wl,-rpath,{PKG_DIR}", "-o", str(CPP_EXAMPLE)], check=True)
    return CPP_EXAMPLE


def build_all(force: bool = False) -> None:
    build_cuda(force=force)
    build_oracle(force=force)
    build_cpp_example(force=force)


if __name__ == "__main__":
    build_all(force=True)
    print(LIB_PATH, ORACLE_LIB)
