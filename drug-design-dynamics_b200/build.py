"""In-tree build of libddd_mc.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
CSRC = HERE / "csrc"
LIB = CSRC / "libddd_mc.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",              # fp64 parity paths must not be contracted; fused ops are explicit fma()
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libddd_mc.so cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    """Everything ddd_mc.cu includes: every .cu/.cuh under csrc/ and every header under include/."""
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources())


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(LIB), str(CSRC / "ddd_mc.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


HOST = HERE / "host"


def build_host() -> list[Path]:
    """C++ mirror of the reference's Go API (host/): the drop-in binary and its metropolis_test.go restatement."""
    build_library()
    outs = []
    common = ["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", str(HOST / "metropolis.cpp")]
    link = [f"-L{CSRC}", "-lddd_mc", f"-Wl,-rpath,{CSRC}", "-Wl,-rpath,$ORIGIN/../csrc", "-lpthread"]
    for name, src in (("ddd_metropolis", "main.cpp"), ("ddd_metropolis_test", "metropolis_test.cpp")):
        out = HOST / name
        res = subprocess.run(common + [str(HOST / src), "-o", str(out)] + link, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("host build failed:\n" + res.stdout + res.stderr)
        outs.append(out)
    return outs


def build_oracle() -> Path:
    """Compile the CPU oracle (test infrastructure; building the checker is not using it)."""
    res = subprocess.run(["make", "-C", str(ROOT / "oracle")], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return ROOT / "oracle" / "libddd_oracle.so"
