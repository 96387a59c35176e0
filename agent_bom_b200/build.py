"""Build libabb200.so (the C-ABI engine) in-tree with nvcc for sm_100a.

    python -m agent_bom_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libabb200.so"
SOURCES = [CSRC / "abb200.cu"]
DEPS = sorted(p for p in CSRC.iterdir() if p.suffix in (".cu", ".cuh", ".inl", ".h")) + sorted((PKG.parent / "include").glob("*.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the engine has no CPU fallback and cannot be built without the CUDA toolkit")


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(d.stat().st_mtime > t for d in DEPS if d.exists())


def build(force: bool = False, verbose: bool = False, variant: str = "", defines: tuple = ()) -> Path:
    """Default: libabb200.so.  `variant` + `defines` build an experiment library next to it (libabb200_<variant>.so with the given
    -D flags; loaded through the ABB_LIB environment variable) — used for in-run A/B measurements, never by the product path."""
    out = LIB if not variant else PKG / f"libabb200_{variant}.so"
    if not variant and not force and not is_stale():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], *(["-Xptxas", "-v"] if verbose else []), "-o", str(out), *map(str, SOURCES)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode:
        raise RuntimeError(f"nvcc failed ({proc.returncode}): {' '.join(cmd)}")
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=args[0] if args else "", defines=tuple(args[1:])))
