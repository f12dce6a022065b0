"""In-tree build of the CUDA library ``agimus_controller_b200/libagx.so`` (nvcc, sm_100a only)."""
from __future__ import annotations

import pathlib
import shutil
import subprocess

HERE = pathlib.Path(__file__).resolve().parent
SRC = HERE / "csrc"
LIB = HERE / "libagx.so"
SOURCES = ["agx_api.cu"]
HEADERS = ["agx_kernels.cuh", "agx_tree.cuh", "agx_riccati_mma.cuh", "agx_sqp.cuh", "agx_node.inl", "agx_dynamics.inl", "agx_octet_base.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "550",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not pathlib.Path(p).exists():
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return p


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [SRC / s for s in SOURCES + HEADERS] + [HERE.parent / "include" / "agx.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> pathlib.Path:
    """Compile the kernels + C ABI for sm_100a.  Cross-compiles without a GPU."""
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []),
           *[str(SRC / s) for s in SOURCES], "-o", str(LIB)]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
