"""The shipped kernels under AddressSanitizer (CPU SIMT emulator build): every "global memory" access of a ragged
small case (B = 5, T = 7: partial warps and CTAs everywhere) stays inside its allocation.  compute-sanitizer is closed
on the GPU pool, so this is the memory-safety evidence for the kernels' indexing."""
import os
import pathlib
import subprocess
import sys

import pytest

HERE = pathlib.Path(__file__).resolve().parent
SRC = HERE.parent / "agimus_controller_b200" / "csrc" / "agx_api.cu"


@pytest.mark.timeout(600)
def test_emulated_kernels_under_asan(tmp_path):
    so = tmp_path / "libagx_emul_asan.so"
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer",
           "-ffp-contract=off", "-fPIC", "-DAGX_EMULATE", "-include", str(HERE / "emul" / "cpu_simt.h"), "-shared",
           "-o", str(so), str(SRC)]
    subprocess.run(cmd, check=True, cwd=str(HERE / "emul"))
    libasan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
    if not os.path.exists(libasan):
        pytest.skip("libasan not available")
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, str(HERE / "emul" / "asan_case.py"), str(so)], capture_output=True, text=True,
                       env=env, timeout=500)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "asan case ok" in r.stdout
    assert "ERROR: AddressSanitizer" not in r.stderr
