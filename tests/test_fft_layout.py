"""The FFT kernel's epilogue layout (dabstar_b200/csrc/tables.cu, host_fft_epilogue_layout): a bipartite edge colouring gives every
nominal carrier a staging slot such that the scatter stores after the last FFT stage and the gather loads of the demapper's rows
are free of shared-memory bank conflicts. tests/native/fft_layout_check.cpp checks it on the host: slots unique, writer and reader
agree, every half warp of every store / load instruction on 16 different bank pairs."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc (host compile of tables.cu)")
def test_epilogue_layout_is_conflict_free(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    csrc = os.path.join(ROOT, "dabstar_b200", "csrc")
    exe = tmp_path / "fft_layout_check"
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-I", csrc, "-I", os.path.join(ROOT, "include"), "-o", str(exe), "-x", "cu",
                        os.path.join(ROOT, "tests", "native", "fft_layout_check.cpp"), os.path.join(csrc, "tables.cu")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "layout ok" in r.stdout, r.stdout + r.stderr
