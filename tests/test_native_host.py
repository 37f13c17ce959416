"""CPU checks of the product's host-side logic and of the kernels' index arithmetic, run through small nvcc-built
host programs (no GPU): FFT decomposition, warp Viterbi algorithm with emulated lanes, puncturing tables."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dabstar_b200", "csrc")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def fnv(data: bytes) -> int:
    h = 2166136261
    for b in data:
        h = ((h ^ b) * 16777619) & 0xFFFFFFFF
    return h


@pytest.fixture(scope="module")
def host_logic(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("native") / "host_logic")
    subprocess.run([NVCC, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", CSRC, "-o", exe,
                    os.path.join(ROOT, "tests", "native", "host_logic.cu"), os.path.join(CSRC, "tables.cu")], check=True)
    return exe


def test_fft_decomposition_emulated(tmp_path):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fft_emu")
    subprocess.run([NVCC, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", CSRC, "-o", exe, os.path.join(ROOT, "tests", "native", "fft_emulation.cu")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout


def test_puncturing_tables_match_oracle(host_logic, oracle):
    out = subprocess.run([host_logic, "addr"], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        _, sf, lvl, br, n, h = line.split()
        sf, lvl, br, n, h = int(sf), int(lvl), int(br), int(n), int(h)
        if sf == 9:
            a = oracle.fic_addresses()
            assert n == 2304
        elif n < 0:
            # profiles the reference's table does not define (it would silently fall back, uep_protection.cpp:163-167)
            assert sf == 1 and (br, lvl) in {(56, 1), (112, 1), (320, 1), (320, 3), (384, 2), (384, 4)}
            continue
        else:
            a = oracle.protection_addresses(sf, br, lvl)
            assert len(a) == n, line
        assert fnv(np.ascontiguousarray(a[:n], np.int32).tobytes()) == h, line
        seen += 1
    assert seen >= 64 + 48


@pytest.mark.parametrize("mode", ["vit", "vit2"])  # warp per code word (viterbi.cuh), thread per code word (viterbi_tpc.cuh)
@pytest.mark.parametrize("frame_bits,sigma", [(768, 0.0), (768, 150.0), (768, 400.0), (192, 200.0), (1728, 170.0), (9216, 300.0)])
def test_warp_viterbi_algorithm_emulated(host_logic, oracle, frame_bits, sigma, mode):
    n = 6
    _, soft = helpers.random_codewords(n, frame_bits, sigma, seed=frame_bits + int(sigma))
    soft[-1] = 0  # all erasures: every comparison ties, path 0 must win
    out = subprocess.run([host_logic, mode, str(frame_bits), str(n)], input=soft.tobytes(), capture_output=True, check=True).stdout.decode()
    lines = out.splitlines()
    assert len(lines) == n
    for i, line in enumerate(lines):
        assert int(line.split()[3]) == fnv(oracle.viterbi(soft[i], frame_bits).tobytes()), i
