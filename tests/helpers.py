"""Shared generators for the parity tests: coded and noisy code words, synthetic recordings."""
from __future__ import annotations

import numpy as np

POLY = (109, 79, 83, 109)


def conv_encode(bits: np.ndarray) -> np.ndarray:
    """Mother code of EN 300 401 11.1 in the register convention of ViterbiSpiral::calculate_BER (viterbi_spiral.cpp:128-164)."""
    n = len(bits)
    out = np.zeros(4 * (n + 6), np.uint8)
    sr = 0
    for i in range(n + 6):
        sr = ((sr << 1) | (int(bits[i]) if i < n else 0)) & 0xFF
        for j, p in enumerate(POLY):
            out[4 * i + j] = bin(sr & p).count("1") & 1
    return out


def soft_from_coded(coded: np.ndarray, sigma: float, rng: np.random.Generator, amp: float = 127.0) -> np.ndarray:
    """Reference polarity: bit 1 -> positive soft value."""
    v = (2.0 * coded.astype(np.float64) - 1.0) * amp + rng.normal(0.0, sigma, coded.shape)
    return np.clip(np.trunc(v), -32768, 32767).astype(np.int16)


def random_codewords(n: int, frame_bits: int, sigma: float, seed: int):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, (n, frame_bits), dtype=np.uint8)
    soft = np.stack([soft_from_coded(conv_encode(b), sigma, rng) for b in bits])
    return bits, soft


def puncture(coded_soft: np.ndarray, addresses: np.ndarray, fragment: int) -> np.ndarray:
    """Keeps the positions `addresses` (in order) of a full 4(N+6) soft vector, zero padded to `fragment`."""
    out = np.zeros(fragment, np.int16)
    out[:len(addresses)] = coded_soft[addresses]
    return out


# sub-channel profiles used across tests: (short_form, prot_level, bit_rate, size_cu)
EEP_A_72 = [(0, 0, 72, 108), (0, 1, 72, 72), (0, 2, 72, 54), (0, 3, 72, 36)]
EEP_B_64 = [(0, 4, 64, 54), (0, 5, 64, 42), (0, 6, 64, 36), (0, 7, 64, 30)]
UEP_128 = [(1, 1, 128, 140), (1, 2, 128, 116), (1, 3, 128, 96), (1, 4, 128, 84), (1, 5, 128, 64)]
ALL_PROFILES = EEP_A_72 + EEP_B_64 + UEP_128
