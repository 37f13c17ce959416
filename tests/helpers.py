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


# ---- DAB+ super-frames (ETSI TS 102 563): test-signal generator for the outer-code stage
def _crc16_ccitt(data: bytes) -> int:
    crc = 0xFFFF
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return ~crc & 0xFFFF


def _firecode_parity(nine: bytes) -> int:
    """16 parity bits over header bytes 2..10, g(x) = x^16+x^14+x^13+x^12+x^11+x^5+x^3+x^2+x+1 (message * x^16 mod g)."""
    crc = 0
    for b in nine + b"\x00\x00":
        for i in range(8):
            top = (crc >> 15) & 1
            crc = ((crc << 1) & 0xFFFF) | ((b >> (7 - i)) & 1)
            if top:
                crc ^= 0x782F
    return crc


def dabplus_superframe(bit_rate: int, rng: np.random.Generator, rs_encode, dac_rate: int = 1, sbr_flag: int = 1):
    """One DAB+ audio super-frame: header + access units (each with its CRC) + RS parity, as 5 logical frames of 24*bit_rate bits.
    rs_encode: 110 data bytes -> 120 code-word bytes. Returns (bits[5, 24*bit_rate], list of AU payloads)."""
    rs_dims = bit_rate // 8
    size = 110 * rs_dims
    num_aus = {(0, 0): 4, (0, 1): 2, (1, 0): 6, (1, 1): 3}[(dac_rate, sbr_flag)]
    first = {4: 8, 2: 5, 6: 11, 3: 6}[num_aus]
    step = (size - first) / num_aus  # evenly spread with a little jitter: every access unit stays below the 960-byte limit
    cuts = [int(first + step * (i + 1) + rng.integers(-8, 9)) for i in range(num_aus - 1)]
    starts = [first] + [int(c) for c in cuts] + [size]
    sf = bytearray(size)
    sf[2] = (dac_rate << 6) | (sbr_flag << 5) | (1 << 4)
    nib = []
    for a in starts[1:num_aus]:
        nib += [(a >> 8) & 0xF, (a >> 4) & 0xF, a & 0xF]
    if len(nib) % 2:
        nib.append(0)
    for i in range(0, len(nib), 2):
        sf[3 + i // 2] = (nib[i] << 4) | nib[i + 1]
    aus = []
    for u in range(num_aus):
        n = starts[u + 1] - starts[u] - 2
        au = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        crc = _crc16_ccitt(au)
        sf[starts[u]:starts[u] + n] = au
        sf[starts[u] + n] = crc >> 8
        sf[starts[u] + n + 1] = crc & 0xFF
        aus.append(au)
    par = _firecode_parity(bytes(sf[2:11]))
    sf[0], sf[1] = par >> 8, par & 0xFF
    block = np.zeros(120 * rs_dims, np.uint8)
    data = np.frombuffer(bytes(sf), np.uint8)
    for j in range(rs_dims):
        block[j::rs_dims] = rs_encode(data[j::rs_dims])
    return np.unpackbits(block).reshape(5, 24 * bit_rate), aus


# ---- TII (EN 300 401 clause 14.8): null-symbol spectra for the detector tests
TII_PATTERNS = [b for b in range(256) if bin(b).count("1") == 4]   # main id -> group pattern (MSB = group 0)


def tii_pair_fft_index(i: int) -> int:
    """fft index of the first carrier of pair i (carriers k = -768 + 2 i and k + 1, DC skipped)."""
    k = -768 + 2 * i
    return k + 2048 if k < 0 else k + 1


def tii_spectrum(ids, rng, amp=30.0, noise=1.0, non_etsi=False, prs=None, single_carriers=()):
    """One null-symbol spectrum (complex64[2048], fft order) carrying the transmitters `ids` = [(main, sub, phase_rad), ...]:
    every active carrier pair has both carriers at the same phase (ETSI) or at the PRS phases (the 'non-ETSI' transmitters the
    reference also looks for); plus white noise and optional lone carriers (which the detector must not take for TII)."""
    x = (rng.normal(size=2048) + 1j * rng.normal(size=2048)) * (noise / np.sqrt(2.0))
    for main, sub, ph in ids:
        pat = TII_PATTERNS[main]
        for blk in range(4):
            for grp in range(8):
                if not pat & (0x80 >> grp):
                    continue
                f = tii_pair_fft_index(blk * 192 + grp * 24 + sub)
                a = amp * np.exp(1j * ph)
                if non_etsi:
                    x[f] += a * prs[f]
                    x[f + 1] += a * prs[f + 1]
                else:
                    x[f] += a
                    x[f + 1] += a
    for i, a in single_carriers:
        x[tii_pair_fft_index(i)] += a
    return x.astype(np.complex64)
