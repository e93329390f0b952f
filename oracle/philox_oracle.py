"""ORACLE — numpy restatement of the in-kernel counter RNG (csrc/philox.cuh); test infrastructure.

Philox4x32 (Salmon et al., SC'11; same constants as Random123 / cuRAND) with
counter = (col4 -- col8 for the packed reverse-noise stream --, row_lo, row_hi, stream << 16 | step), key = (seed_lo, seed_hi);
10 rounds for every stream except the reverse-step noise (stream 0), which uses 7 (the paper's smallest BigCrush-clean round count;
csrc/philox.cuh: philox_rounds).  The reference has no counterpart (it calls torch.randn, models/diffusion.py:335,409,443); the
10-round form is pinned by the Random123 known-answer vectors in tests/test_philox.py and the 7-round form is the same loop run
seven times (plus the algebraic check there that rounds compose).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


ROUNDS, ROUNDS_REVERSE = 10, 7


def rounds_of(stream: int) -> int:
    return ROUNDS_REVERSE if stream == 0 else ROUNDS


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    return philox4x32(ctr, key, ROUNDS)


def philox4x32(ctr: np.ndarray, key: np.ndarray, rounds: int = 10, first_round: int = 0) -> np.ndarray:
    """ctr [..., 4] uint32, key [..., 2] uint32 (broadcastable) -> [..., 4] uint32 after `rounds` Philox rounds; `first_round` offsets
    the key schedule (round r uses key + r * W), so philox(philox(c, k, a), k, b, first_round=a) == philox(c, k, a + b)."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    with np.errstate(over="ignore"):
        k0 = (np.asarray(key[..., 0], dtype=np.uint32) + np.uint32((int(W0) * first_round) & 0xFFFFFFFF)).astype(np.uint32)
        k1 = (np.asarray(key[..., 1], dtype=np.uint32) + np.uint32((int(W1) * first_round) & 0xFFFFFFFF)).astype(np.uint32)
        for _ in range(rounds):
            p0 = M0 * c[0]
            p1 = M1 * c[2]
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
            c = [hi1 ^ c[1] ^ k0.astype(np.uint64), lo1, hi0 ^ c[3] ^ k1.astype(np.uint64), lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def words(seed: int, rows: np.ndarray, ncol4: int, stream: int, step: int) -> np.ndarray:
    """uint32 [len(rows), ncol4, 4] exactly as osteo_philox_words writes them."""
    rows = np.asarray(rows, dtype=np.uint64)
    n = rows.shape[0]
    ctr = np.zeros((n, ncol4, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(ncol4, dtype=np.uint32)[None, :]
    ctr[..., 1] = (rows & MASK).astype(np.uint32)[:, None]
    ctr[..., 2] = (rows >> np.uint64(32)).astype(np.uint32)[:, None]
    ctr[..., 3] = np.uint32(((stream & 0xFFFF) << 16) | (step & 0xFFFF))
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32(ctr, key, rounds_of(stream))


def unit_1_2(w: np.ndarray) -> np.ndarray:
    """Mantissa trick of csrc/philox.cuh: float32 in [1, 2) from the top 23 bits."""
    return (np.uint32(0x3F800000) | (w >> np.uint32(9))).astype(np.uint32).view(np.float32)


def u01(w: np.ndarray) -> np.ndarray:
    """23-bit uniform in [0, 1) (dropout draws); exact in fp32."""
    return (unit_1_2(w) - np.float32(1.0)).astype(np.float32)


STREAM_REVERSE = 0      # csrc/philox.cuh: the reverse-step noise uses the PACKED mapping (one word per Box-Muller pair)


def normals(seed: int, rows: np.ndarray, d: int, stream: int, step: int) -> np.ndarray:
    """fp64-evaluated Box-Muller on the same words, in the stream's mapping (csrc/philox.cuh): float64 [len(rows), d]."""
    if stream == STREAM_REVERSE:
        return normals_packed(seed, rows, d, stream, step)
    ncol4 = (d + 3) // 4
    w = words(seed, rows, ncol4, stream, step)
    f = unit_1_2(w)
    u_r = (np.float32(2.0) - f).astype(np.float64)        # (0, 1]: radius uniforms (words 0 and 2)
    u_a = (f - np.float32(1.0)).astype(np.float64)        # [0, 1): angle uniforms (words 1 and 3)
    r0 = np.sqrt(-2.0 * np.log(u_r[..., 0]))
    r1 = np.sqrt(-2.0 * np.log(u_r[..., 2]))
    t0 = 2.0 * np.pi * u_a[..., 1]
    t1 = 2.0 * np.pi * u_a[..., 3]
    z = np.stack([r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=-1)
    return z.reshape(len(rows), ncol4 * 4)[:, :d]


def normals_packed(seed: int, rows: np.ndarray, d: int, stream: int, step: int) -> np.ndarray:
    """PACKED mapping: counter word 0 = column / 8, every 32-bit word gives one pair -- radius uniform u1 = 1 - (w >> 12) / 2^20 in
    (0, 1], angle 2 pi (w & 0xFFF) / 2^12 -- so a Philox block yields the 8 normals of columns 8b .. 8b + 7 (word j -> 2j, 2j + 1)."""
    ncol8 = (d + 7) // 8
    w = words(seed, rows, ncol8, stream, step)                                   # [n, ncol8, 4]
    f_r = (np.uint32(0x3F800000) | ((w >> np.uint32(12)) << np.uint32(3))).astype(np.uint32).view(np.float32)
    f_a = (np.uint32(0x3F800000) | ((w & np.uint32(0xFFF)) << np.uint32(11))).astype(np.uint32).view(np.float32)
    u_r = (np.float32(2.0) - f_r).astype(np.float64)
    ang = 2.0 * np.pi * (f_a - np.float32(1.0)).astype(np.float64)
    r = np.sqrt(-2.0 * np.log(u_r))
    z = np.stack([r * np.cos(ang), r * np.sin(ang)], axis=-1)                    # [n, ncol8, 4, 2]
    return z.reshape(len(rows), ncol8 * 8)[:, :d]
