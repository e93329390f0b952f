"""CPU: the numpy Philox4x32-10 restatement against the Random123 known-answer vectors (kat_vectors, philox4x32 10 rounds)."""
import numpy as np

from oracle import philox_oracle as P

KAT = [
    ((0x00000000,) * 4, (0x00000000,) * 2, (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_random123_known_answers():
    for ctr, key, out in KAT:
        got = P.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(v) for v in got) == out


def test_counter_layout_and_uniforms():
    w = P.words(seed=(7 << 32) | 5, rows=np.array([3, (1 << 32) + 3], dtype=np.uint64), ncol4=2, stream=2, step=999)
    assert w.shape == (2, 2, 4) and w.dtype == np.uint32
    direct = P.philox4x32_10(np.array([1, 3, 1, (2 << 16) | 999], dtype=np.uint32), np.array([5, 7], dtype=np.uint32))
    assert np.array_equal(w[1, 1], direct)
    u = P.u01(np.array([0, 0xFFFFFFFF], dtype=np.uint32))
    assert u[0] == 0.0 and 0.999 < u[1] < 1.0      # [0, 1); the radius uniform is 2 - f in (0, 1], never 0


def test_normals_have_unit_moments():
    z = P.normals(1, np.arange(512, dtype=np.uint64), 1024, 0, 10)
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1.0) < 1e-2


def test_packed_mapping_of_the_reverse_noise_stream():
    """oracle.normals: stream 0 (reverse-step noise) maps ONE word to a Box-Muller pair (20-bit radius uniform, 12-bit angle), so a block
    covers 8 columns; the other streams keep two words per pair. Checks the bit layout on hand-computed values and the moments."""
    seed, rows, d, step = 9, np.arange(3, dtype=np.uint64), 20, 5
    z = P.normals(seed, rows, d, 0, step)
    w = P.words(seed, rows, 3, 0, step)                   # ceil(20 / 8) blocks
    for (r, b, j) in [(0, 0, 0), (1, 1, 3), (2, 2, 1)]:
        word = int(w[r, b, j])
        u1 = 1.0 - (word >> 12) / 2.0 ** 20
        ang = 2.0 * np.pi * (word & 0xFFF) / 4096.0
        rad = np.sqrt(-2.0 * np.log(u1))
        col = 8 * b + 2 * j
        assert abs(z[r, col] - rad * np.cos(ang)) < 1e-12 and abs(z[r, col + 1] - rad * np.sin(ang)) < 1e-12
    assert z.shape == (3, 20)
    other = P.normals(seed, rows, d, 1, step)             # two-word mapping: different values, same shape
    assert other.shape == z.shape and not np.allclose(other, z)
    big = P.normals(3, np.arange(256, dtype=np.uint64), 4096, 0, 1)
    assert abs(big.mean()) < 4e-3 and abs(big.var() - 1.0) < 6e-3 and abs((big ** 4).mean() - 3.0) < 0.05
    assert np.abs(big).max() <= np.sqrt(2 * 20 * np.log(2)) + 1e-12


def test_seven_round_stream_is_a_prefix_of_the_pinned_ten_round_function():
    """The reverse-noise stream uses Philox4x32-7 (csrc/philox.cuh: philox_rounds). Random123's known answers pin the 10-round function;
    rounds compose -- ten rounds == seven rounds followed by rounds 7..9 of the same key schedule -- so the 7-round words are pinned by
    the same vectors, and they differ from the 10-round words."""
    for ctr, key, out in KAT:
        c, k = np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32)
        seven = P.philox4x32(c, k, 7)
        assert tuple(int(v) for v in P.philox4x32(seven, k, 3, first_round=7)) == out
        assert tuple(int(v) for v in seven) != out
    assert P.rounds_of(0) == 7 and all(P.rounds_of(s) == 10 for s in (1, 2, 3, 16))
    w7 = P.words(5, np.arange(4, dtype=np.uint64), 3, 0, 17)
    w10 = P.words(5, np.arange(4, dtype=np.uint64), 3, 1, 17)
    assert not np.array_equal(w7, w10)
    # avalanche sanity of the 7-round function: flipping one counter bit flips about half of the 128 output bits
    base = np.array([[i, 7, 0, 99] for i in range(256)], dtype=np.uint32)
    flipped = base.copy()
    flipped[:, 0] ^= np.uint32(1)
    k = np.array([123, 456], dtype=np.uint32)
    d = P.philox4x32(base, k, 7) ^ P.philox4x32(flipped, k, 7)
    bits = np.unpackbits(d.view(np.uint8)).reshape(256, 128).sum(1)
    assert 56 < bits.mean() < 72
