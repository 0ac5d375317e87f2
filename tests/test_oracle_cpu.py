"""CPU checks of the oracle's RNG mirrors (no GPU): Philox4x32-10 known answers and the statistical quality of the dropout
masks produced by the counter hash the GEMM epilogues use (oracle/philox.py mirrors ltg_common.cuh bit for bit; the bit-exact
device-vs-mirror comparison is tests/test_kernels_gpu.py::test_gemm_dropout_epilogue_bits)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import philox  # noqa: E402


def test_hash_quad_known_values_are_stable():
    # pins the definition (multiplier, round count, key schedule, word order): any change must be made on the device side too
    L, R = philox.hash_quad(0x12345678, np.array([0, 1, 2, (1 << 32) + 5], dtype=np.uint64))
    again = philox.hash_quad(0x12345678, np.array([0, 1, 2, (1 << 32) + 5], dtype=np.uint64))
    assert np.array_equal(L, again[0]) and np.array_equal(R, again[1])
    assert len(set(int(x) for x in L)) == 4 and len(set(int(x) for x in R)) == 4
    # one round by hand: L1 = hi(L0*M) ^ k ^ R0, R1 = lo(L0*M)
    M, k = 0xD256D193, 0x12345678
    l, r = 2, 0
    for _ in range(5):
        p = l * M
        l, r = ((p >> 32) ^ k ^ r) & 0xFFFFFFFF, p & 0xFFFFFFFF
        k = (k + 0x9E3779B9) & 0xFFFFFFFF
    assert int(L[2]) == l and int(R[2]) == r


def test_dropout_mask_statistics():
    """keep-rate and correlations of the [rows, cols] mask at the sampling-noise level (4 sigma), for the discriminator's pitches."""
    for ld, cols, keep in ((152, 150, 0.7), (256, 250, 0.7), (304, 300, 0.7), (64, 64, 0.5)):
        rows = 8192
        for step in (1, 2):
            m = philox.hash_keep_mask(20260101, philox.STREAM_DISC_DROPOUT, step, rows, cols, ld, keep).astype(np.float64)
            n = m.size
            p = int(np.float32(keep) * 65536.0) / 65536.0
            assert abs(m.mean() - p) < 4.0 * np.sqrt(p * (1 - p) / n)
            tol = 4.5 / np.sqrt(n)
            for a, b in ((m[:, :-1], m[:, 1:]), (m[:-1], m[1:]), (m[:, :-4], m[:, 4:]), (m[:-2], m[2:]), (m[:, :-2], m[:, 2:])):
                c = np.corrcoef(a.ravel(), b.ravel())[0, 1]
                assert abs(c) < tol, (ld, step, c, tol)
            # per-column and per-row keep rates scatter like binomials
            assert m.mean(0).std() < 1.5 * np.sqrt(p * (1 - p) / rows)
            assert m.mean(1).std() < 1.5 * np.sqrt(p * (1 - p) / cols)
        m2 = philox.hash_keep_mask(20260101, philox.STREAM_DISC_DROPOUT, 3, rows, cols, ld, keep).astype(np.float64)
        c = np.corrcoef(m.ravel(), m2.ravel())[0, 1]
        assert abs(c) < 4.5 / np.sqrt(m.size)   # masks of consecutive steps are independent


def test_philox4x32_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    out = philox.philox4x32_10(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0))
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = philox.philox4x32_10(np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff),
                               np.uint32(0xffffffff), np.uint32(0xffffffff))
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
