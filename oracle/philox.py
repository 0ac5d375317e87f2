"""TEST INFRASTRUCTURE -- numpy mirror of the stateless Philox4x32-10 streams used by the CUDA kernels
(long-tail-gan_b200/csrc/ltg_common.cuh). The reference draws its randomness from TensorFlow's and NumPy's
global generators (MultiVAE.py:149,178; discriminator.py:25-55; sample.py:54; train.py:236), whose streams are
not reproducible outside those libraries; parity tests therefore *inject* identical randomness on both sides,
and this module produces bit-identical copies of what the device generates.
"""
import numpy as np

STREAM_ENC_DROPOUT = 1
STREAM_EPS = 2
STREAM_DISC_DROPOUT = 3
STREAM_SAMPLE = 8
STREAM_PARTNER = 9

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds. All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def rand_u32(seed, stream, step, idx):
    """ltg_rand_u32: element `idx` takes word idx&3 of the Philox block with counter (idx>>2, stream, step)."""
    idx = np.asarray(idx, dtype=np.uint64)
    blk = idx >> np.uint64(2)
    r = philox4x32_10((blk & _MASK32).astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32), np.uint32(stream), np.uint32(step),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = (idx & np.uint64(3)).astype(np.int64)
    out = np.where(w == 0, r[0], np.where(w == 1, r[1], np.where(w == 2, r[2], r[3])))
    return out.astype(np.uint32)


def keep_threshold(keep):
    t = float(np.float32(keep)) * 4294967296.0
    if t >= 4294967295.0:
        return 0xFFFFFFFF
    if t <= 0.0:
        return 0
    return int(t)


def keep_mask(seed, stream, step, idx, keep):
    """Boolean dropout keep-mask (True = kept), bit-identical to the device."""
    if not (0.0 < keep < 1.0):
        return np.ones(np.shape(idx), dtype=bool)
    return rand_u32(seed, stream, step, idx) < np.uint32(keep_threshold(keep))


def u01(r):
    return ((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def normal_eps(seed, step, idx):
    """Box-Muller normal of latent_fwd_kernel: one Philox block per element (counter = idx), words x,y."""
    idx = np.asarray(idx, dtype=np.uint64)
    r = philox4x32_10((idx & _MASK32).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32), np.uint32(STREAM_EPS), np.uint32(step),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1, u2 = u01(r[0]), u01(r[1])
    return (np.sqrt(np.float32(-2.0) * np.log(u1)) * np.cos(np.float32(2.0 * np.pi) * u2)).astype(np.float32)


def gumbel(seed, step, idx):
    r = rand_u32(seed, STREAM_SAMPLE, step, idx)
    return (-np.log(-np.log(u01(r)))).astype(np.float32)


# ---- cheap counter hash used by the GEMM-epilogue dropout (ltg_common.cuh: ltg_lowbias32 / ltg_hash_key / ltg_hash_pair) ----
def lowbias32(x):
    x = np.asarray(x, dtype=np.uint64) & _MASK32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7feb352d)) & _MASK32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846ca68b)) & _MASK32
    x ^= x >> np.uint64(16)
    return x


def hash_key(seed, stream, step):
    inner = lowbias32((np.uint64(stream) * np.uint64(0x9E3779B9) + np.uint64(step)) & _MASK32)
    mid = lowbias32(np.uint64((seed >> 32) & 0xFFFFFFFF) ^ inner)
    return lowbias32(np.uint64(seed & 0xFFFFFFFF) ^ mid)


def hash_quad(key, g):
    """ltg_hash_quad (ltg_common.cuh): Philox2x32, 5 rounds, multiplier 0xD256D193, round keys key + r * 0x9E3779B9, counter
    (lo32(g), hi32(g)). Returns the two 32-bit output words as uint64 arrays."""
    g = np.asarray(g, dtype=np.uint64)
    L = g & _MASK32
    R = (g >> np.uint64(32)) & _MASK32
    k = np.uint64(int(key) & 0xFFFFFFFF)
    for _ in range(5):
        p = L * np.uint64(0xD256D193)
        L = ((p >> np.uint64(32)) ^ k ^ R) & _MASK32
        R = p & _MASK32
        k = (k + np.uint64(0x9E3779B9)) & _MASK32
    return L, R


def hash_keep_mask(seed, stream, step, n_rows, n_cols, rng_ld, keep):
    """Keep-mask [n_rows, n_cols] of the GEMM dropout epilogues: group g = (row*rng_ld + col) // 4 -> 64 bits; column 4g+0 / +1
    take the low / high 16 bits of the first word, +2 / +3 those of the second; kept iff bits < floor(keep * 65536)."""
    if not (0.0 < keep < 1.0):
        return np.ones((n_rows, n_cols), dtype=bool)
    key = hash_key(seed, stream, step)
    idx = np.arange(n_rows, dtype=np.uint64)[:, None] * np.uint64(rng_ld) + np.arange(n_cols, dtype=np.uint64)[None, :]
    L, R = hash_quad(key, idx >> np.uint64(2))
    f = idx & np.uint64(3)
    w = np.where(f < np.uint64(2), L, R)
    bits = np.where((f & np.uint64(1)) == 0, w & np.uint64(0xFFFF), w >> np.uint64(16))
    thr = int(float(np.float32(keep)) * 65536.0)
    return bits < np.uint64(thr)
