"""TEST INFRASTRUCTURE — not part of the product path.

CPU twin of the counter-based noise the CUDA kernels generate in-register.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this module.

The reference draws its MC-dropout masks and reparameterisation normals from
torch's global CPU generator in call order (src/torchmodel.py:44-50,55,96-118,
131; draw order measured in SURVEY.md §8 a5).  A GPU kernel cannot replay that
stream, so parity is defined on a *keyed* noise function instead:

    noise(key, step, sample, site, row, element)

with Philox4x32-10 (Salmon et al., SC'11; the published Random123 algorithm,
pinned below by its known-answer vectors in tests/test_philox.py):

    key      = seed + call_index                (64 bit, one API call = one key)
    counter  = (block | site << 16, row, sample, step)
    mask bit = bit (e & 31) of word (e >> 5) & 3 of block (e >> 7)   -> keep*2
    normal   = sqrt(-2 ln u1) * cos(2 pi u2), u = (word + 0.5) * 2^-32 in
               float64 from words 0,1 of block e, rounded once to float32
    uniform  = (word0 >> 8) * 2^-24 as float32 (categorical draws)

Sites (one per reference RNG draw inside one MC sample, same order as a5):
see SITES below.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

# site ids: <net>_<slot> base + draw index inside the net (3 dropout + 1 tail)
SITES = {
    "PS_A": 0,    # loop 2a transition: d512,d512,d512,r10   (torchmodel.py:274)
    "PO_A": 4,    # loop 2a decoder:    d256,d256,d256,d16384 (:275)
    "QS_A": 8,    # loop 2a encoder:    d256,d256,d256,r10   (:276)
    "PS_B": 12,   # loop 2b transition                          (:288)
    "PO_B1": 16,  # loop 2b decoder of the fresh transition     (:288)
    "RP_B": 20,   # loop 2b reparameterize(ps1_mean, ps1_logvar) (:291)
    "PO_B2": 21,  # loop 2b decoder of the reparameterised s    (:291)
    "QS_ROOT": 32,  # root encoder + reparam (:228-229, :248-249)
    "CAT": 40,    # categorical action draw in mcts_step_simulate (:364,379)
}


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        np.asarray(c0, dtype=np.uint32), np.asarray(c1, dtype=np.uint32),
        np.asarray(c2, dtype=np.uint32), np.asarray(c3, dtype=np.uint32))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = c0.astype(np.uint64) * M0
        p1 = c2.astype(np.uint64) * M1
        hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
        lo0 = (p0 & MASK32).astype(np.uint32)
        hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
        lo1 = (p1 & MASK32).astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint32(k0), lo1, hi0 ^ c3 ^ np.uint32(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def split_key(key64):
    key64 = int(key64) & 0xFFFFFFFFFFFFFFFF
    return key64 & 0xFFFFFFFF, key64 >> 32


def dropout_mask(key64, step, sample, site, rows, n):
    """float32 (len(rows), n) of {0, 2}: the inverted-dropout multiplier at p=0.5."""
    rows = np.asarray(rows, dtype=np.uint32)
    nblk = (n + 127) // 128
    blk = np.arange(nblk, dtype=np.uint32)
    k0, k1 = split_key(key64)
    w = philox4x32_10(blk[None, :] | np.uint32(site << 16), rows[:, None],
                      np.uint32(sample), np.uint32(step), k0, k1)
    words = np.stack(w, axis=-1).reshape(len(rows), nblk * 4)       # word index = blk*4 + j
    e = np.arange(n)
    bits = (words[:, e >> 5] >> (e & 31).astype(np.uint32)) & np.uint32(1)
    return (bits.astype(np.float32) * np.float32(2.0))


def normals(key64, step, sample, site, rows, n):
    """float32 (len(rows), n) standard normals (Box-Muller in float64, one rounding)."""
    rows = np.asarray(rows, dtype=np.uint32)
    e = np.arange(n, dtype=np.uint32)
    k0, k1 = split_key(key64)
    w0, w1, _, _ = philox4x32_10(e[None, :] | np.uint32(site << 16), rows[:, None],
                                 np.uint32(sample), np.uint32(step), k0, k1)
    u1 = (w0.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    u2 = (w1.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586 * u2)
    return z.astype(np.float32)


def uniform24(key64, step, sample, site, row):
    """One float32 uniform in [0,1) with 24 random bits."""
    k0, k1 = split_key(key64)
    w0, _, _, _ = philox4x32_10(np.uint32(site << 16), np.uint32(row), np.uint32(sample),
                                np.uint32(step), k0, k1)
    return np.float32(int(w0) >> 8) * np.float32(1.0 / 16777216.0)
