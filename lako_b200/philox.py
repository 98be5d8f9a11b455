"""Host restatement of the kernels' counter-based dropout mask (csrc/common.cuh: philox4x32 (7 rounds) / drop_keep8),
used by the tests to rebuild the exact keep-masks a kernel applied and by users who need to audit them."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


ROUNDS = 7


def philox4x32(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(ROUNDS):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def thresh_of(p):
    return int(p * 65536.0 + 0.5) if p > 0 else 0


def scale_of(p):
    return 1.0 / (1.0 - thresh_of(p) / 65536.0) if p > 0 else 1.0


def keep_mask(seed, step, site, n_elems, p):
    """bool numpy array [n_elems]: keep decision of element e (group e>>3, 16-bit lane e&7)."""
    if p <= 0:
        return np.ones(n_elems, dtype=bool)
    groups = (n_elems + 7) // 8
    g = np.arange(groups, dtype=np.uint64)
    r = philox4x32((g & MASK32).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                      np.full(groups, site, dtype=np.uint32), np.full(groups, step, dtype=np.uint32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lanes = np.empty((groups, 8), dtype=np.uint32)
    for j in range(4):
        lanes[:, 2 * j] = r[j] & np.uint32(0xFFFF)
        lanes[:, 2 * j + 1] = r[j] >> np.uint32(16)
    return (lanes >= thresh_of(p)).reshape(-1)[:n_elems]


def attn_keep_words(seed, step, site, group4, p):
    """uint32 numpy array like group4: the keep word of one (query row, 32-key chunk) of the attention-probability dropout
    (csrc/common.cuh drop_keep32_sliced).  group4 = index of the first of its 4 consecutive Philox counters
    (((passage * H + head) * L + query) * 32 + chunk * 4).  The 16 output words are bit planes of 32 independent 16-bit
    uniforms, compared with the threshold from the least significant plane up."""
    group4 = np.asarray(group4, dtype=np.uint64)
    ge = np.full(group4.shape, 0xFFFFFFFF, dtype=np.uint32)
    if p <= 0:
        return ge
    t = thresh_of(p)
    for u in range(4):
        g = group4 + np.uint64(u)
        r = philox4x32((g & MASK32).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                          np.full(g.shape, site, dtype=np.uint32), np.full(g.shape, step, dtype=np.uint32),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for j in range(4):
            ge = (r[j] & ge) if (t >> (u * 4 + j)) & 1 else (r[j] | ge)
    return ge


def attn_keep_mask(seed, step, site, P, H, L, p):
    """bool numpy array [P, H, L, 32 * nch]: keep decision of every (passage, head, query, key), keys beyond L included
    (key j of chunk c is bit (j >> 1) + 16 (j & 1) of that chunk's word).  nch = max(8, ceil(L / 32)) chunks per query row
    -- 8 for the single-tile kernels (L <= 256), more for the tiled path (csrc/enc_attn_long.cu enc_attn_keep_stride)."""
    nch = max(8, (L + 31) // 32)
    rows = np.arange(P * H * L, dtype=np.uint64).reshape(P, H, L, 1)
    chunks = np.arange(nch, dtype=np.uint64).reshape(1, 1, 1, nch)
    words = attn_keep_words(seed, step, site, rows * np.uint64(4 * nch) + chunks * np.uint64(4), p)  # [P, H, L, nch]
    j = np.arange(32)
    bit = (j >> 1) + 16 * (j & 1)
    return ((words[..., None] >> bit.astype(np.uint32)) & np.uint32(1)).astype(bool).reshape(P, H, L, 32 * nch)
