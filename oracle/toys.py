"""Oracle for on-device toy generation (test infrastructure; only tests/ may import this).

The reference draws toys with NumPy's global Mersenne Twister (model.py:69-91 -> source.py:248-264 -> multihist
Histdd.get_random); the device generator uses counter-based Philox4x32-10 instead, so parity with the reference is
distributional.  What IS restated exactly here, given the random stream:
  * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; pinned by the Random123
    known-answer vectors in tests/test_oracle_pins.py) and NumPy's 53-bit uniform construction,
  * Histdd.get_random's event stage: bin = min(searchsorted(cdf / cdf[-1], u), n - 1) on the flattened pmf,
    unravel, position lo + u * (hi - lo) per dimension,
  * Model.simulate's layout: the events of a toy grouped by source, in source order (model.py:88).
The Poisson stage (NumPy-legacy multiplication / PTRS algorithms) is checked statistically only: device exp/log may
round differently from libm and flip a rare accept/reject."""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: uint32 [n, 4]; key: (k0, k1).  Returns uint32 [n, 4]."""
    c = [np.asarray(counter[:, i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & _MASK,
             (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & _MASK]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


def uniform53(a, b):
    """NumPy's random_double: ((a >> 5) * 2^26 + (b >> 6)) / 2^53."""
    return ((a.astype(np.uint64) >> np.uint64(5)) * np.uint64(1 << 26) + (b.astype(np.uint64) >> np.uint64(6))) \
        .astype(np.float64) * (1.0 / 9007199254740992.0)


def toy_uniforms(seed, toy, index, domain):
    """Two uniforms per (toy, index) in `domain`; toy / index are integer arrays."""
    toy = np.asarray(toy, dtype=np.uint64)
    n = len(toy)
    ctr = np.empty((n, 4), dtype=np.uint32)
    ctr[:, 0] = np.asarray(index, dtype=np.uint64) & _MASK
    ctr[:, 1] = toy & _MASK
    ctr[:, 2] = toy >> np.uint64(32)
    ctr[:, 3] = domain
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return uniform53(r[:, 0], r[:, 1]), uniform53(r[:, 2], r[:, 3])


def toy_events(edges, cdf, counts, seed=0, first_toy=0):
    """Events of the toys given their per-source counts [T, S].  Returns (coords [n_space, N], source [N], offsets)."""
    counts = np.asarray(counts, dtype=np.int64)
    n_toys, n_sources = counts.shape
    per_toy = counts.sum(axis=1)
    offsets = np.concatenate([[0], np.cumsum(per_toy)]).astype(np.int64)
    n = int(offsets[-1])
    toy = np.repeat(np.arange(n_toys, dtype=np.int64), per_toy)
    j = np.arange(n, dtype=np.int64) - offsets[toy]
    source = np.concatenate([np.repeat(np.arange(n_sources), counts[t]) for t in range(n_toys)]) if n else np.zeros(0, dtype=int)
    shape = tuple(len(e) - 1 for e in edges)
    u_bin, u0 = toy_uniforms(seed, first_toy + toy, j, 1)
    us = [u0]
    if len(edges) > 1:
        us += list(toy_uniforms(seed, first_toy + toy, j, 2))
    if len(edges) > 3:
        us.append(toy_uniforms(seed, first_toy + toy, j, 3)[0])
    coords = np.empty((len(edges), n))
    flat = np.empty(n, dtype=np.int64)
    for s in range(n_sources):
        m = source == s
        flat[m] = np.minimum(np.searchsorted(cdf[s], u_bin[m]), cdf.shape[1] - 1)
    multi = np.unravel_index(flat, shape)
    for d, e in enumerate(edges):
        lo, hi = e[multi[d]], e[multi[d] + 1]
        coords[d] = lo + us[d] * (hi - lo)
    return coords, source.astype(np.int32), offsets
