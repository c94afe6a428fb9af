"""numpy restatement of the engine's variant sampler (dmx_qp_sample_kernel): Philox4x32-10 with counter
(sample lo, sample hi, slot // 4, 0) and key = seed; slot s takes word s % 4 as u = (word + 0.5) / 2^32,
idx = first cdf entry > u.  Test infrastructure."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, seed):
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def sample_variants(qps, batch, seed, first_sample=0):
    S = len(qps)
    g = np.arange(first_sample, first_sample + batch, dtype=np.uint64)
    rows = np.zeros((batch, S), np.uint8)
    neg = np.zeros(batch, np.int64)
    zero = np.zeros(batch, bool)
    words = None
    for s, qp in enumerate(qps):
        qp = np.asarray(qp, np.float64)
        if s % 4 == 0:
            words = philox4x32_10(g & MASK, g >> np.uint64(32), np.full(batch, s // 4, np.uint64), np.zeros(batch, np.uint64), seed)
        u = (words[s % 4].astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
        total = np.sum(np.abs(qp))
        acc, cdf = 0.0, np.empty(len(qp))
        for i, v in enumerate(np.abs(qp)):      # same accumulation order as the C code
            acc += v / total
            cdf[i] = acc
        cdf[-1] = 1.0
        idx = np.minimum(np.searchsorted(cdf, u, side="right"), len(qp) - 1)
        rows[:, s] = idx
        neg += qp[idx] < 0
        zero |= qp[idx] == 0
    return rows, np.where(zero, 0.0, np.where(neg % 2 == 1, -1.0, 1.0))


def shot_expectations(probs, shots, masks, seed, first_row=0):
    """numpy restatement of dmx_shot_expect_kernel: per row, shot s takes word s % 4 of Philox call (s // 4 lo, s // 4 hi,
    row lo, row hi), outcome = first entry of cumsum(p) / sum(p) (last entry forced to 1) that exceeds u; every mask's
    expectation 2 P(even parity) - 1 comes from the same shots."""
    probs = np.asarray(probs, np.float64)
    rows, dim = probs.shape
    out = np.zeros((rows, len(masks)))
    parity = np.array([[bin(i & int(m)).count("1") & 1 for i in range(dim)] for m in masks])
    all_shots = np.broadcast_to(np.asarray(shots, np.int64), (rows,))
    for r in range(rows):
        shots = int(all_shots[r])
        calls = (shots + 3) // 4
        c = np.arange(calls, dtype=np.uint64)
        row = np.uint64(first_row + r)
        words = philox4x32_10(c & MASK, c >> np.uint64(32), np.full(calls, row & MASK, np.uint64), np.full(calls, row >> np.uint64(32), np.uint64), seed)
        u = (np.stack(words, axis=1).reshape(-1)[:shots].astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
        cdf = np.cumsum(probs[r])
        cdf = cdf / cdf[-1]
        cdf[-1] = 1.0
        outcome = np.minimum(np.searchsorted(cdf, u, side="right"), dim - 1)
        hist = np.bincount(outcome, minlength=dim)
        for j in range(len(masks)):
            out[r, j] = 2.0 * (hist[parity[j] == 0].sum() / shots) - 1.0
    return out
