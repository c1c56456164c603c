"""The tile-parallel cumulative sum of csrc/hmx_stage.cuh (K4a), restated in NumPy: while the running sum stays inside one
binade and no addend is an exact rounding tie, RN(s + w) = (S + RN(w / ulp)) ulp, so a tile's prefixes are an INTEGER prefix
sum; everything else takes the literal chain.  Must equal np.cumsum bit for bit (the GPU kernel is tested the same way in
tests/test_gpu_stage.py)."""
import numpy as np

TILE = 4096


def _round_weights(w, e):
    """q = RN(w / 2^(e-1075)) per weight; bad where the parallel path may not be used."""
    wb = w.view(np.uint64)
    ew = ((wb >> np.uint64(52)) & np.uint64(0x7FF)).astype(np.int64)
    bad = ((wb >> np.uint64(63)) != 0) | (ew == 0x7FF)
    mw = (wb & np.uint64((1 << 52) - 1)) | np.where(ew > 0, np.uint64(1 << 52), np.uint64(0))
    sh = e - np.where(ew > 0, ew, 1)
    bad |= sh < 2
    shc = np.clip(sh, 1, 63).astype(np.uint64)
    fl = mw >> shc
    rem = mw - (fl << shc)
    half = np.uint64(1) << (shc - np.uint64(1))
    q = np.where(sh >= 64, np.uint64(0), fl + (rem > half).astype(np.uint64))
    bad |= (sh < 64) & (rem == half)
    return q, bad


def cumsum_tiled(w):
    out = np.empty_like(w)
    s = 0.0
    n_fast = n_chain = 0
    for a in range(0, len(w), TILE):
        t = w[a:a + TILE]
        sb = int(np.float64(s).view(np.uint64))
        e = (sb >> 52) & 0x7FF
        fast = not (sb >> 63) and 0 < e < 0x7FF
        if fast:
            S = (sb & ((1 << 52) - 1)) | (1 << 52)
            q, bad = _round_weights(t, e)
            pre = np.cumsum(q.astype(np.uint64))
            total = int(np.sum(q.astype(object)))
            fast = not bad.any() and total < (1 << 53) and S + total < (1 << 53)
        if fast:
            tot = np.uint64(S) + pre
            out[a:a + TILE] = ((np.uint64(e) << np.uint64(52)) | (tot & np.uint64((1 << 52) - 1))).view(np.float64)
            n_fast += 1
        else:
            acc = np.float64(s)
            for i, x in enumerate(t):
                acc = acc + x
                out[a + i] = acc
            n_chain += 1
        s = float(out[min(a + TILE, len(w)) - 1])
    return out, n_fast, n_chain


def test_tiled_cumsum_equals_numpy_bitwise():
    rng = np.random.default_rng(0)
    cases = {}
    w = rng.random(120001)
    cases["uniform"] = w / w.sum()
    w = np.exp(rng.normal(size=100000) * 6)
    cases["lognormal"] = w / w.sum()
    w = np.sort(np.exp(rng.normal(size=40000) * 10))
    cases["ascending"] = w / w.sum()
    cases["descending"] = cases["ascending"][::-1].copy()
    cases["constant"] = np.full(50000, 1.0 / 50000)
    w = rng.random(40000)
    w[rng.random(40000) < 0.3] = 0.0
    cases["zeros"] = w / w.sum()
    cases["ties"] = np.concatenate([[0.6], (rng.integers(1, 1000, 20000) + 0.5) * 2.0 ** -53, rng.random(1000) * 1e-6])
    cases["subnormal"] = np.concatenate([np.full(3000, 5e-324), np.full(3000, 1e-310), rng.random(20000) * 1e-5])
    cases["jumps"] = np.concatenate([rng.random(9000) * 1e-9, [0.7], rng.random(9000) * 1e-9, [0.2999], rng.random(9000) * 1e-7])
    cases["heavy"] = np.concatenate([[1e-3], np.full(5000, 4e-4), rng.random(9000) * 1e-6])
    fast_total = 0
    for name, w in cases.items():
        w = np.ascontiguousarray(w, dtype=np.float64)
        got, n_fast, n_chain = cumsum_tiled(w)
        assert np.array_equal(got.view(np.uint64), np.cumsum(w).view(np.uint64)), name
        fast_total += n_fast
        if name in ("uniform", "lognormal", "constant", "descending"):
            assert n_fast > n_chain, (name, n_fast, n_chain)  # the parallel path carries the bulk (more so at larger N)
        if name == "ties":
            assert n_chain >= 5  # the tiles that hold exact ties are decided by the chain
    assert fast_total > 50
