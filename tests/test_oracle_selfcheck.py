"""Independent restatements inside the oracle must agree with each other:
adder model A (host float RNE + FTZ) vs model B (bit-level FloPoCo datapath), and the clean heap
walker vs the PU-memory address-literal walker, on random data."""
import numpy as np
import pytest

from helpers import oracle_cfg, L
from oracle import oracle as O


def _normals(rng, n, emin, emax):
    e = rng.integers(emin, emax, n).astype(np.uint32)
    f = rng.integers(0, 1 << 23, n).astype(np.uint32)
    s = rng.integers(0, 2, n).astype(np.uint32)
    return (s << 31) | (e << 23) | f


def test_adder_models_agree_random():
    rng = np.random.default_rng(7)
    n = 60000
    a, b = _normals(rng, n, 60, 200), _normals(rng, n, 60, 200)
    assert (O.fpadd_many(a, b) == O.fpadd_many(a, b, literal=True)).all()
    # close exponents, opposite signs: cancellation, renormalisation, sticky/round paths
    b2 = (a ^ np.uint32(0x80000000)) + rng.integers(-4, 5, n).astype(np.int32).view(np.uint32)
    assert (O.fpadd_many(a, b2) == O.fpadd_many(a, b2, literal=True)).all()
    # exponent differences around the 24..27 alignment boundary
    for d in range(20, 30):
        bb = (a & np.uint32(0x807FFFFF)) | (((a >> 23) & 0xFF) - d).astype(np.uint32) << 23
        assert (O.fpadd_many(a, bb) == O.fpadd_many(a, bb, literal=True)).all()


def test_adder_special_values():
    one, mone = 0x3F800000, 0xBF800000
    for lit in (False, True):
        assert O.fpadd(0, 0, lit) == 0
        assert O.fpadd(one, 0, lit) == one
        assert O.fpadd(0, mone, lit) == mone
        assert O.fpadd(one, mone, lit) == 0                       # +0, never -0
        assert O.fpadd(0x3F800001, 0xBF800000, lit) == 0x34000000  # 2^-23 exactly
        assert O.fpadd(0x4B800000, one, lit) == 0x4B800000        # 2^24 + 1 ties to even
        assert O.fpadd(0x4B800001, one, lit) == 0x4B800002        # 2^24+2 + 1 ties to even (up)
    # commutative
    rng = np.random.default_rng(3)
    a, b = _normals(rng, 2000, 100, 150), _normals(rng, 2000, 100, 150)
    assert (O.fpadd_many(a, b, True) == O.fpadd_many(b, a, True)).all()


@pytest.mark.parametrize("D,T,F,K,S", [
    (4, 16, 32, 2, 1), (4, 16, 32, 1, 2), (1, 8, 4, 1, 1), (2, 24, 8, 1, 3), (3, 64, 16, 8, 1),
    (6, 40, 64, 4, 2), (8, 64, 128, 8, 1), (10, 16, 256, 2, 1), (12, 8, 64, 1, 1), (5, 13, 2044, 2, 1),
])
def test_walkers_agree_random(D, T, F, K, S):
    rng = np.random.default_rng(D * 1000 + T)
    for negative in (False, True):
        W, FI = L.synth_ensemble(T, D, F, seed=int(rng.integers(1 << 40)), negative=negative)
        x = L.synth_tuples(0, 257, F, seed=int(rng.integers(1 << 40)), missing_ppm=30000, signed=negative)
        wl, fl = L.pack_streams(W, FI, D)
        cfg = oracle_cfg(D, K, S, L.MISSING_DEFAULT, F, T)
        a = O.scores(cfg, wl, fl, x)
        b = O.scores(cfg, wl, fl, x, literal_adder=True)
        c = O.scores_literal(cfg, wl, fl, x)
        assert (a == b).all() and (a == c).all()


def test_cfg1_shape_and_threads():
    # BASELINE cfg1: 16 trees, depth 4, 32 features, 10k tuples
    W, FI = L.synth_ensemble(16, 4, 32)
    x = L.synth_tuples(0, 10000, 32)
    wl, fl = L.pack_streams(W, FI, 4)
    cfg = oracle_cfg(4, 2, 1, L.MISSING_DEFAULT, 32, 16)
    s1 = O.scores(cfg, wl, fl, x)
    s4 = O.scores(cfg, wl, fl, x, threads=4)
    assert (s1 == s4).all()
    lab = O.labels(s1)
    assert 0 < lab.sum() < lab.size          # both labels occur
    # golden checksum of the oracle on the canonical synthetic set (pins generator + oracle together)
    assert int(s1.astype(np.uint64).sum()) == int(np.load(_golden("cfg1_scores_sum.npy")))


def _golden(name):
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)


def test_blocked_loop_is_bit_identical_and_thread_count_is_bounded():
    """The tree-blocked loop (the one timed as the CPU baseline) performs the same adds in the same per-tuple order."""
    rng = np.random.default_rng(7)
    for (D, T, F, K, S, n) in [(4, 16, 32, 2, 1, 1000), (6, 37, 64, 4, 2, 333), (9, 100, 128, 8, 2, 130), (3, 5, 8, 1, 1, 65),
                               (12, 24, 256, 8, 1, 200), (5, 64, 16, 2, 2, 64)]:     # S*K*8 < T: trailing trees never walked
        W, FI = L.synth_ensemble(T, D, F, seed=int(rng.integers(1 << 30)), bias=0.0)
        x = L.synth_tuples(0, n, F, seed=int(rng.integers(1 << 30)), missing_ppm=30000)
        wl, fl = L.pack_streams(W, FI, D)
        cfg = oracle_cfg(D, K, S, L.MISSING_DEFAULT, F, T)
        want = O.scores(cfg, wl, fl, x)
        for th in (1, 3):
            assert (O.scores_blocked(cfg, wl, fl, x, threads=th) == want).all(), (D, T, th)
    assert 1 <= O.max_threads() <= O.online_cpus()
