"""Shared test helpers: build streams/config for a case and run the oracle on it."""
import numpy as np

import ddt_b200 as ddt
from oracle import oracle as O

L = ddt.layout


def kat_arrays(case):
    """KAT case -> (W uint32 [T, 2^(D+1)-1], FI uint16 [T, 2^D-1], tuples uint32 [n, F])."""
    W = np.array([t["W"] for t in case["trees"]], dtype=np.uint32)
    FI = np.array([t["FI"] for t in case["trees"]], dtype=np.uint16)
    x = np.array(case["tuples"], dtype=np.uint32)
    return W, FI, x


def oracle_cfg(D, K, S, missing, F, T):
    w_cls, f_cls = L.tree_cls(D)
    return O.make_cfg(D, K, S, missing, w_cls, f_cls, F // 4, T)


def oracle_scores(W, FI, x, D, K, S, missing, **kw):
    wl, fl = L.pack_streams(W, FI, D)
    cfg = oracle_cfg(D, K, S, missing, x.shape[1], W.shape[0])
    return O.scores(cfg, wl, fl, x, **kw)


def geometry_regs(T, D, F, K, S, missing, n_tuples=0):
    """Register values 201..208 for an explicit (K, S) — S may differ from the profile default."""
    from ddt_b200 import engine as E
    regs = E.csr_from_profile(T, D, 4 * F, K, missing, n_tuples)
    r205 = regs[205]
    r205 = (r205 & ~(0xFF << 36)) | ((S & 0xFF) << 36)
    regs[205] = r205
    return regs
