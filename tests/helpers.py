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


def multi_node_regs(T, D, F, K, n, ndev, mode, batch_tuples=2):
    """Register values of a multi-device run as a Catapult host would write them (EngineCSR.sv:194-216)."""
    w_cls, f_cls = L.tree_cls(D)
    per = -(-T // ndev)
    from ddt_b200 import engine as E
    regs = E.csr_from_profile(T, D, 4 * F, K, L.MISSING_DEFAULT, n)
    flags = 0x2 | 0x20 | 0x40                       # host_node | multiple_nodes | pcie_receiver_enabled
    if mode == "ensemble":
        flags |= 0x4 | 0x10                         # broadcast_data | aggreg_enabled
        S = -(-per // (8 * K))
        regs[205] = (regs[205] & ~(0xFF << 36)) | (S << 36)
    else:
        flags |= 0x8                                # broadcast_trees
    regs[201] = flags | ((batch_tuples * (F // 4)) << 32)
    regs[203] = ((per * w_cls - 1) & 0xFFFF) | ((per * f_cls) << 16) | (ndev << 32)
    return regs
