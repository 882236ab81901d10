"""One-off randomised parity stress (GPU): python tools/stress_parity.py <seed> <configs>.
Every kernel variant against the oracle, raw 32-bit feature/threshold words, random geometry.
Round 1: seed 12345, 200 configs x 4 variants -> 0 mismatches."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import ddt_b200 as ddt
from ddt_b200 import engine as E
from helpers import oracle_cfg, geometry_regs, L
from oracle import oracle as O
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
bad = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 150):
    D = int(rng.integers(1, 13)); K = int(rng.choice([1, 2, 4, 8])); T = int(rng.integers(1, 120))
    S = max(1, -(-T // (8 * K)) + int(rng.integers(-1, 2))); F = 4 * int(rng.integers(1, 130)); n = int(rng.integers(1, 1500))
    if rng.random() < 0.15: F = 4 * int(rng.integers(130, 512))
    W, FI = L.synth_ensemble(T, D, F, seed=int(rng.integers(1 << 40)))
    n_int = (1 << D) - 1
    W[:, :n_int] = rng.integers(0, 1 << 32, size=(T, n_int), dtype=np.uint64).astype(np.uint32)
    x = rng.integers(0, 1 << 32, size=(n, F), dtype=np.uint64).astype(np.uint32)
    missing = int(rng.integers(0, 1 << 32)); x[rng.random((n, F)) < 0.03] = missing
    wl, fl = L.pack_streams(W, FI, D)
    want = O.scores(oracle_cfg(D, K, S, missing, F, T), wl, fl, x, threads=16)
    e = ddt.Engine(0)
    for a, v in sorted(geometry_regs(T, D, F, K, S, missing).items()): e.softreg_write(a, v)
    e.load_ensemble(wl, fl)
    for v in (0, 1, 2, 3):
        e.set_kernel_variant(v)
        sc, lb = e.infer_host(x)
        if not (sc.view(np.uint32) == want).all() or not (lb == O.labels(want)).all():
            bad += 1; print("MISMATCH", it, D, T, F, K, S, n, v)
    e.close()
print("stress done, mismatches:", bad)
