#!/usr/bin/env python3
"""Where does the multi-device handle lose e2e throughput?  (diagnostic, run under gpurun --gpus 2)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ddt_b200 as ddt
from ddt_b200 import engine as E
L = ddt.layout
T, D, F, K, n = 1024, 12, 256, 8, 4_000_000
W, FI = L.synth_ensemble(T, D, F); wl, fl = L.pack_streams(W, FI, D)
torch.cuda.set_device(0)
hx_t = torch.empty((n, F), dtype=torch.int32, pin_memory=True); hx_t.random_(0, 1 << 30)
def regs_for(eng, G, mode):
    regs = E.csr_from_profile(T, D, 4 * F, K, L.MISSING_DEFAULT, n)
    if G > 1:
        regs[201] = (0x2 | 0x20 | 0x40 | (0x8 if mode == "data" else 0x14)) | ((4096 * (F // 4)) << 32)
        regs[203] = G << 32
    for a, v in sorted(regs.items()): eng.softreg_write(a, v)
def run(devs, src, label):
    eng = ddt.Engine(devs if len(devs) > 1 else devs[0])
    regs_for(eng, len(devs), "data")
    eng.load_ensemble(wl, fl)
    if src == "dte":
        hx = eng.host_alloc((n, F), np.int32); hx[:] = hx_t.numpy()
        hs = eng.host_alloc((n,), np.float32)
    else:
        hx, hs = hx_t, torch.empty(n, dtype=torch.float32, pin_memory=True)
    for _ in range(2): eng.infer_host(hx, want_labels=False, out_scores=hs)
    t0 = time.perf_counter()
    for _ in range(3): eng.infer_host(hx, want_labels=False, out_scores=hs)
    dt = (time.perf_counter() - t0) / 3
    print("%-28s devs=%s src=%s: %.1f M tuples/s" % (label, devs, src, n / dt / 1e6), flush=True)
    eng.close()
ng = torch.cuda.device_count()
run([0], "torch", "single gpu0")
if ng > 1:
    run([1], "torch", "single gpu1, torch-pinned")
    run([1], "dte", "single gpu1, dte_host_alloc")
    run([0, 1], "torch", "multi, torch-pinned")
    run([0, 1], "dte", "multi, dte_host_alloc")
