#!/usr/bin/env python3
"""dte_autotune on the BASELINE geometries: which launch plan does measurement pick, and what does the built-in choice cost?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddt_b200 as ddt
L = ddt.layout
for name, (T, D, F) in (("cfg2", (512, 8, 128)), ("cfg3", (1024, 12, 256)), ("cfg4 shard", (1024, 10, 256)), ("other", (256, 11, 512))):
    W, FI = L.synth_ensemble(T, D, F)
    wl, fl = L.pack_streams(W, FI, D)
    with ddt.Engine(0) as e:
        e.configure(T, D, 4 * F, clusters=8)
        e.load_ensemble(wl, fl)
        print("== %s: %d trees, D=%d, F=%d; built-in plan: %s" % (name, T, D, F, e.kernel_name()))
        print(e.autotune(4_000_000))
