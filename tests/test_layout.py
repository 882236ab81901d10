"""Host-side layout logic: stream packing, synthetic generators, sharding arithmetic."""
import numpy as np

import ddt_b200 as ddt

L, S = ddt.layout, ddt.sharding


def test_tree_cls_formula():
    assert L.tree_cls(4) == (8, 2) and L.tree_cls(12) == (2048, 512) and L.tree_cls(8) == (128, 32)
    assert L.tree_cls(1) == (1, 1) and L.tree_cls(2) == (2, 1) and L.tree_cls(3) == (4, 1)
    for d in range(3, 13):                       # 10 * 2^D bytes per tree for D >= 3
        w, f = L.tree_cls(d)
        assert (w + f) * 16 == 10 * 2 ** d


def test_pack_unpack_roundtrip_and_word_order():
    W, FI = L.synth_ensemble(5, 3, 16)
    wl, fl = L.pack_streams(W, FI, 3)
    assert wl.shape == (5 * 4, 4) and fl.shape == (5 * 1, 8)
    # little-endian: word i of a line at byte offset 4i, index i at byte offset 2i (PipelinedMUX.sv:63-65)
    raw = wl.tobytes()
    assert int.from_bytes(raw[4:8], "little") == int(W[0, 1])
    assert int.from_bytes(fl.tobytes()[2:4], "little") == int(FI[0, 1])
    W2, FI2 = L.unpack_streams(wl, fl, 3)
    assert (W2 == W).all() and (FI2 == FI).all()
    # padding words are zero
    assert wl.reshape(5, 16)[:, 15].sum() == 0 and fl.reshape(5, 8)[:, 7].sum() == 0


def test_synth_is_deterministic_and_in_contract():
    a = L.synth_tuples(100, 50, 32)
    b = L.synth_tuples(0, 200, 32)[100:150]
    assert (a == b).all()                        # counter based: any window regenerates identically
    v = a.view(np.float32)
    miss = a == L.MISSING_DEFAULT
    assert ((v >= 0) & (v < 1))[~miss].all()
    m = L.synth_tuples(0, 4000, 64, missing_ppm=50000)
    frac = (m == L.MISSING_DEFAULT).mean()
    assert 0.04 < frac < 0.06
    W, FI = L.synth_ensemble(32, 6, 40)
    assert ((FI & 0x7FF) < 40).all() and (FI & 0x4000).sum() == 0 and 0 < ((FI >> 13) & 1).mean() < 1
    leaves = W[:, 63:].view(np.float32)
    assert np.isfinite(leaves).all() and (np.abs(leaves) < 2.0 / 32).all()
    # golden pin of the generator itself
    assert int(L.splitmix64(0, np.arange(1, dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF


def test_result_lines_drop_partial_line():
    s = np.arange(10, dtype=np.float32)
    r = L.result_lines(s)
    assert r.shape == (2, 4) and r[1, 3] == 7.0


def test_sharding_arithmetic():
    assert [S.ensemble_chunk(8192, g, 8) for g in range(8)] == [(1024 * g, 1024) for g in range(8)]
    chunks = [S.ensemble_chunk(10, g, 4) for g in range(4)]
    assert chunks == [(0, 3), (3, 3), (6, 3), (9, 1)]
    assert sum(c for _, c in chunks) == 10
    shards = [S.data_shard(10, g, 4) for g in range(4)]
    assert shards == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert S.shard_geometry(8192, 10, 8, 8) == (8, 16)
    deal = S.deal_batches(10, 4, 3)
    assert deal == [[(0, 4)], [(4, 4)], [(8, 2)]]
    deal = S.deal_batches(20, 4, 2)
    assert deal[0] == [(0, 4), (8, 4), (16, 4)] and deal[1] == [(4, 4), (12, 4)]


def test_performance_model_cli():
    """tools/dte_model.cpp — the B200 counterpart of profiler/profiler.cpp; same three inputs."""
    import re
    import subprocess
    from ddt_b200 import build as B
    exe = B.build_model()
    out = subprocess.run([exe, "512", "12", "128", "8"], capture_output=True, text=True)
    assert out.returncode == 0
    txt = out.stdout
    # the reference's law f*Ncu*Npe/(depth*Ntrees) with 64 PEs at 150 MHz, D=12, 512 trees
    m = re.search(r"engine\s*:\s*([0-9.e+]+) tuples/s per FPGA", txt)
    assert m and abs(float(m.group(1)) - 150e6 * 64 / (12 * 512)) / (150e6 * 64 / (12 * 512)) < 1e-3
    m = re.search(r"walk-bound\s*:\s*([0-9.e+]+)", txt)
    assert m and float(m.group(1)) > 1e7
    assert subprocess.run([exe], capture_output=True).returncode == 2


def test_repacker_on_the_cpu_against_the_oracle(tmp_path):
    """The device layout (8-byte top records, 32/64-byte bottom records, D = 1 extension, early leaves expanded into
    complete subtrees) decoded on the HOST exactly as the kernels decode it (tools/pack_check.cu) gives the oracle's
    leaf for every (tuple, tree) — the repacker is covered without a GPU."""
    import subprocess
    import numpy as np
    from ddt_b200 import build as B
    from helpers import oracle_cfg
    from oracle import oracle as O
    exe = B.build_pack_check()
    rng = np.random.default_rng(3)
    for D, T, F, early in [(1, 5, 4, False), (2, 9, 8, False), (3, 8, 8, True), (5, 16, 32, True), (8, 7, 600, True), (12, 3, 256, True), (10, 4, 2044, False)]:
        W, FI = L.synth_ensemble(T, D, F, seed=50 + D)
        n_int = (1 << D) - 1
        if early:
            for t in range(T):
                for _ in range(3):
                    lvl = int(rng.integers(0, D))
                    FI[t, (1 << lvl) - 1 + int(rng.integers(0, 1 << lvl))] |= 1 << 14
        x = L.synth_tuples(0, 97, F, seed=60 + D, missing_ppm=30000)
        wl, fl = L.pack_streams(W, FI, D)
        for name, arr in (("w", wl), ("f", fl), ("x", x)):
            np.ascontiguousarray(arr).tofile(tmp_path / (name + ".bin"))
        out = subprocess.run([exe, str(D), str(F), str(T), "97", str(L.MISSING_DEFAULT), str(tmp_path / "w.bin"), str(tmp_path / "f.bin"),
                              str(tmp_path / "x.bin")], capture_output=True, timeout=120)
        assert out.returncode == 0, out.stderr
        got = np.frombuffer(out.stdout, dtype=np.uint32).reshape(97, T)
        cfg = oracle_cfg(D, 1, 1, L.MISSING_DEFAULT, F, T)
        w_cls, f_cls = L.tree_cls(D)
        lib = O.lib()
        import ctypes as C
        for t in range(T):
            wt = np.ascontiguousarray(wl.reshape(T, -1)[t]); ft = np.ascontiguousarray(fl.reshape(T, -1)[t])
            for i in range(0, 97, 7):
                xi = np.ascontiguousarray(x[i])
                want = lib.dteo_leaf(C.byref(cfg), wt.ctypes.data, ft.ctypes.data, xi.ctypes.data)
                assert got[i, t] == want, (D, t, i)


def test_partition_arithmetic_against_a_brute_force_deal():
    """csrc/dte_partition.hpp (global tuple -> ring position / local index; covered prefix from per-device counts) against a
    literal simulation of PCIeReceiver's batch dealing (PCIeReceiver.sv:298-307) over random (batch, devices, tuples)."""
    import subprocess
    from ddt_b200 import build as B
    out = subprocess.run([B.build_partition_check()], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_dealing_and_chunking_kats(kats):
    """Hand-derived vectors for the two partition rules of PCIeReceiver (batch dealing :298-307, tree chunks :241-264)
    against the host-side helpers the multi-process path uses."""
    import ddt_b200 as ddt
    for k in kats["deal"]:
        got = ddt.sharding.deal_batches(k["n_lines"], k["batch_cls"], k["num_devs"])
        assert [[list(x) for x in dev] for dev in got] == k["expect"], k["name"]
    for k in kats["chunks"]:
        got = [list(ddt.sharding.ensemble_chunk(k["n_trees"], r, k["num_devs"])) for r in range(k["num_devs"])]
        assert got == k["expect"], k["name"]
