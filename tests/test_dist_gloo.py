"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups.

The per-device partial scores come from the ORACLE here (no GPU in this tier of tests); what is
under test is the product's partitioning and combine code (distributed-decisiontrees_b200/sharding.py):
ensemble chunks as PCIeReceiver cuts them, ring-order vs one-collective combine, data shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ddt_b200 as ddt
from helpers import oracle_cfg
from oracle import oracle as O

L, S = ddt.layout, ddt.sharding

T, D, F, K, N = 48, 5, 32, 2, 203


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _full_reference():
    W, FI = L.synth_ensemble(T, D, F, seed=11)
    x = L.synth_tuples(0, N, F, seed=12)
    return W, FI, x


def _worker(rank, world, port, mode, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, FI, x = _full_reference()
        if mode in ("reduce", "ring"):
            # ensemble-sharded: contiguous chunk of trees, every rank sees every tuple
            first, count = S.ensemble_chunk(T, rank, world)
            Kd, Sd = S.shard_geometry(T, D, K, world)
            wl, fl = L.pack_streams(W[first:first + count], FI[first:first + count], D)
            part = O.scores(oracle_cfg(D, Kd, Sd, L.MISSING_DEFAULT, F, count), wl, fl, x)
            t = torch.from_numpy(part.view(np.float32).copy())
            add = (lambda a, b: torch.from_numpy(O.fpadd_many(a.numpy().view(np.uint32), b.numpy().view(np.uint32)).view(np.float32)))
            out = S.combine_partials(t, dist, mode=mode, dst=0, add=add)
            if rank == 0:
                np.save(os.path.join(outdir, "out.npy"), out.numpy())
            else:
                assert out is None
            np.save(os.path.join(outdir, "part%d.npy" % rank), part)
        else:
            # data-sharded: whole ensemble everywhere, contiguous tuple shards, no collective
            first, count = S.data_shard(N, rank, world)
            wl, fl = L.pack_streams(W, FI, D)
            sc = O.scores(oracle_cfg(D, K, -(-T // (8 * K)), L.MISSING_DEFAULT, F, T), wl, fl, x[first:first + count])
            np.save(os.path.join(outdir, "shard%d.npy" % rank), sc)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "ring"), (2, "reduce"), (3, "ring"), (2, "data")])
def test_sharded_modes(tmp_path, world, mode):
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)
    W, FI, x = _full_reference()
    if mode == "data":
        got = np.concatenate([np.load(tmp_path / ("shard%d.npy" % r)) for r in range(world)])
        wl, fl = L.pack_streams(W, FI, D)
        want = O.scores(oracle_cfg(D, K, -(-T // (8 * K)), L.MISSING_DEFAULT, F, T), wl, fl, x)
        assert (got == want).all()               # bit-exact, global tuple order kept
        return
    parts = [np.load(tmp_path / ("part%d.npy" % r)) for r in range(world)]
    want = O.ring_combine(parts)                 # the reference's ring order, host first
    got = np.load(tmp_path / "out.npy").view(np.uint32)
    if mode == "ring":
        assert (got == want).all()               # bit-exact scores and therefore labels
    else:
        g, w = got.view(np.float32), want.view(np.float32)
        assert np.allclose(g, w, rtol=1e-5, atol=1e-7)     # north_star tolerance for the one-collective combine
    assert (O.labels(got) == O.labels(want)).mean() > 0.99


def test_ensemble_chunk_never_hands_out_an_empty_shard():
    """(first > 0, count = 0) used to reach dte_load_ensemble on trailing ranks (T < world * ceil(T / world)),
    which rejects it while the peers wait in a collective.  Now every rank gets the same ValueError up front."""
    for T, world in [(48, 3), (50, 3), (1024, 8), (8, 8), (9, 8)]:
        try:
            chunks = [S.ensemble_chunk(T, r, world) for r in range(world)]
        except ValueError:
            assert T == 9                      # ceil(9/8) = 2 trees per chunk would leave ranks 5..7 empty
            continue
        assert all(c > 0 for _, c in chunks) and sum(c for _, c in chunks) == T
        assert [f for f, _ in chunks] == [sum(c for _, c in chunks[:r]) for r in range(world)]
    for T, world in [(2, 3), (5, 4), (9, 8)]:
        for r in range(world):                 # the SAME decision on every rank
            with pytest.raises(ValueError):
                S.ensemble_chunk(T, r, world)
