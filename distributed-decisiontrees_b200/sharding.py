"""Multi-GPU partitioning — the InputDistributor / ResultsCombiner semantics, one process per GPU.

Reference (paths relative to the reference root):
  * ensemble-sharded (broadcast_data=1, aggreg_enabled=1): device g keeps a CONTIGUOUS chunk of
    trees, the weights stream cut every `numcls_local_weights` lines and the index stream every
    `numcls_local_findexes` lines (rtl/DTEngine/PCIeReceiver.sv:241-264); every device sees every
    tuple (rtl/DTEngine/InputDistributor.sv:199-204); partial scores are summed over the ring in
    device order, host first (rtl/DTEngine/ResultsCombiner.sv:292-311,359-368).
  * data-sharded (broadcast_trees=1): every device holds the whole ensemble, tuples are dealt out
    (rtl/DTEngine/PCIeReceiver.sv:298-307) and results are forwarded unsummed
    (rtl/DTEngine/ResultsCombiner.sv:370-391).  No collective.

torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def ensemble_chunk(n_trees, rank, world):
    """Contiguous tree range [first, first+count) of device `rank` out of `world`."""
    per = -(-int(n_trees) // int(world))
    first = min(int(n_trees), rank * per)
    return first, max(0, min(int(n_trees), first + per) - first)


def data_shard(n_tuples, rank, world):
    """Contiguous tuple range [first, first+count) of device `rank` (global tuple order is kept)."""
    base, rem = divmod(int(n_tuples), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def shard_geometry(n_trees_total, depth_levels, clusters, world):
    """Per-device (K, S) for an ensemble shard: the same K, S sized for the local chunk."""
    per = -(-int(n_trees_total) // int(world))
    return int(clusters), -(-per // (8 * int(clusters)))


def combine_partials(partial, dist, mode="reduce", dst=0, add=None):
    """Sum per-device partial scores onto rank `dst`.

    mode="reduce": ONE collective, dist.reduce(SUM) (NCCL over NVLink on GPUs) — summation order is
                   the library's, so scores match the ring order to ~1 ulp per hop, not bit-exactly.
    mode="ring":   gather to dst, then ((p0 + p1) + p2) + ... in device order with `add`
                   (Engine.ring_add_device on GPUs) — bit-exact with the reference's ring.
    Returns the combined tensor on rank dst, None elsewhere."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return partial
    if mode == "reduce":
        out = partial.clone()
        dist.reduce(out, dst=dst, op=dist.ReduceOp.SUM)
        return out if rank == dst else None
    if mode != "ring":
        raise ValueError(mode)
    bufs = [torch.empty_like(partial) for _ in range(world)] if rank == dst else None
    dist.gather(partial, gather_list=bufs, dst=dst)
    if rank != dst:
        return None
    acc = bufs[0]
    for g in range(1, world):
        acc = add(bufs[g], acc) if add is not None else bufs[g] + acc
    return acc


def deal_batches(n_lines, batch_cls, world):
    """The reference's round-robin deal of data lines in batches of core_data_batch_cls
    (PCIeReceiver.sv:298-307): returns, per device, the list of (first_line, n_lines) it receives."""
    out = [[] for _ in range(world)]
    dev, pos = 0, 0
    while pos < n_lines:
        take = min(batch_cls, n_lines - pos)
        out[dev].append((pos, take))
        pos += take
        dev = (dev + 1) % world
    return out
