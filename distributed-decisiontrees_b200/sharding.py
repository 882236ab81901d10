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


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can view a raw device pointer (zero copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (int(n),), "typestr": "<f4", "version": 2}


class FusedCombine:
    """Ensemble-sharded combine fused into the walk kernel.

    Rank `dst` owns a peer-visible fp32[n] buffer (CUDA IPC through libdte.so); every rank's walk
    kernel adds its partial scores straight into it over NVLink (`red.relaxed.sys.global.add.f32` in
    the epilogue) — no separate reduce kernel, no collective on the data path.  torch.distributed is
    used only to hand the 64-byte handle around and for the two host barriers of a step."""

    def __init__(self, engine, dist, n, dst=0):
        import torch
        self.engine, self.dist, self.n, self.dst = engine, dist, int(n), dst
        multi = dist is not None and dist.get_world_size() > 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.owner = self.rank == dst
        self.ptr = None
        # every rank takes the same decision (no rank may be left waiting in a collective)
        box = [None]
        if self.owner:
            try:
                self.ptr, handle = engine.ipc_alloc(4 * self.n)
                box[0] = handle
            except Exception as ex:                      # noqa: BLE001
                box[0] = "error: %s" % ex
        if multi:
            dist.broadcast_object_list(box, src=dst)
        err = box[0] if isinstance(box[0], str) else None
        if err is None and not self.owner:
            try:
                self.ptr = engine.ipc_open(box[0])
            except Exception as ex:                      # noqa: BLE001
                err = "error: %s" % ex
        if multi:
            flags = [None] * dist.get_world_size()
            dist.all_gather_object(flags, err)
            err = next((f for f in flags if f), None)
        if err:
            if self.ptr is not None:
                try:
                    engine.ipc_close(self.ptr, self.owner)
                except Exception:                        # noqa: BLE001
                    pass
                self.ptr = None
            raise RuntimeError("fused combine unavailable (%s)" % err)
        self.view = torch.as_tensor(_CudaArray(self.ptr, self.n), device="cuda") if self.owner else None

    def step(self, d_tuples, n, stream):
        """One combined inference: returns the summed scores (torch view) on dst, None elsewhere."""
        import torch
        if self.owner:
            self.view[:n].zero_()
        torch.cuda.synchronize()
        if self.dist is not None and self.dist.get_world_size() > 1:
            self.dist.barrier()
        self.engine.infer_device_accumulate(d_tuples, n, self.ptr, stream=stream)
        torch.cuda.synchronize()
        if self.dist is not None and self.dist.get_world_size() > 1:
            self.dist.barrier()
        return self.view[:n] if self.owner else None

    def close(self):
        self.view = None
        self.engine.ipc_close(self.ptr, self.owner)
