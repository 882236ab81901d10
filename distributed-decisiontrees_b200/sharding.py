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
    """Contiguous tree range [first, first+count) of device `rank` out of `world`: chunks of ceil(T/world) trees in
    ring order, as PCIeReceiver cuts the streams by numcls_local_weights (PCIeReceiver.sv:241-264).  Every device
    must end up with at least one tree — (first>0, count=0) would read as "all trees" downstream and a rank that
    raises while its peers wait in a collective hangs the job — so that is refused here, on every rank alike."""
    n_trees, world = int(n_trees), int(world)
    per = -(-n_trees // world)
    if per * (world - 1) >= n_trees:
        raise ValueError("%d trees cannot be cut into %d non-empty chunks of %d" % (n_trees, world, per))
    first = rank * per
    return first, min(n_trees, first + per) - first


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


class RingCombine:
    """Ensemble-sharded combine in the reference's RING ORDER at the cost of one kernel (one process per GPU).

    Every rank's walk writes its partial scores into a peer-visible buffer (CUDA IPC through libdte.so); rank `dst`
    (the host node, ring position 0) runs `ring_combine_kernel` over all of them — peers are read over NVLink — and
    forms score = ((p0 + p1) + p2) + ... with add.rn.ftz.f32, exactly ResultsCombiner.sv:292-311,359-368, then the
    labels.  torch.distributed only carries the 64-byte handles and the two host barriers of a step."""

    def __init__(self, engine, dist, n):
        import torch
        self.engine, self.dist, self.n = engine, dist, int(n)
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        err, handle = None, None
        try:
            self.ptr, handle = engine.ipc_alloc(4 * self.n)
        except Exception as ex:                          # noqa: BLE001
            self.ptr, err = None, "rank %d: %s" % (self.rank, ex)
        boxes = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(boxes, (handle, err))
        else:
            boxes = [(handle, err)]
        err = next((b[1] for b in boxes if b[1]), None)
        self.ptrs = None
        if err is None and self.rank == 0:
            try:
                self.ptrs = [self.ptr] + [engine.ipc_open(boxes[r][0]) for r in range(1, self.world)]
            except Exception as ex:                      # noqa: BLE001
                err = "rank 0: %s" % ex
        if self.world > 1:
            flags = [None] * self.world
            dist.all_gather_object(flags, err)
            err = next((f for f in flags if f), None)
        if err:
            self.close()
            raise RuntimeError("ring combine unavailable (%s)" % err)
        self.part = torch.as_tensor(_CudaArray(self.ptr, self.n), device="cuda")

    def step(self, d_tuples, n, stream, d_out=None, d_labels=None):
        """walk -> barrier -> (rank 0) ONE combine kernel -> barrier.  Rank 0 gets scores in d_out, labels in d_labels."""
        import torch
        self.engine.infer_device(d_tuples, n, self.ptr, None, stream=stream)
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        if self.rank == 0:
            self.engine.ring_combine_device(self.ptrs, n, d_out, d_labels, stream=stream)
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()                          # the peers may now overwrite their partials

    def close(self):
        self.part = None
        if getattr(self, "ptrs", None):
            for p in self.ptrs[1:]:
                try:
                    self.engine.ipc_close(p, False)
                except Exception:                        # noqa: BLE001
                    pass
            self.ptrs = None
        if getattr(self, "ptr", None) is not None:
            try:
                self.engine.ipc_close(self.ptr, True)
            except Exception:                            # noqa: BLE001
                pass
            self.ptr = None


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
        self.engine.infer_device_accumulate(d_tuples, n, self.ptr, stream=stream)     # stream: torch handle (0 = legacy default)
        torch.cuda.synchronize()
        if self.dist is not None and self.dist.get_world_size() > 1:
            self.dist.barrier()
        return self.view[:n] if self.owner else None

    def close(self):
        self.view = None
        self.engine.ipc_close(self.ptr, self.owner)
