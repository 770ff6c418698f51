"""Multi-GPU plumbing for the barcode-sharded path (one process per GPU, torch.distributed).

Barcodes are independent in the reference (/root/reference/cmd_cram_demuxlet.cpp:576-875; its own way to
parallelise is --group-list, :68), so ranks own contiguous barcode ranges balanced by pileup entries, every
rank holds the whole genotype-probability matrix, and the only exchange of the path is one all-gather of the
fixed-size per-barcode best records.  Works on "nccl" (GPU tensors over NVLink) and on "gloo" (CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .plp import Pileup


def shard_ranges(cell_ptr: np.ndarray, nranks: int) -> List[Tuple[int, int]]:
    """contiguous [lo, hi) barcode ranges, one per rank, with (nearly) equal numbers of pileup entries"""
    ncells = int(cell_ptr.shape[0] - 1)
    total = int(cell_ptr[-1])
    cuts = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        c = int(np.searchsorted(cell_ptr, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), ncells))
    cuts.append(ncells)
    return [(cuts[i], cuts[i + 1]) for i in range(nranks)]


def shard_pileup(plp: Pileup, rank: int, nranks: int) -> Tuple[Pileup, Tuple[int, int]]:
    lo, hi = shard_ranges(plp.cell_ptr, nranks)[rank]
    return plp.subset_cells(np.arange(lo, hi)), (lo, hi)


def allgather_records(local: np.ndarray, counts: List[int], device=None) -> np.ndarray:
    """all-gather of per-barcode records (any fixed-size numpy dtype); `counts[r]` = records of rank r.
    Returns the concatenation in rank order on every rank.  One collective: ranks pad to the largest shard."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    assert len(counts) == world and local.shape[0] == counts[dist.get_rank()]
    item = local.dtype.itemsize
    width = max(counts) * item
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = np.zeros(width, dtype=np.uint8)
    buf[: local.shape[0] * item] = local.view(np.uint8).reshape(-1)
    send = torch.from_numpy(buf).to(device)
    recv = torch.empty(world * width, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send)
    out = recv.cpu().numpy()
    parts = [out[r * width: r * width + counts[r] * item].view(local.dtype) for r in range(world)]
    return np.concatenate(parts)


class RecordExchange:
    """The path's one collective, with its buffers kept between steps: all-gather of the per-barcode records of every
    rank's shard (`counts[r]` records of `itemsize` bytes on rank r; shards padded to the largest).  On "nccl" the
    records go pinned host -> device -> NCCL all-gather over NVLink -> pinned host."""

    def __init__(self, counts: List[int], itemsize: int, device=None):
        import torch
        import torch.distributed as dist
        self.dist, self.torch = dist, torch
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        assert len(counts) == self.world
        self.counts, self.item = list(counts), int(itemsize)
        self.width = max(max(counts), 1) * self.item
        nccl = dist.get_backend() == "nccl"
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
        self.h_send = torch.zeros(self.width, dtype=torch.uint8)
        self.h_recv = torch.zeros(self.world * self.width, dtype=torch.uint8)
        if nccl:
            self.h_send, self.h_recv = self.h_send.pin_memory(), self.h_recv.pin_memory()
        self.d_send = torch.zeros(self.width, dtype=torch.uint8, device=device)
        self.d_recv = torch.zeros(self.world * self.width, dtype=torch.uint8, device=device)
        self.on_device = nccl

    def send_view(self, dtype) -> np.ndarray:
        """this rank's slot of the pinned send buffer as `dtype` records: fill it in place (Demuxlet.best(out=...)) and
        call the exchange with that very array to skip one copy"""
        n = self.counts[self.rank] * self.item
        return self.h_send.numpy()[:n].view(dtype)

    def __call__(self, local: np.ndarray) -> np.ndarray:
        """returns a view of all ranks' records (rank order) valid until the next call"""
        n = self.counts[self.rank] * self.item
        assert local.dtype.itemsize == self.item and local.shape[0] == self.counts[self.rank]
        mine = self.h_send.numpy()[:n]
        if local.ctypes.data != mine.ctypes.data:
            mine[:] = local.view(np.uint8).reshape(-1)
        self.d_send.copy_(self.h_send, non_blocking=True)
        self.dist.all_gather_into_tensor(self.d_recv, self.d_send)
        self.h_recv.copy_(self.d_recv, non_blocking=True)
        if self.on_device:
            self.torch.cuda.current_stream().synchronize()
        out = self.h_recv.numpy()
        if len(set(self.counts)) == 1:
            return out[: self.world * n].view(local.dtype) if self.width == n else np.concatenate(
                [out[r * self.width: r * self.width + n] for r in range(self.world)]).view(local.dtype)
        return np.concatenate([out[r * self.width: r * self.width + self.counts[r] * self.item]
                               for r in range(self.world)]).view(local.dtype)
