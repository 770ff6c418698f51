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
