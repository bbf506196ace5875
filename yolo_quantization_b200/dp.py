"""Data-parallel plumbing for the batch-sharded inference path.

The path is embarrassingly data-parallel: one process per GPU, a full replica of the ~8.7 MB of quantized
weights per GPU, images sharded by rank.  The only communication is ONE broadcast of the ``.weights`` byte
stream from rank 0 at load time (NCCL on GPUs, gloo in the CPU tests); there is no per-step collective.
``torch.distributed`` is plumbing here, not the product.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n_items`` images owned by ``rank``; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_weights_file(path: str, rank: int, world: int, device=None, dist=None) -> int:
    """Rank 0 holds ``path``; every other rank receives the bytes and writes them to its own ``path``.

    Returns the number of bytes.  ``dist`` is ``torch.distributed`` (already initialised) or None for world 1.
    """
    if world == 1 or dist is None:
        return os.path.getsize(path)
    import torch
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([os.path.getsize(path) if rank == 0 else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, 0)
    if rank == 0:
        blob = torch.from_numpy(np.fromfile(path, dtype=np.uint8)).to(dev)
    else:
        blob = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(blob, 0)
    if rank != 0:
        blob.cpu().numpy().tofile(path)
    return int(n.item())
