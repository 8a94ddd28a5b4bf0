"""Rank -> work assignment of the inference path: volumes are independent
(eval_3d_sagittal_twostage.py:153) and so are the slices of one volume (:201), so the path shards
with NO data-path collective.  Pure host logic (no CUDA), shared by bench.py and the drivers."""
from __future__ import annotations

import os
from typing import Iterator, List, Sequence, Tuple


def world_from_env() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) as exported by torchrun; (0, 1, 0) for a plain python launch."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_round_robin(n_items: int, rank: int, world: int) -> List[int]:
    """Items r, r+world, ... (config 5: rank r takes volumes {v : v mod world == r})."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


def shard_contiguous(n_items: int, rank: int, world: int) -> range:
    """Contiguous chunk of a single volume's (orientation, z) slices (config 3); the first
    n_items % world ranks get one extra item so the chunks differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def batches(items: Sequence[int], batch: int) -> Iterator[Sequence[int]]:
    """Stage-major batching: consecutive groups of at most `batch` items."""
    if batch < 1:
        raise ValueError("batch must be >= 1")
    for i in range(0, len(items), batch):
        yield items[i:i + batch]


def volume_slices(n_sagittal: int, n_coronal: int) -> List[Tuple[str, int]]:
    """The (orientation, index) work list of one straightened volume: sagittal slices [:, :, z]
    (eval_3d_sagittal_twostage.py:201) then coronal slices [:, z, :] (RHLV_quantification_coronal.py:51-54)."""
    return [("sagittal", z) for z in range(n_sagittal)] + [("coronal", z) for z in range(n_coronal)]


def allreduce_mean_(tensors, world: int, scale_inplace) -> None:
    """Gradient exchange of the data-parallel training step: ``all_reduce(SUM)`` every tensor in place, then
    ``scale_inplace(t, 1 / world)``.  BatchNorm statistics stay per rank and the spectral-norm u / v buffers need no broadcast
    (they are a deterministic function of the replicated weights).  ``scale_inplace`` is the device kernel on the GPU path
    (hv_axpby) - the CPU tests pass a torch lambda; the collective itself is torch.distributed (NCCL over NVLink / gloo)."""
    if world <= 1:
        return
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t)
        scale_inplace(t, 1.0 / world)
