"""Host-side sharding logic of the multi-GPU inference path, exercised with a world-size-2 gloo group on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from healthivert_gan_b200 import sharding


def test_partitions_cover_everything_once():
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            rr = [sharding.shard_round_robin(n, r, world) for r in range(world)]
            cc = [list(sharding.shard_contiguous(n, r, world)) for r in range(world)]
            assert sorted(sum(rr, [])) == list(range(n))
            assert sum(cc, []) == list(range(n))
            assert max(len(c) for c in cc) - min(len(c) for c in cc) <= 1
    with pytest.raises(ValueError):
        sharding.shard_round_robin(4, 2, 2)
    assert [list(b) for b in sharding.batches(list(range(5)), 2)] == [[0, 1], [2, 3], [4]]
    work = sharding.volume_slices(3, 2)
    assert work == [("sagittal", 0), ("sagittal", 1), ("sagittal", 2), ("coronal", 0), ("coronal", 1)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_volumes, n_slices, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.world_from_env() == (rank, world, rank)
        # config 5: volumes round-robin; each "result" is a deterministic function of (volume, slice)
        mine = sharding.shard_round_robin(n_volumes, rank, world)
        res = torch.zeros(n_volumes, n_slices, dtype=torch.int64)
        for v in mine:
            for b in sharding.batches(list(range(n_slices)), 16):
                for z in b:
                    res[v, z] = v * 1000 + z + 1
        # results are written per volume with no data-path collective; this gather only checks them
        dist.all_reduce(res)
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            torch.save({"res": res, "tmax": float(t)}, out)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    n_volumes, n_slices = 5, 40
    mp.spawn(_worker, args=(2, _free_port(), n_volumes, n_slices, out), nprocs=2, join=True)
    got = torch.load(out)
    want = torch.tensor([[v * 1000 + z + 1 for z in range(n_slices)] for v in range(n_volumes)])
    assert torch.equal(got["res"], want)       # every (volume, slice) produced exactly once
    assert got["tmax"] == 2.0


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        grads = [torch.randn(7, 5, generator=g), torch.randn(11, generator=g)]
        sharding.allreduce_mean_(grads, world, lambda t, f: t.mul_(f))
        if rank == 0:
            torch.save(grads, out)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_mean_world_size_2(tmp_path):
    """The training step's only exchange: mean of the per-rank gradients, checked against a local mean."""
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    want = None
    for rank in range(2):
        g = torch.Generator().manual_seed(100 + rank)
        gs = [torch.randn(7, 5, generator=g), torch.randn(11, generator=g)]
        want = gs if want is None else [a + b for a, b in zip(want, gs)]
    for a, b in zip(got, want):
        assert torch.allclose(a, b / 2, atol=1e-7)
