"""The N > 1 path on CPU: world_size-2 `gloo` processes.  The partition + single-reduce + resolve logic is the product's
(`spp_slice`, `row_bands`, `sharded_frame` in simd-raytracer_b200/__init__.py - the same functions bench.py drives over
NCCL); the per-rank renderer is the oracle here because the product has no CPU path."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest

from .conftest import REPO, resized, scene_bytes


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    import importlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rt = importlib.import_module("simd-raytracer_b200")
    from tests.helpers import oracle
    data = resized(scene_bytes("hw15_scene2"), 96, 64)
    o = oracle.Oracle(data)
    spp_total = 5
    kw = dict(gi_rays=1, max_ray_depth=3)

    def render_slice(first, n):
        if n == 0:
            return torch.zeros((o.height, o.width, 3))
        img, _ = o.render(oracle.default_params(spp=n, sample_offset=first, spp_total=spp_total, raw_sum=1, **kw), threads=2)
        return torch.from_numpy(img)

    frame = rt.sharded_frame(render_slice, spp_total, rank, world, lambda fb: dist.reduce(fb, dst=0, op=dist.ReduceOp.SUM),
                             lambda fb: fb / np.float32(spp_total))
    # tile sharding: interleaved row bands, gathered on rank 0
    band = torch.zeros((o.height, o.width, 3))
    for y0, y1 in rt.row_bands(o.height, 16, rank, world):
        o.render(oracle.default_params(spp=2, **kw), rect=(0, y0, o.width, y1), out=band.numpy(), threads=2)
    dist.reduce(band, dst=0, op=dist.ReduceOp.SUM)      # bands are disjoint and zero elsewhere: sum == gather
    if rank == 0:
        np.save(os.path.join(out_dir, "spp.npy"), frame.numpy())
        np.save(os.path.join(out_dir, "bands.npy"), band.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(oracle_mod, rt, tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    o = oracle_mod.Oracle(resized(scene_bytes("hw15_scene2"), 96, 64))
    full, _ = o.render(oracle_mod.default_params(spp=5, gi_rays=1, max_ray_depth=3))
    np.testing.assert_allclose(np.load(tmp_path / "spp.npy"), full, rtol=1e-6, atol=2e-6)      # slice sums re-associate the sample sum
    full2, _ = o.render(oracle_mod.default_params(spp=2, gi_rays=1, max_ray_depth=3))
    assert np.array_equal(np.load(tmp_path / "bands.npy"), full2)                           # tile sharding is exact
