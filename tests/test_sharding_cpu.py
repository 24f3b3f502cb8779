"""CPU tests (-m "not gpu"): the N>1 path (ray sharding, BLAS broadcast, hit gather) with gloo, world_size 2."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radeonrays_sdk_b200 import sharding


def test_shard_ranges_cover_and_align():
    for count in (1, 31, 32, 33, 1000, 8294400, 16777216):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
                assert e0 == b1 and (b1 % 64 == 0 or b1 == count)
            assert sum(e - b for b, e in spans) == count


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as O
    from radeonrays_sdk_b200 import workloads as W
    pos, idx, _ = W.load_mesh("cornell_box")
    # rank 0 builds, everybody receives the same bytes (a BLAS is position independent)
    nodes = O.build_blas(pos, idx)[0] if rank == 0 else np.zeros(2 * idx.shape[0] - 1, O.NODE_DTYPE)
    t = torch.from_numpy(nodes.view(np.uint8).reshape(-1))
    sharding.broadcast_bytes(t, src=0)
    nodes = t.numpy().view(O.NODE_DTYPE)
    rays = W.cornell_primary_rays(50)[:2477]           # ragged count on purpose

    def trace_fn(slice_):                                # the CPU oracle stands in for the device tracer here
        return torch.from_numpy(O.trace(nodes, slice_).view(np.uint8).reshape(-1).copy())

    allhits = sharding.trace_sharded(trace_fn, rays).numpy().view(O.HIT_DTYPE)
    np.save(os.path.join(tmp, f"hits{rank}.npy"), allhits)
    dist.destroy_process_group()


def test_two_rank_trace_equals_single_rank(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from oracle import binding as O
    from radeonrays_sdk_b200 import workloads as W
    pos, idx, _ = W.load_mesh("cornell_box")
    want = O.trace(O.build_blas(pos, idx)[0], W.cornell_primary_rays(50)[:2477])
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"hits{r}.npy"))
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
