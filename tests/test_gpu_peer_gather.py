"""-m gpu: the fused trace + gather path of the multi-GPU split (sharding.PeerHitBuffer, rrCudaExportDeviceMemory /
rrCudaImportDeviceMemory) with two PROCESSES.  Both ranks use cuda:0 when the box has one GPU (the IPC mapping is then a same-device
mapping; the driver's 2/4/8-GPU bench runs the same code over NVLink), rendezvous over gloo on 127.0.0.1."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radeonrays_sdk_b200 import api, sharding, workloads as W
    from radeonrays_sdk_b200.host import Engine
    device = rank % torch.cuda.device_count()
    eng = Engine(device)
    ctx = eng.ctx
    pos, idx, _ = W.load_mesh("sponza")
    # rank 0 builds the BLAS, everybody receives its bytes (indices, not pointers)
    g = eng.build_geometry(pos, idx, build_flags=0)
    if rank != 0:
        g.d_nodes.zero_()
    host = g.d_nodes.cpu()
    sharding.broadcast_bytes(host, src=0)
    g.d_nodes.copy_(host)
    rays = W.sponza_primary_rays(500, 301)[:150_437]          # ragged on purpose
    n = rays.shape[0]
    b, e = sharding.shard_range(n, rank, world)
    for query, output, item in ((api.RR_INTERSECT_QUERY_CLOSEST, api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, 16),
                                (api.RR_INTERSECT_QUERY_ANY, api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID, 4)):
        peer = sharding.PeerHitBuffer(ctx, item * n, root=0)
        rb = eng.make_ray_buffers(e - b, output)
        rb.d_rays[: 32 * (e - b)].copy_(torch.from_numpy(rays[b:e].view(np.uint8).reshape(-1)).to(eng.device))
        p_hits = peer.ptr(item * b)
        ctx.run(lambda s: ctx.cmd_intersect(g.p_nodes, query, rb.p_rays, e - b, None, output, p_hits, rb.p_scratch, s))
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            np.save(os.path.join(tmp, f"hits{item}.npy"), peer.read(np.uint8))
        dist.barrier()
        peer.close()
    if rank == 0:
        np.save(os.path.join(tmp, "nodes.npy"), g.nodes())
    eng.close()
    dist.destroy_process_group()


def test_two_processes_store_hits_into_rank0_buffer(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    port = 29700 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import binding as O
    from radeonrays_sdk_b200 import workloads as W
    from helpers import assert_hits_equal
    pos, idx, _ = W.load_mesh("sponza")
    nodes = np.load(os.path.join(str(tmp_path), "nodes.npy"))
    rays = W.sponza_primary_rays(500, 301)[:150_437]
    got = np.load(os.path.join(str(tmp_path), "hits16.npy")).view(W.HIT_DTYPE)
    assert_hits_equal(got, O.trace(nodes, rays, init=np.zeros(rays.shape[0], W.HIT_DTYPE)), what="peer gather closest", mesh=(pos, idx), rays=rays)
    ids = np.load(os.path.join(str(tmp_path), "hits4.npy")).view(np.uint32)
    assert np.array_equal(ids, O.trace(nodes, rays, O.QUERY_ANY, O.OUTPUT_INSTANCE_ID))


def _scene_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radeonrays_sdk_b200 import api, sharding, workloads as W
    from radeonrays_sdk_b200.host import Engine
    eng = Engine(rank % torch.cuda.device_count())
    pos, idx, first = W.load_mesh("sponza")
    shapes = [73, 288, 87, 39, 8]                                   # five OBJ shapes of Sponza as five meshes
    meshes = [(pos, idx[first[s]:first[s + 1]]) for s in shapes]
    inst = [0, 1, 2, 3, 4, 2, 0]
    xf = W.grid_instances(2, 40.0, 13.0)[: len(inst)]
    geoms, scene = sharding.build_scene_sharded(eng, meshes, inst, xf)
    built_here = [i for i in range(len(meshes)) if i % world == rank]
    rays = W.random_rays(60_000, (-100, -10, -60), (130, 110, 100), seed=17)
    hits = eng.intersect(scene, rays)
    np.save(os.path.join(tmp, f"scene_hits{rank}.npy"), hits)
    np.save(os.path.join(tmp, f"built{rank}.npy"), np.array(built_here))
    if rank == 0:
        for i, g in enumerate(geoms):
            np.save(os.path.join(tmp, f"blas{i}.npy"), g.d_nodes[: 64 * (2 * g.triangle_count - 1)].cpu().numpy())
    eng.close()
    dist.destroy_process_group()


def test_multi_mesh_scene_blas_builds_are_distributed(tmp_path):
    """SURVEY.md section 8e, multi-mesh scenes: BLAS i is built on rank i mod N, its bytes are sent to everybody, every rank builds
    the TLAS over its local copies and traces -- both ranks must return the oracle's two-level hits."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    port = 29900 + os.getpid() % 2000
    mp.spawn(_scene_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import binding as O
    from radeonrays_sdk_b200 import workloads as W
    from helpers import assert_hits_equal, assert_nodes_equal
    pos, idx, first = W.load_mesh("sponza")
    shapes = [73, 288, 87, 39, 8]
    inst = [0, 1, 2, 3, 4, 2, 0]
    xf = W.grid_instances(2, 40.0, 13.0)[: len(inst)]
    blas = [O.build_blas(pos, idx[first[s]:first[s + 1]], restructure=True)[0] for s in shapes]
    for i, want in enumerate(blas):     # rank 0 holds every BLAS, also those rank 1 built
        assert_nodes_equal(np.load(os.path.join(str(tmp_path), f"blas{i}.npy")).view(W.NODE_DTYPE), want, what=f"blas {i}")
    assert list(np.load(os.path.join(str(tmp_path), "built0.npy"))) == [0, 2, 4] and list(np.load(os.path.join(str(tmp_path), "built1.npy"))) == [1, 3]
    tlas, oxf = O.build_tlas(blas, inst, xf)
    rays = W.random_rays(60_000, (-100, -10, -60), (130, 110, 100), seed=17)
    want = O.trace_2l(tlas, oxf, blas, inst, rays, init=np.zeros(rays.shape[0], W.HIT_DTYPE))
    assert (want["inst_id"] != O.INVALID).sum() > 1000
    for r in range(2):
        assert_hits_equal(np.load(os.path.join(str(tmp_path), f"scene_hits{r}.npy")), want, what=f"rank {r}")
