"""ctypes binding of oracle/librr_oracle.so (the CPU restatement of the reference shaders).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs -- never by the product package radeonrays_sdk_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NODE_DTYPE = np.dtype([("aabb0_min_or_v0", "<f4", 3), ("child0", "<u4"),
                       ("aabb0_max_or_v1", "<f4", 3), ("child1", "<u4"),
                       ("aabb1_min_or_v2", "<f4", 3), ("parent", "<u4"),
                       ("aabb1_max_or_v3", "<f4", 3), ("update", "<u4")])
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("min_t", "<f4"), ("direction", "<f4", 3), ("max_t", "<f4")])
HIT_DTYPE = np.dtype([("uv", "<f4", 2), ("inst_id", "<u4"), ("prim_id", "<u4")])
STATS_DTYPE = np.dtype([("nodes_visited", "<u4"), ("triangles_tested", "<u4"), ("max_stack", "<u4"), ("t", "<f4")])
assert NODE_DTYPE.itemsize == 64 and RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16

INVALID = 0xFFFFFFFF
QUERY_CLOSEST, QUERY_ANY = 0, 1
OUTPUT_FULL_HIT, OUTPUT_INSTANCE_ID = 0, 1
TIE_LOWEST_ID, TIE_FIRST_FOUND = 0, 1


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "librr_oracle.so")
    src = os.path.join(_HERE, "rr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "librr_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.rro_sah.restype = C.c_float
        _LIB.rro_depth.restype = C.c_uint32
        _LIB.rro_check_consistency.restype = C.c_int
        _LIB.rro_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def scene_aabb(verts, idx):
    verts, idx = _f32(verts), _u32(idx)
    lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib().rro_scene_aabb(_p(verts), C.c_uint32(verts.shape[1]), _p(idx), C.c_uint32(idx.shape[0]), _p(lo), _p(hi))
    return lo, hi


def morton_codes(verts, idx, lo, hi):
    verts, idx = _f32(verts), _u32(idx)
    codes = np.zeros(idx.shape[0], np.uint32)
    lib().rro_morton_codes(_p(verts), C.c_uint32(verts.shape[1]), _p(idx), C.c_uint32(idx.shape[0]),
                           _p(_f32(lo)), _p(_f32(hi)), _p(codes))
    return codes


def sort_pairs(keys, vals):
    keys, vals = _u32(keys), _u32(vals)
    ok, ov = np.zeros_like(keys), np.zeros_like(vals)
    lib().rro_sort_pairs(_p(keys), _p(vals), C.c_uint32(keys.shape[0]), _p(ok), _p(ov))
    return ok, ov


def build_blas(verts, idx, restructure=False):
    """-> (nodes[2N-1], sorted_codes[N], sorted_refs[N])"""
    verts, idx = _f32(verts), _u32(idx)
    n = idx.shape[0]
    nodes = np.zeros(2 * n - 1, NODE_DTYPE)
    sc, sr = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    lib().rro_build_blas(_p(verts), C.c_uint32(verts.shape[1]), _p(idx), C.c_uint32(n), C.c_int(int(restructure)),
                         _p(nodes), _p(sc), _p(sr))
    return nodes, sc, sr


def build_blas63(verts, idx, restructure=False):
    """63-bit Morton extension -> (nodes[2N-1], sorted_codes u64[N], sorted_refs[N])"""
    verts, idx = _f32(verts), _u32(idx)
    n = idx.shape[0]
    nodes = np.zeros(2 * n - 1, NODE_DTYPE)
    sc, sr = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    lib().rro_build_blas63(_p(verts), C.c_uint32(verts.shape[1]), _p(idx), C.c_uint32(n), C.c_int(int(restructure)),
                           _p(nodes), _p(sc), _p(sr))
    return nodes, sc, sr


def refit(nodes, verts, idx):
    verts, idx = _f32(verts), _u32(idx)
    nodes = nodes.copy()
    n = idx.shape[0]
    lib().rro_fit_mesh(_p(nodes), C.c_uint32(n), _p(verts), C.c_uint32(verts.shape[1]), _p(idx))
    return nodes


def restructure(nodes):
    nodes = nodes.copy()
    lib().rro_restructure(_p(nodes), C.c_uint32((nodes.shape[0] + 1) // 2))
    return nodes


def build_tlas(blas_list, instance_blas, transforms, reference_corner_quirk=False):
    """blas_list: list of node arrays; instance_blas[i] = index into blas_list; transforms (n,3,4).
    -> (tlas_nodes[2n-1], out_transforms (2n,12))"""
    n = len(instance_blas)
    transforms = _f32(transforms).reshape(n, 12)
    roots = (C.c_void_p * n)(*[blas_list[b].ctypes.data for b in instance_blas])
    tris = _u32([(blas_list[b].shape[0] + 1) // 2 for b in instance_blas])
    nodes = np.zeros(2 * n - 1, NODE_DTYPE)
    out = np.zeros((2 * n, 12), np.float32)
    lib().rro_build_tlas(roots, _p(tris), _p(transforms), C.c_uint32(n), C.c_int(int(reference_corner_quirk)),
                         _p(nodes), _p(out))
    return nodes, out


def check_consistency(nodes):
    return bool(lib().rro_check_consistency(_p(nodes), C.c_uint32((nodes.shape[0] + 1) // 2)))


def sah(nodes):
    return float(lib().rro_sah(_p(nodes), C.c_uint32((nodes.shape[0] + 1) // 2)))


def depth(nodes):
    return int(lib().rro_depth(_p(nodes), C.c_uint32((nodes.shape[0] + 1) // 2)))


def _hits_buffer(count, output, init):
    if output == OUTPUT_FULL_HIT:
        h = np.zeros(count, HIT_DTYPE) if init is None else np.ascontiguousarray(init, dtype=HIT_DTYPE).copy()
    else:
        h = np.zeros(count, np.uint32) if init is None else _u32(init).copy()
    return h


def trace(nodes, rays, query=QUERY_CLOSEST, output=OUTPUT_FULL_HIT, tie=TIE_LOWEST_ID, init=None, want_stats=False):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    hits = _hits_buffer(rays.shape[0], output, init)
    stats = np.zeros(rays.shape[0], STATS_DTYPE) if want_stats else None
    lib().rro_trace(_p(nodes), _p(rays), C.c_uint32(rays.shape[0]), C.c_int(query), C.c_int(output), C.c_int(tie),
                    _p(hits), _p(stats))
    return (hits, stats) if want_stats else hits


def trace_2l(tlas, out_transforms, blas_list, instance_blas, rays, query=QUERY_CLOSEST, output=OUTPUT_FULL_HIT,
             tie=TIE_LOWEST_ID, init=None, want_stats=False):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    n = len(instance_blas)
    roots = (C.c_void_p * n)(*[blas_list[b].ctypes.data for b in instance_blas])
    hits = _hits_buffer(rays.shape[0], output, init)
    stats = np.zeros(rays.shape[0], STATS_DTYPE) if want_stats else None
    lib().rro_trace_2l(_p(tlas), _p(_f32(out_transforms)), roots, _p(rays), C.c_uint32(rays.shape[0]), C.c_int(query),
                       C.c_int(output), C.c_int(tie), _p(hits), _p(stats))
    return (hits, stats) if want_stats else hits


def brute_force(verts, idx, rays):
    verts, idx = _f32(verts), _u32(idx)
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    hits = np.zeros(rays.shape[0], HIT_DTYPE)
    t = np.zeros(rays.shape[0], np.float32)
    lib().rro_brute_force(_p(verts), C.c_uint32(verts.shape[1]), _p(idx), C.c_uint32(idx.shape[0]), _p(rays),
                          C.c_uint32(rays.shape[0]), _p(hits), _p(t))
    return hits, t


def num_threads():
    return int(lib().rro_num_threads())


BRUTE_DTYPE = np.dtype([("t", "<f4"), ("prim_id", "<u4"), ("count", "<u4"), ("pad", "<u4")])


def ref_bvh_analyzer_trace(nodes, rays, repeats=1, workdir=None, want_hits=False, threads=None, want_brute=False,
                           stock_schedule=False):
    """Run the compiled REFERENCE CPU tracer (oracle/_ref/bvh_analyzer_trace) on a node dump. Returns its JSON.
    `threads` sets OMP_NUM_THREADS for the child (torchrun exports OMP_NUM_THREADS=1 to its ranks).
    want_hits: res["hits"] = the bvh::Hit array BvhIntersect<2> returned (bvh_analyzer/bvh.h:226-319).
    want_brute: res["brute"] = per ray the (t, prim) minimum and the count of accepted triangles from the reference's
    Triangle::Intersect (bvh_analyzer/triangle.h:34-70) run over every triangle of the dump (O(rays x triangles))."""
    import json, tempfile
    exe = os.path.join(_HERE, "_ref", "bvh_analyzer_trace")
    if not os.path.exists(exe):
        return None
    d = workdir or tempfile.mkdtemp()
    n = (nodes.shape[0] + 1) // 2
    nodes.tofile(os.path.join(d, "bvh.bin"))
    np.ascontiguousarray(rays, dtype=RAY_DTYPE).tofile(os.path.join(d, "rays.bin"))
    cmd = [exe, os.path.join(d, "bvh.bin"), str(n - 1), str(n), os.path.join(d, "rays.bin"), str(rays.shape[0]), str(repeats)]
    if want_hits or want_brute:
        cmd.append(os.path.join(d, "hits.bin"))
    if want_brute:
        cmd.append(os.path.join(d, "brute.bin"))
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(int(threads))
    if stock_schedule:
        env["RR_REF_STOCK_SCHEDULE"] = "1"
    out = subprocess.run(cmd, capture_output=True, text=True, env=env)
    res = json.loads(out.stdout.strip().splitlines()[-1])
    if want_hits or want_brute:
        res["hits"] = np.fromfile(os.path.join(d, "hits.bin"), dtype=HIT_DTYPE)
    if want_brute:
        res["brute"] = np.fromfile(os.path.join(d, "brute.bin"), dtype=BRUTE_DTYPE)
    return res


def write_bvh_analyzer_config(directory, nodes, rays, width, height):
    """The on-disk input of the STOCK bvh_analyzer binary: VkBvhNode dump + RRRay dump + the 7-line text config
    (bvh_analyzer/config.h:46-63: bvh path, type string, internal-node count, triangle count, rays path, width, height;
    the reference writes the two dumps under DUMP_BVH, test/test_vk/basic_test.h:519-555).  Returns the config path."""
    n = (nodes.shape[0] + 1) // 2
    assert rays.shape[0] == width * height
    bvh_path, ray_path, cfg = (os.path.join(directory, f) for f in ("bvh.bin", "rays.bin", "config.txt"))
    np.ascontiguousarray(nodes, dtype=NODE_DTYPE).tofile(bvh_path)
    np.ascontiguousarray(rays, dtype=RAY_DTYPE).tofile(ray_path)
    with open(cfg, "w") as f:
        f.write("\n".join([bvh_path, "vkbvh2", str(n - 1), str(n), ray_path, str(width), str(height)]) + "\n")
    return cfg


def ref_bvh_analyzer_stock(nodes, rays, width, height, threads=None):
    """Run the UNMODIFIED reference binary (oracle/_ref/bvh_analyzer = bvh_analyzer/main.cpp) end to end on the 7-line config:
    IsValid, SAH, the `#pragma omp parallel for` trace with its per-ray omp critical, two JPEG writes (bvh.h:78-121).
    Returns {is_valid, sah, avg_primary_node_tests, avg_primary_aabb_tests, avg_primary_triangle_tests, wall_s}, or None
    when the binary is absent."""
    import tempfile, time
    exe = os.path.join(_HERE, "_ref", "bvh_analyzer")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as d:
        cfg = write_bvh_analyzer_config(d, nodes, rays, width, height)
        env = dict(os.environ)
        if threads:
            env["OMP_NUM_THREADS"] = str(int(threads))
        t0 = time.time()
        out = subprocess.run([exe, cfg], capture_output=True, text=True, env=env, cwd=d)   # the JPEGs land in d
        wall = time.time() - t0
        res = {"wall_s": wall, "returncode": out.returncode, "wrote_jpegs": os.path.exists(os.path.join(d, "isect_result.jpg"))}
    for line in out.stdout.splitlines():
        k, _, v = line.partition(": ")
        if k in ("is_valid", "sah", "avg_primary_node_tests", "avg_primary_aabb_tests", "avg_primary_triangle_tests"):
            res[k] = bool(int(v)) if k == "is_valid" else float(v)
    return res
