import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from oracle import binding as O
from radeonrays_sdk_b200 import api, workloads as W
from radeonrays_sdk_b200.host import Engine
e = Engine(0)
pos, idx, _ = W.load_mesh("sponza")
for flags in (1, 3):
    g = e.build_geometry(pos, idx, build_flags=flags)
    before = g.nodes().copy()
    moved = pos.copy(); moved[:, 1] -= np.float32(40.0)
    e.update_geometry(g, moved)
    got = g.nodes()
    want = O.refit(before, moved, idx)
    a = np.ascontiguousarray(got["aabb0_min_or_v0"]).view(np.uint32); b = np.ascontiguousarray(want["aabb0_min_or_v0"]).view(np.uint32)
    bad = np.nonzero((a != b).any(axis=1))[0]
    n = idx.shape[0]
    print("flags", flags, "bad", bad.size, "internal", (bad < n - 1).sum(), "leaves", (bad >= n - 1).sum(), bad[:10], bad[-5:])
    if bad.size:
        d = np.diff(bad); print("runs", (d != 1).sum() + 1, "first run len", np.argmax(d != 1) + 1 if (d != 1).any() else bad.size)
e.close()
