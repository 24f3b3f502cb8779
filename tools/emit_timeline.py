"""Phase timeline of k_emit_fit (library built with EXTRA=-DRR_EMIT_TIMELINE): every 64th CTA stamps %globaltimer at
start | leaves gathered | climbed | flushed | handed over.  Prints mean phase durations in us.
  make -C radeonrays_sdk_b200/csrc clean all EXTRA=-DRR_EMIT_TIMELINE && python tools/emit_timeline.py 4000x1000"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radeonrays_sdk_b200 import api
from radeonrays_sdk_b200.host import Engine, Geometry, _dev_bytes
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_build import heightfield_device

size = sys.argv[1] if len(sys.argv) > 1 else "4000x1000"
nx, nz = (int(v) for v in size.split("x"))
eng = Engine(0)
ctx, dev = eng.ctx, eng.device
pos, idx = heightfield_device(nx, nz, 0.0, dev)
n = idx.shape[0]
g = Geometry()
g.engine, g.triangle_count, g.vertex_count, g.vertex_stride = eng, n, pos.shape[0], 12
g.d_vertices, g.d_indices = pos.view(torch.uint8).reshape(-1), idx.view(torch.uint8).reshape(-1)
g.options = api.RRBuildOptions(api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, None)
g.p_vertices, g.p_indices = ctx.tensor_ptr(g.d_vertices), ctx.tensor_ptr(g.d_indices)
g.input = ctx.geometry_input(g.p_vertices, g.vertex_count, 12, g.p_indices, n)
g.req = ctx.geometry_requirements(g.input, g.options)
g.d_temp, g.d_nodes = _dev_bytes(g.req.temporary_build_buffer_size, dev), _dev_bytes(g.req.result_buffer_size, dev)
g.p_temp, g.p_nodes = ctx.tensor_ptr(g.d_temp), ctx.tensor_ptr(g.d_nodes)
for _ in range(3):
    eng.rebuild(g)
L = ctx.build_scratch_layout(n)
ctas = (n + 511) // 512
samples = (ctas + 63) // 64
t = g.d_temp[L.sort_tmp_values_offset: L.sort_tmp_values_offset + samples * 64].cpu().numpy().view(np.uint64).reshape(samples, 8)[:, [0, 1, 3, 4, 5]].astype(np.int64)
d = np.diff(t, axis=1) / 1e3
names = ["gather", "climb", "flush", "handover"]
print(f"{size}: {n} triangles, {ctas} CTAs, {samples} sampled; kernel span {(t[:, 4].max() - t[:, 0].min()) / 1e3:.1f} us")
for k, nm in enumerate(names):
    print(f"  {nm:9s} mean {d[:, k].mean():7.2f} us   p50 {np.median(d[:, k]):7.2f}   p95 {np.percentile(d[:, k], 95):7.2f}")
print(f"  CTA life  mean {(t[:, 4] - t[:, 0]).mean() / 1e3:7.2f} us")
eng.close()
