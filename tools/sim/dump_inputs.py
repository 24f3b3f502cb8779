"""Inputs of tools/sim/packet_walk_model.c from the CPU oracle (no GPU): the Sponza BVH as VkBvhNode[] after a quality build
(sponza_q.bin) and a fast build (sponza_f.bin), and the C2 ray batch (rays4k.bin).   python tools/sim/dump_inputs.py OUT_DIR"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as O
from radeonrays_sdk_b200 import workloads as W

out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/sim"
os.makedirs(out, exist_ok=True)
pos, idx, _ = W.load_mesh("sponza")
O.build_blas(pos, idx, restructure=True)[0].tofile(os.path.join(out, "sponza_q.bin"))
O.build_blas(pos, idx)[0].tofile(os.path.join(out, "sponza_f.bin"))
W.sponza_primary_rays(3840, 2160).tofile(os.path.join(out, "rays4k.bin"))
print("wrote sponza_q.bin, sponza_f.bin, rays4k.bin to", out)
