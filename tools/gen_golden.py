"""Regenerate tests/golden/oracle_digests.json (digests of the oracle's outputs on the committed mesh fixtures)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import binding as O
from radeonrays_sdk_b200 import workloads as W
from test_oracle_cpu import digest

out = {}
for name, rays in (("cornell_box", W.cornell_primary_rays(128)), ("sponza", W.sponza_primary_rays(160, 90))):
    pos, idx, _ = W.load_mesh(name)
    nodes, sc, sr = O.build_blas(pos, idx)
    out[name] = {"sorted_codes": digest(sc), "sorted_refs": digest(sr),
                 "nodes": digest(np.stack([nodes[f].view(np.uint32).reshape(nodes.shape[0], -1) for f in ("child0", "child1", "parent")], 1)),
                 "boxes": digest(np.concatenate([nodes[f] for f in ("aabb0_min_or_v0", "aabb0_max_or_v1", "aabb1_min_or_v2", "aabb1_max_or_v3")], 1)),
                 "closest_hits": digest(O.trace(nodes, rays)), "any_ids": digest(O.trace(nodes, rays, O.QUERY_ANY, O.OUTPUT_INSTANCE_ID)),
                 "treelet_nodes": digest(O.restructure(nodes)["child0"])}
json.dump(out, open(os.path.join(ROOT, "tests/golden/oracle_digests.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
