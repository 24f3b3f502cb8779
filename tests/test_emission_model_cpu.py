"""CPU test (-m "not gpu") of the claim the CUDA emission kernels rest on (rr_build.cu group_merge):

  the Karras hierarchy the reference emits with FindSpan / FindSplit (lbvh_emit_hierarchy_mesh.comp:105-217) is the
  Cartesian tree of the deltas between neighbouring sorted leaves -- the node split right of leaf s spans the leaves between
  the nearest SMALLER delta on either side; it is stored at the right end of its range if delta(R, R+1) > delta(L-1, L) (left
  child) else at the left end; its children are the split / split + 1 (leaves when the child range is a single leaf); its
  parent is the node split at the larger of the two bounding deltas; its child boxes are the unions of the leaf boxes of the
  two child ranges.

The model below is that statement in numpy, checked against the oracle's restatement of the shader on random, duplicate-heavy
and chain-shaped inputs -- topology and boxes bit for bit.  The GPU tests then check the kernels against the same oracle.
"""
import numpy as np
import pytest

from oracle import binding as O
from radeonrays_sdk_b200 import workloads as W

INVALID = 0xFFFFFFFF


def _deltas(codes):
    """delta(a, a+1) as the kernels compute it: clz of the xor of the codes, index tie-break for equal codes; index -1 and
    n-1 (outside the array) are 0, smaller than every real delta."""
    n = codes.shape[0]
    d = np.zeros(n + 1, np.int64)                      # d[a + 1] = delta(a, a + 1), a = -1 .. n-1
    for a in range(n - 1):
        x = int(codes[a]) ^ int(codes[a + 1])
        d[a + 1] = (32 - x.bit_length()) if x else 32 + (32 - (a ^ (a + 1)).bit_length())
    return d


def _ordered(x):
    """Order-preserving uint key of a float array, -0 below +0 (common.h:68-90): min / max through it behave like the
    kernels' fminf / fmaxf and the oracle's, which numpy's min / max do not for signed zeros."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u >> 31 == 1, 0xFFFFFFFF - u, u | 0x80000000)


def _fmin(a):
    k = _ordered(a)
    return np.take_along_axis(a, k.argmin(0)[None], 0)[0]


def _fmax(a):
    k = _ordered(a)
    return np.take_along_axis(a, k.argmax(0)[None], 0)[0]


def cartesian_emission(codes, refs, leaf_lo, leaf_hi):
    """Closed-form emission: returns the node array fields the oracle produces (children, parent, child boxes)."""
    n = codes.shape[0]
    leaf0 = n - 1
    d = _deltas(codes)
    D = lambda a: d[a + 1]
    child0 = np.full(2 * n - 1, INVALID, np.uint32); child1 = np.zeros(2 * n - 1, np.uint32); parent = np.full(2 * n - 1, INVALID, np.uint32)
    box = np.zeros((2 * n - 1, 4, 3), np.float32)      # aabb0_min, aabb0_max, aabb1_min, aabb1_max
    child1[leaf0:] = refs
    # prefix structure for range boxes (plain loops: the sizes are small)
    index_of_split = {}
    spans = {}
    for s in range(n - 1):
        L = s
        while L - 1 >= -1 and D(L - 1) >= D(s):        # nearest smaller delta on the left
            L -= 1
        R = s + 1
        while D(R) >= D(s):                            # ... on the right (D(n-1) = 0 stops it)
            R += 1
        is_root = L == 0 and R == n - 1
        left = D(R) > D(L - 1)
        idx = 0 if is_root else (R if left else L)
        index_of_split[s] = idx
        spans[s] = (L, R, is_root, left)
    for s, (L, R, is_root, left) in spans.items():
        idx = index_of_split[s]
        c0 = leaf0 + s if L == s else s                # Karras: left child = split, right child = split + 1
        c1 = leaf0 + s + 1 if R == s + 1 else s + 1
        child0[idx], child1[idx] = c0, c1
        box[idx, 0] = _fmin(leaf_lo[L:s + 1]); box[idx, 1] = _fmax(leaf_hi[L:s + 1])
        box[idx, 2] = _fmin(leaf_lo[s + 1:R + 1]); box[idx, 3] = _fmax(leaf_hi[s + 1:R + 1])
        if not is_root:
            q = R if left else L - 1                   # the parent is split at the larger bounding delta
            parent[idx] = index_of_split[q]
    for j in range(n):                                 # leaves: parent = the node split at the larger neighbouring delta
        if n > 1:
            parent[leaf0 + j] = index_of_split[j if D(j) > D(j - 1) else j - 1]
    return child0, child1, parent, box


def _check(pos, idx):
    nodes, sc, sr = O.build_blas(pos, idx)
    n = idx.shape[0]
    tri = pos[idx[sr]]                                 # [n, 3, 3] in sorted order
    tri_t = np.ascontiguousarray(tri.transpose(1, 0, 2))  # [3, n, 3]: per-leaf min / max over the three vertices
    c0, c1, par, box = cartesian_emission(sc, sr, _fmin(tri_t), _fmax(tri_t))
    assert np.array_equal(nodes["child0"], c0)
    assert np.array_equal(nodes["child1"], c1)
    assert np.array_equal(nodes["parent"], par)
    internal = slice(0, n - 1)
    for k, f in enumerate(("aabb0_min_or_v0", "aabb0_max_or_v1", "aabb1_min_or_v2", "aabb1_max_or_v3")):
        assert np.array_equal(np.ascontiguousarray(nodes[f][internal]).view(np.uint32), np.ascontiguousarray(box[internal, k]).view(np.uint32)), f


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 513, 1500])
def test_random_meshes(n):
    rng = np.random.default_rng(n)
    pos = rng.random((3 * n, 3), dtype=np.float32)
    idx = rng.permutation(3 * n).astype(np.uint32).reshape(n, 3)
    _check(pos, idx)


def test_equal_codes_use_the_index_tie_break():
    n = 700
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    _check(pos, np.tile(np.array([[0, 1, 2]], np.uint32), (n, 1)))


def test_chain_along_the_octree_diagonal():
    pos, idx = [], []
    for lvl in range(1, 11):
        for rep in range(3 * lvl):
            c = np.float32(2.0 ** -lvl) * (np.float32(1.0) + np.float32(0.3) * np.float32(rep) / np.float32(3 * lvl))
            b = len(pos)
            pos.extend([np.array([c, c, c], np.float32), np.array([c, c, c], np.float32) * np.float32(1.0001), np.array([c, c * np.float32(1.0002), c], np.float32)])
            idx.append((b, b + 1, b + 2))
    pos.append(np.zeros(3, np.float32)); pos.append(np.ones(3, np.float32))
    idx.append((len(pos) - 2, len(pos) - 1, len(pos) - 2))
    _check(np.asarray(pos, np.float32), np.asarray(idx, np.uint32))


def test_cornell_box(cornell):
    pos, idx, _ = cornell
    _check(pos, idx)
