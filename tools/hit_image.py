"""The picture the reference's API tests write after tracing (test/test_vk/basic_test.h:720-741: stbi_write_jpg of
0xff000000 | u * 255 << 8 | v * 255 << 16 per hit, 0xff101010 per miss, flipped vertically), from this backend, as a PNG:

  python tools/hit_image.py [--mesh sponza|cornell_box] [--res 2048] [--two-level] [--out gpurun_out/isect.png] [--check]

--check also traces a 1/16-resolution frame with the CPU oracle and compares its picture byte for byte (test infrastructure)."""
import argparse, os, struct, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radeonrays_sdk_b200 import api, workloads as W


def write_png(path, rgba):
    h, w, _ = rgba.shape
    raw = b"".join(b"\x00" + rgba[y].tobytes() for y in range(h))
    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="sponza")
    ap.add_argument("--res", type=int, default=2048)
    ap.add_argument("--two-level", action="store_true", help="one BLAS per OBJ shape under a TLAS (BuildObj2Level, basic_test.h:752-1069)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "isect.png"))
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    from radeonrays_sdk_b200.host import Engine
    eng = Engine(0)
    pos, idx, first = W.load_mesh(a.mesh)
    rays_of = (lambda r: W.sponza_primary_rays(r, r)) if a.mesh == "sponza" else W.cornell_primary_rays
    if a.two_level:
        geoms = [eng.build_geometry(pos, idx[first[s]:first[s + 1]]) for s in range(len(first) - 1)]
        xf = np.zeros((len(geoms), 3, 4), np.float32)
        xf[:, 0, 0] = xf[:, 1, 1] = xf[:, 2, 2] = 1
        target = eng.build_scene(geoms, list(range(len(geoms))), xf)
    else:
        target = eng.build_geometry(pos, idx, build_flags=0)
    hits = eng.intersect(target, rays_of(a.res))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    write_png(a.out, W.hits_to_image(hits, a.res, a.res))
    print("wrote", a.out, "hit fraction %.4f" % float((hits["inst_id"] != W.INVALID).mean()))
    if a.check and not a.two_level:
        from oracle import binding as O
        small = max(16, a.res // 16)
        r = rays_of(small)
        want = W.hits_to_image(O.trace(target.nodes(), r), small, small)
        got = W.hits_to_image(eng.intersect(target, r), small, small)
        diff = int(np.count_nonzero((want != got).any(axis=2)))
        print("oracle picture at %dx%d: %d pixels differ" % (small, small, diff))
        assert diff <= 2          # (t, prim) ties at the ulp level, tests/helpers.py assert_closest_hits_equal
    eng.close()
