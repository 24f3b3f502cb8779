"""Build tests/golden/{cornell_box,sponza}.npz from the reference's OBJ files.

Runs only in the build container (needs /root/reference).  Compiles tools/gen_mesh_fixture.cpp
against the reference's own loader (test/test_vk/mesh_data.h + tiny_obj_loader), runs it on
resources/*.obj, and stores positions/indices/shape offsets as compressed .npz.  The meshes are
CC-BY 3.0 (see tests/golden/LICENSES.txt); no reference source code is copied.
"""
import os, subprocess, sys, tempfile
import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tmp = tempfile.mkdtemp()
    exe = os.path.join(tmp, "gen_mesh")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", f"{REF}/test/test_vk",
                           os.path.join(ROOT, "tools/gen_mesh_fixture.cpp"),
                           f"{REF}/test/test_vk/tiny_obj_loader.cc", "-o", exe])
    for name in ("cornell_box", "sponza"):
        out = os.path.join(tmp, name + ".bin")
        subprocess.check_call([exe, f"{REF}/resources/{name}.obj", out])
        raw = np.fromfile(out, dtype=np.uint8)
        V, T = np.frombuffer(raw[:8], dtype=np.uint32)
        o = 8
        pos = np.frombuffer(raw[o:o + 12 * V], dtype=np.float32).reshape(V, 3); o += 12 * V
        idx = np.frombuffer(raw[o:o + 12 * T], dtype=np.uint32).reshape(T, 3); o += 12 * T
        S = int(np.frombuffer(raw[o:o + 4], dtype=np.uint32)[0]); o += 4
        first = np.frombuffer(raw[o:o + 4 * (S + 1)], dtype=np.uint32)
        dst = os.path.join(ROOT, "tests/golden", name + ".npz")
        np.savez_compressed(dst, positions=pos, indices=idx, shape_first_triangle=first)
        print(dst, pos.shape, idx.shape, S, os.path.getsize(dst))


if __name__ == "__main__":
    sys.exit(main())
