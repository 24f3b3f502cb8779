// Fixture generator (run in the build container only; /root/reference is not on the GPU box).
//
// Loads an OBJ through the reference's OWN test loader (test/test_vk/mesh_data.h:61-108, which
// de-duplicates (v,vn,vt) triples and emits u32 indices / 12-byte positions) so that the
// committed fixtures are byte-identical to what RadeonRays' own tests feed rrCmdBuildGeometry
// (test/test_vk/basic_test.h:387).  The reference headers are included from where they lie;
// nothing is copied into this repository.
//
// Output (little endian): u32 vertex_count, u32 triangle_count, f32 positions[3*V], u32 indices[3*T],
// then u32 shape_count, u32 shape_first_index[shape_count+1] (offsets into `indices`, in elements/3)
// -- the per-shape split is what BuildObj2Level uses (basic_test.h:752-1069).
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <vector>
#include "mesh_data.h"

int main(int argc, char** argv)
{
    if (argc != 3) { std::fprintf(stderr, "usage: %s in.obj out.bin\n", argv[0]); return 1; }
    MeshData mesh(argv[1]);
    // Per-shape triangle ranges: reload shapes to count indices per shape.
    tinyobj::attrib_t attrib; std::vector<tinyobj::shape_t> shapes; std::vector<tinyobj::material_t> mats;
    std::string warn, err;
    tinyobj::LoadObj(&attrib, &shapes, &mats, &warn, &err, argv[1], "");
    std::vector<uint32_t> first{0};
    for (auto& s : shapes) first.push_back(first.back() + (uint32_t)s.mesh.indices.size() / 3);
    FILE* f = std::fopen(argv[2], "wb");
    uint32_t V = (uint32_t)mesh.positions.size() / 3, T = (uint32_t)mesh.indices.size() / 3;
    std::fwrite(&V, 4, 1, f); std::fwrite(&T, 4, 1, f);
    std::fwrite(mesh.positions.data(), 4, mesh.positions.size(), f);
    std::fwrite(mesh.indices.data(), 4, mesh.indices.size(), f);
    uint32_t S = (uint32_t)shapes.size();
    std::fwrite(&S, 4, 1, f); std::fwrite(first.data(), 4, first.size(), f);
    std::fclose(f);
    std::printf("%s: V=%u T=%u shapes=%u\n", argv[1], V, T, S);
    return 0;
}
