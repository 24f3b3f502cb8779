// test_rr_c_api.cpp -- the scenarios of the reference's API-level tests, written in C++ against libradeonrays_b200.so
// exactly as a RadeonRays client would call it (plain rr* C ABI + CUDA runtime for the client's own buffers):
//   CreateContext, BuildSingleTriangle, BuildObj, UpdateObj   test/test_vk/basic_test.h:266-372, 373-750, 1071-1345
//   BuildObj2Level (one BLAS per OBJ shape, identity transforms) test/test_vk/basic_test.h:752-1069
//   InternalResources (rrCreateContext + rrAllocateDeviceBuffer + rrMapDevicePtr)  test/test_vk/internal_resources_test.h:51-236
// The reference's versions only assert RR_SUCCESS and write JPEGs; this program additionally dumps every hit buffer so
// that tests/test_gpu_c_client.py can compare them bit for bit with the CPU oracle.
//
// usage: test_rr_c_api <dir>     reads  <dir>/positions.bin (f32 xyz), indices.bin (u32 x3), shapes.bin (u32 first triangle
//                                       per shape, S+1 entries)
//                                writes <dir>/<scenario>.hits (RRHit[]) and <dir>/<scenario>.nodes (VkBvhNode[])
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "radeonrays.h"
#include "radeonrays_cuda.h"

#define CHECK_RR_CALL(x)                                                                     \
    do                                                                                       \
    {                                                                                        \
        RRError e_ = (x);                                                                    \
        if (e_ != RR_SUCCESS) { std::fprintf(stderr, "%s:%d %s -> RRError %d\n", __FILE__, __LINE__, #x, (int)e_); std::exit(1); } \
    } while (0)
#define CHECK_CUDA(x)                                                                        \
    do                                                                                       \
    {                                                                                        \
        cudaError_t e_ = (x);                                                                \
        if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); std::exit(1); } \
    } while (0)
#define EXPECT(c)                                                                            \
    do                                                                                       \
    {                                                                                        \
        if (!(c)) { std::fprintf(stderr, "%s:%d expectation failed: %s\n", __FILE__, __LINE__, #c); std::exit(1); } \
    } while (0)

namespace
{
template <class T>
std::vector<T> read_file(const std::string& path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
    const size_t bytes = (size_t)f.tellg();
    std::vector<T> v(bytes / sizeof(T));
    f.seekg(0);
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
    return v;
}
template <class T>
void write_file(const std::string& path, const std::vector<T>& v)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}

// A client-owned device buffer wrapped for the library (the role VkBuffer + rrGetDevicePtrFromVkBuffer play in the reference tests).
struct Buffer
{
    void*       mem = nullptr;
    RRDevicePtr ptr = nullptr;
    size_t      size = 0;
    Buffer(RRContext ctx, size_t bytes) : size(bytes)
    {
        CHECK_CUDA(cudaMalloc(&mem, bytes ? bytes : 1));
        CHECK_RR_CALL(rrGetDevicePtrFromCudaPtr(ctx, mem, 0, &ptr));
    }
    template <class T>
    void upload(const std::vector<T>& v)
    {   // cudaMemcpy from pageable memory may return before the DMA has landed, and the library's command streams are
        // non-blocking streams (no implicit ordering with the legacy stream): wait for the copy before anything is submitted
        CHECK_CUDA(cudaMemcpy(mem, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        CHECK_CUDA(cudaDeviceSynchronize());
    }
    template <class T>
    std::vector<T> download(size_t count) const
    {
        std::vector<T> v(count);
        CHECK_CUDA(cudaMemcpy(v.data(), mem, count * sizeof(T), cudaMemcpyDeviceToHost));
        return v;
    }
    void release(RRContext ctx)
    {
        CHECK_RR_CALL(rrReleaseDevicePtr(ctx, ptr));
        CHECK_CUDA(cudaFree(mem));
    }
};

void submit_and_wait(RRContext ctx, RRCommandStream cs)
{
    RREvent ev = nullptr;
    CHECK_RR_CALL(rrSumbitCommandStream(ctx, cs, nullptr, &ev));
    CHECK_RR_CALL(rrWaitEvent(ctx, ev));
    CHECK_RR_CALL(rrReleaseEvent(ctx, ev));
}

// canonical ray set of the reference tests, basic_test.h:527-550
std::vector<RRRay> sponza_rays(uint32_t res)
{
    std::vector<RRRay> rays(size_t(res) * res);
    for (uint32_t x = 0; x < res; ++x)
        for (uint32_t y = 0; y < res; ++y)
        {
            RRRay& r       = rays[size_t(res) * y + x];
            r.origin[0]    = 0.f;
            r.origin[1]    = 15.f;
            r.origin[2]    = 0.f;
            r.direction[0] = -1.f;
            r.direction[1] = -1.f + (2.f / res) * y;
            r.direction[2] = -1.f + (2.f / res) * x;
            r.min_t        = 0.001f;
            r.max_t        = 100000.f;
        }
    return rays;
}

struct Mesh
{
    std::vector<float>    positions;
    std::vector<uint32_t> indices;
    std::vector<uint32_t> shapes;
};

struct BuiltGeometry
{
    Buffer               vertices, indices, scratch, geometry;
    RRGeometryBuildInput input{};
    RRTriangleMeshPrimitive mesh{};
    RRMemoryRequirements reqs{};
    uint32_t             triangles;
    BuiltGeometry(RRContext ctx, const std::vector<float>& pos, const uint32_t* idx, uint32_t tri_count, const RRBuildOptions* options)
        : vertices(ctx, pos.size() * sizeof(float)), indices(ctx, size_t(tri_count) * 3 * sizeof(uint32_t)), scratch(ctx, 0), geometry(ctx, 0),
          triangles(tri_count)
    {
        vertices.upload(pos);
        CHECK_CUDA(cudaMemcpy(indices.mem, idx, size_t(tri_count) * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CHECK_CUDA(cudaDeviceSynchronize());
        mesh.vertices         = vertices.ptr;
        mesh.vertex_count     = (uint32_t)(pos.size() / 3);
        mesh.vertex_stride    = 3 * sizeof(float);
        mesh.triangle_indices = indices.ptr;
        mesh.triangle_count   = tri_count;
        mesh.index_type       = RR_INDEX_TYPE_UINT32;
        input.primitive_type  = RR_PRIMITIVE_TYPE_TRIANGLE_MESH;
        input.primitive_count = 1;
        CHECK_RR_CALL(rrGetGeometryBuildMemoryRequirements(ctx, bound_input(), options, &reqs));
        scratch.release(ctx);
        geometry.release(ctx);
        scratch  = Buffer(ctx, reqs.temporary_build_buffer_size);
        geometry = Buffer(ctx, reqs.result_buffer_size);
    }
    const RRGeometryBuildInput* bound_input()
    {
        input.triangle_mesh_primitives = &mesh;  // re-bound on every use: the object may have moved
        return &input;
    }
    void release(RRContext ctx)
    {
        vertices.release(ctx); indices.release(ctx); scratch.release(ctx); geometry.release(ctx);
    }
};

std::vector<RRHit> trace(RRContext ctx, RRDevicePtr scene, const std::vector<RRRay>& rays, RRIntersectQuery query, std::vector<uint32_t>* ids = nullptr)
{
    Buffer d_rays(ctx, rays.size() * sizeof(RRRay)), d_hits(ctx, rays.size() * sizeof(RRHit)), d_ids(ctx, rays.size() * sizeof(uint32_t));
    d_rays.upload(rays);
    CHECK_CUDA(cudaMemset(d_hits.mem, 0, d_hits.size));
    size_t scratch_size = 0;
    CHECK_RR_CALL(rrGetTraceMemoryRequirements(ctx, (uint32_t)rays.size(), &scratch_size));
    Buffer          d_scratch(ctx, scratch_size);
    RRCommandStream cs = nullptr;
    CHECK_RR_CALL(rrAllocateCommandStream(ctx, &cs));
    CHECK_RR_CALL(rrCmdIntersect(ctx, scene, query, d_rays.ptr, (uint32_t)rays.size(), nullptr, RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, d_hits.ptr,
                                 d_scratch.ptr, cs));
    if (ids)
        CHECK_RR_CALL(rrCmdIntersect(ctx, scene, query, d_rays.ptr, (uint32_t)rays.size(), nullptr, RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID,
                                     d_ids.ptr, d_scratch.ptr, cs));
    submit_and_wait(ctx, cs);
    CHECK_RR_CALL(rrReleaseCommandStream(ctx, cs));
    auto hits = d_hits.download<RRHit>(rays.size());
    if (ids) *ids = d_ids.download<uint32_t>(rays.size());
    d_rays.release(ctx); d_hits.release(ctx); d_ids.release(ctx); d_scratch.release(ctx);
    return hits;
}

void build(RRContext ctx, BuiltGeometry& g, RRBuildOperation op, const RRBuildOptions* options)
{
    RRCommandStream cs = nullptr;
    CHECK_RR_CALL(rrAllocateCommandStream(ctx, &cs));
    CHECK_RR_CALL(rrCmdBuildGeometry(ctx, op, g.bound_input(), options, g.scratch.ptr, g.geometry.ptr, cs));
    submit_and_wait(ctx, cs);
    CHECK_RR_CALL(rrReleaseCommandStream(ctx, cs));
}

// ---- scenarios ---------------------------------------------------------------------------------------------------------
void CreateContext()
{
    RRContext ctx = nullptr;
    CHECK_RR_CALL(rrCreateContextCuda(RR_API_VERSION, 0, nullptr, &ctx));
    CHECK_RR_CALL(rrDestroyContext(ctx));
    EXPECT(rrCreateContext(RR_API_VERSION, RR_API_VK, &ctx) == RR_ERROR_UNSUPPORTED_API);
    EXPECT(rrCreateContext(RR_API_VERSION, RR_API_CUDA, nullptr) == RR_ERROR_INVALID_PARAMETER);
}

void BuildSingleTriangle(const std::string& dir)
{
    RRContext ctx = nullptr;
    CHECK_RR_CALL(rrCreateContextCuda(RR_API_VERSION, 0, nullptr, &ctx));
    const float           o = 0.7f;
    std::vector<float>    pos = {0, -o, o, -o, o, 1.0f, o, o, 1.0f};
    std::vector<uint32_t> idx = {0, 1, 2};
    RRBuildOptions        options{};
    options.build_flags = RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD;
    BuiltGeometry g(ctx, pos, idx.data(), 1, &options);
    build(ctx, g, RR_BUILD_OPERATION_BUILD, &options);
    std::vector<RRRay> rays(2);
    rays[0] = RRRay{{0, 0, 0}, 0.001f, {0, 0, 1}, 1000.f};
    rays[1] = RRRay{{0, 0, 0}, 0.001f, {0, 0, -1}, 1000.f};
    auto hits = trace(ctx, g.geometry.ptr, rays, RR_INTERSECT_QUERY_CLOSEST);
    EXPECT(hits[0].inst_id == 0 && hits[0].prim_id == 0);
    EXPECT(hits[1].inst_id == RR_INAVLID_VALUE);
    write_file(dir + "/single_triangle.nodes", g.geometry.download<uint8_t>(64));
    g.release(ctx);
    CHECK_RR_CALL(rrDestroyContext(ctx));
}

void BuildObj(const std::string& dir, const Mesh& m, uint32_t res)
{
    RRContext ctx = nullptr;
    CHECK_RR_CALL(rrCreateContextCuda(RR_API_VERSION, 0, nullptr, &ctx));
    const uint32_t n = (uint32_t)(m.indices.size() / 3);
    for (int quality = 0; quality < 2; ++quality)
    {
        RRBuildOptions options{};
        options.build_flags = quality ? 0 : RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD;
        BuiltGeometry g(ctx, m.positions, m.indices.data(), n, &options);
        for (int rep = 0; rep < 10; ++rep) build(ctx, g, RR_BUILD_OPERATION_BUILD, &options);  // 10 rebuilds like basic_test.h:485-517
        const auto            rays = sponza_rays(res);
        std::vector<uint32_t> ids;
        auto hits = trace(ctx, g.geometry.ptr, rays, RR_INTERSECT_QUERY_CLOSEST, &ids);
        const std::string tag = quality ? "obj_quality" : "obj_fast";
        write_file(dir + "/" + tag + ".hits", hits);
        write_file(dir + "/" + tag + ".ids", ids);
        write_file(dir + "/" + tag + ".nodes", g.geometry.download<uint8_t>(size_t(2 * n - 1) * 64));
        auto anyhits = trace(ctx, g.geometry.ptr, rays, RR_INTERSECT_QUERY_ANY);
        write_file(dir + "/" + tag + "_any.hits", anyhits);
        g.release(ctx);
    }
    CHECK_RR_CALL(rrDestroyContext(ctx));
}

void UpdateObj(const std::string& dir, const Mesh& m, uint32_t res)
{
    RRContext ctx = nullptr;
    CHECK_RR_CALL(rrCreateContextCuda(RR_API_VERSION, 0, nullptr, &ctx));
    const uint32_t n = (uint32_t)(m.indices.size() / 3);
    RRBuildOptions options{};
    options.build_flags = RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD | RR_BUILD_FLAG_BITS_ALLOW_UPDATE;
    BuiltGeometry g(ctx, m.positions, m.indices.data(), n, &options);
    build(ctx, g, RR_BUILD_OPERATION_BUILD, &options);
    std::vector<float> moved = m.positions;
    for (size_t i = 1; i < moved.size(); i += 3) moved[i] -= 40.f;   // basic_test.h:1218-1224: y -= 40
    g.vertices.upload(moved);
    build(ctx, g, RR_BUILD_OPERATION_UPDATE, &options);
    auto rays = sponza_rays(res);
    for (auto& r : rays) r.origin[1] -= 40.f;
    write_file(dir + "/update.hits", trace(ctx, g.geometry.ptr, rays, RR_INTERSECT_QUERY_CLOSEST));
    write_file(dir + "/update.nodes", g.geometry.download<uint8_t>(size_t(2 * n - 1) * 64));
    g.release(ctx);
    CHECK_RR_CALL(rrDestroyContext(ctx));
}

void BuildObj2Level(const std::string& dir, const Mesh& m, uint32_t res)
{
    RRContext ctx = nullptr;
    CHECK_RR_CALL(rrCreateContextCuda(RR_API_VERSION, 0, nullptr, &ctx));
    RRBuildOptions options{};
    options.build_flags = RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD;
    const size_t shapes = m.shapes.size() - 1;
    std::vector<BuiltGeometry> geoms;
    geoms.reserve(shapes);
    for (size_t s = 0; s < shapes; ++s)
    {
        const uint32_t first = m.shapes[s], count = m.shapes[s + 1] - first;
        geoms.emplace_back(ctx, m.positions, m.indices.data() + size_t(first) * 3, count, &options);
    }
    RRCommandStream cs = nullptr;
    CHECK_RR_CALL(rrAllocateCommandStream(ctx, &cs));
    for (auto& g : geoms)
        CHECK_RR_CALL(rrCmdBuildGeometry(ctx, RR_BUILD_OPERATION_BUILD, g.bound_input(), &options, g.scratch.ptr, g.geometry.ptr, cs));
    std::vector<RRInstance> instances(shapes);
    for (size_t s = 0; s < shapes; ++s)
    {
        instances[s].geometry = geoms[s].geometry.ptr;
        const float identity[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) instances[s].transform[r][c] = identity[r][c];
    }
    RRSceneBuildInput scene_input{};
    scene_input.instances      = instances.data();
    scene_input.instance_count = (uint32_t)shapes;
    RRMemoryRequirements reqs{};
    CHECK_RR_CALL(rrGetSceneBuildMemoryRequirements(ctx, &scene_input, &options, &reqs));
    Buffer scene(ctx, reqs.result_buffer_size), scratch(ctx, reqs.temporary_build_buffer_size);
    CHECK_RR_CALL(rrCmdBuildScene(ctx, &scene_input, &options, scratch.ptr, scene.ptr, cs));   // same stream: BLASes first, then the TLAS
    submit_and_wait(ctx, cs);
    CHECK_RR_CALL(rrReleaseCommandStream(ctx, cs));
    const auto            rays = sponza_rays(res);
    std::vector<uint32_t> ids;
    write_file(dir + "/two_level.hits", trace(ctx, scene.ptr, rays, RR_INTERSECT_QUERY_CLOSEST, &ids));
    write_file(dir + "/two_level.ids", ids);
    for (auto& g : geoms) g.release(ctx);
    scene.release(ctx); scratch.release(ctx);
    CHECK_RR_CALL(rrDestroyContext(ctx));
}

// internal_resources_test.h:51-236: the library owns device and buffers; build_flags = 0 => treelet restructuring on.
void InternalResources(const std::string& dir, const Mesh& m, uint32_t res)
{
    RRContext ctx = nullptr;
    CHECK_RR_CALL(rrCreateContext(RR_API_VERSION, RR_API_CUDA, &ctx));
    const uint32_t n = (uint32_t)(m.indices.size() / 3);
    RRDevicePtr vertex_ptr = nullptr, index_ptr = nullptr;
    CHECK_RR_CALL(rrAllocateDeviceBuffer(ctx, m.positions.size() * sizeof(float), &vertex_ptr));
    CHECK_RR_CALL(rrAllocateDeviceBuffer(ctx, m.indices.size() * sizeof(uint32_t), &index_ptr));
    void* map = nullptr;
    CHECK_RR_CALL(rrMapDevicePtr(ctx, vertex_ptr, &map));
    std::copy(m.positions.begin(), m.positions.end(), static_cast<float*>(map));
    CHECK_RR_CALL(rrUnmapDevicePtr(ctx, vertex_ptr, &map));
    CHECK_RR_CALL(rrMapDevicePtr(ctx, index_ptr, &map));
    std::copy(m.indices.begin(), m.indices.end(), static_cast<uint32_t*>(map));
    CHECK_RR_CALL(rrUnmapDevicePtr(ctx, index_ptr, &map));

    RRTriangleMeshPrimitive mesh{};
    mesh.vertices = vertex_ptr; mesh.vertex_count = (uint32_t)(m.positions.size() / 3); mesh.vertex_stride = 12;
    mesh.triangle_indices = index_ptr; mesh.triangle_count = n; mesh.index_type = RR_INDEX_TYPE_UINT32;
    RRGeometryBuildInput input{};
    input.primitive_type = RR_PRIMITIVE_TYPE_TRIANGLE_MESH; input.primitive_count = 1; input.triangle_mesh_primitives = &mesh;
    RRBuildOptions options{};
    options.build_flags = 0;
    RRMemoryRequirements reqs{};
    CHECK_RR_CALL(rrGetGeometryBuildMemoryRequirements(ctx, &input, &options, &reqs));
    RRDevicePtr scratch_ptr = nullptr, geometry_ptr = nullptr;
    CHECK_RR_CALL(rrAllocateDeviceBuffer(ctx, reqs.temporary_build_buffer_size, &scratch_ptr));
    CHECK_RR_CALL(rrAllocateDeviceBuffer(ctx, reqs.result_buffer_size, &geometry_ptr));
    RRCommandStream cs = nullptr;
    CHECK_RR_CALL(rrAllocateCommandStream(ctx, &cs));
    CHECK_RR_CALL(rrCmdBuildGeometry(ctx, RR_BUILD_OPERATION_BUILD, &input, &options, scratch_ptr, geometry_ptr, cs));
    submit_and_wait(ctx, cs);
    CHECK_RR_CALL(rrReleaseCommandStream(ctx, cs));

    const auto  rays = sponza_rays(res);
    RRDevicePtr rays_ptr = nullptr, hits_ptr = nullptr, trace_scratch = nullptr;
    CHECK_RR_CALL(rrAllocateDeviceBuffer(ctx, rays.size() * sizeof(RRRay), &rays_ptr));
    CHECK_RR_CALL(rrAllocateDeviceBuffer(ctx, rays.size() * sizeof(RRHit), &hits_ptr));
    CHECK_RR_CALL(rrMapDevicePtr(ctx, rays_ptr, &map));
    std::copy(rays.begin(), rays.end(), static_cast<RRRay*>(map));
    CHECK_RR_CALL(rrUnmapDevicePtr(ctx, rays_ptr, &map));
    size_t scratch_size = 0;
    CHECK_RR_CALL(rrGetTraceMemoryRequirements(ctx, (uint32_t)rays.size(), &scratch_size));
    CHECK_RR_CALL(rrAllocateDeviceBuffer(ctx, scratch_size, &trace_scratch));
    CHECK_RR_CALL(rrAllocateCommandStream(ctx, &cs));
    CHECK_RR_CALL(rrCmdIntersect(ctx, geometry_ptr, RR_INTERSECT_QUERY_CLOSEST, rays_ptr, (uint32_t)rays.size(), nullptr,
                                 RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, hits_ptr, trace_scratch, cs));
    submit_and_wait(ctx, cs);
    CHECK_RR_CALL(rrReleaseCommandStream(ctx, cs));
    CHECK_RR_CALL(rrMapDevicePtr(ctx, hits_ptr, &map));
    std::vector<RRHit> hits(static_cast<RRHit*>(map), static_cast<RRHit*>(map) + rays.size());
    CHECK_RR_CALL(rrUnmapDevicePtr(ctx, hits_ptr, &map));
    write_file(dir + "/internal.hits", hits);
    for (RRDevicePtr p : {vertex_ptr, index_ptr, scratch_ptr, geometry_ptr, rays_ptr, hits_ptr, trace_scratch}) CHECK_RR_CALL(rrReleaseDevicePtr(ctx, p));
    CHECK_RR_CALL(rrDestroyContext(ctx));
}
}  // namespace

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <dir> [resolution]\n", argv[0]); return 2; }
    const std::string dir = argv[1];
    const uint32_t    res = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 640;   // internal_resources_test.h uses 640x640
    Mesh m;
    m.positions = read_file<float>(dir + "/positions.bin");
    m.indices   = read_file<uint32_t>(dir + "/indices.bin");
    m.shapes    = read_file<uint32_t>(dir + "/shapes.bin");
    CHECK_RR_CALL(rrSetLogLevel(RR_LOG_LEVEL_WARN));
    CreateContext();
    std::puts("[ OK ] CreateContext");
    BuildSingleTriangle(dir);
    std::puts("[ OK ] BuildSingleTriangle");
    BuildObj(dir, m, res);
    std::puts("[ OK ] BuildObj");
    UpdateObj(dir, m, res);
    std::puts("[ OK ] UpdateObj");
    BuildObj2Level(dir, m, res);
    std::puts("[ OK ] BuildObj2Level");
    InternalResources(dir, m, res);
    std::puts("[ OK ] InternalResources");
    return 0;
}
