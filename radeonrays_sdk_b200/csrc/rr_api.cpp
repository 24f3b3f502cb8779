// rr_api.cpp -- the rr* C ABI over the CUDA backend.
//
// Mirrors the reference's API shim (src/core/src/radeonrays.cpp): the same null checks returning
// RR_ERROR_INVALID_PARAMETER, every C++ exception mapped to RR_ERROR_INTERNAL (CUDA out-of-memory to
// RR_ERROR_OUT_OF_DEVICE_MEMORY), wrong-backend interop calls to RR_ERROR_UNSUPPORTED_INTEROP, and the
// record-then-submit execution model: rrCmd* only append closures to the command stream, nothing runs on
// the GPU until rrSumbitCommandStream (radeonrays.cpp:504-529, vlk/device.cpp:221-240).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>

#include "radeonrays_cuda.h"
#include "radeonrays_cuda_debug.h"
#include "rr_internal.h"

namespace rr
{
// ---- logging (reference: utils/logger.h:33-79 over spdlog; rrSetLogLevel / rrSetLogFile) ---------------
struct Logger
{
    std::mutex mu;
    int        level = RR_LOG_LEVEL_WARN;
    FILE*      file  = nullptr;
    Logger()
    {
        if (const char* e = std::getenv("RR_LOG_LEVEL")) level = std::atoi(e);
    }
    static Logger& get()
    {
        static Logger L;  // thread-safe initialisation (C++11 magic static)
        return L;
    }
    void log(int lvl, const char* fmt, ...)
    {
        if (lvl < level) return;
        static const char* names[] = {"", "debug", "info", "warn", "error"};
        std::lock_guard<std::mutex> g(mu);
        FILE* out = file ? file : stderr;
        std::fprintf(out, "[rr-cuda][%s] ", names[lvl]);
        va_list ap;
        va_start(ap, fmt);
        std::vfprintf(out, fmt, ap);
        va_end(ap);
        std::fputc('\n', out);
        if (file) std::fflush(file);
    }
};
#define RR_INFO(...) ::rr::Logger::get().log(RR_LOG_LEVEL_INFO, __VA_ARGS__)
#define RR_DEBUG(...) ::rr::Logger::get().log(RR_LOG_LEVEL_DEBUG, __VA_ARGS__)
#define RR_ERR(...) ::rr::Logger::get().log(RR_LOG_LEVEL_ERROR, __VA_ARGS__)

// ---- runtime objects (reference: base/*.h interfaces, vlk/{device,command_stream,event,device_ptr}.h) ---
struct DevicePtr
{
    char*  base        = nullptr;
    size_t offset      = 0;
    size_t size        = 0;        // 0 = unknown (client memory)
    bool   owned       = false;    // allocated by rrAllocateDeviceBuffer
    bool   imported    = false;    // mapped from another process by rrCudaImportDeviceMemory
    void*  host_shadow = nullptr;  // pinned mapping while mapped
    char*  ptr() const { return base + offset; }
    size_t bytes_available() const { return size ? size - offset : (size_t)-1; }
    ~DevicePtr()
    {
        if (host_shadow) cudaFreeHost(host_shadow);
        if (owned && base) cudaFree(base);
        if (imported && base) cudaIpcCloseMemHandle(base);
    }
};

struct Event
{
    cudaEvent_t ev = nullptr;
    ~Event() { if (ev) cudaEventDestroy(ev); }
};

struct CommandStream
{
    cudaStream_t stream   = nullptr;
    bool         external = false;
    std::vector<std::function<void(cudaStream_t)>> commands;
    // A command stream that is submitted again and again (a renderer's per-frame build / update / intersect list) is captured
    // into a CUDA graph on its second submit and replayed from then on: a Sponza-sized build is 11 kernels + 3 memsets of
    // ~10 us each, so launch gaps are a fifth of it.  Streams that record host copies (scene builds) or wrap a client's
    // stream are always replayed command by command.
    bool            graphable         = true;
    int             submits           = 0;
    cudaGraphExec_t exec              = nullptr;
    size_t          captured_commands = 0;
    uint64_t        captured_launches = 0;
    ~CommandStream() { if (exec) cudaGraphExecDestroy(exec); }
};

struct Context
{
    RRApi        api        = RR_API_CUDA;
    DeviceInfo   dev;
    cudaStream_t stream     = nullptr;
    bool         own_stream = false;
    uint64_t     launches   = 0;
    bool         first_found_tie_rule   = false;
    bool         reference_corner_quirk = false;
    // Kernels cannot throw: conditions they cannot handle (a traversal stack deeper than the deep kernel's arena, an emission
    // hand-over list longer than its slots) OR a bit into this host-mapped word, and rrWaitEvent turns it into RR_ERROR_INTERNAL.
    volatile uint32_t* host_error = nullptr;
    ~Context()
    {
        if (own_stream && stream) cudaStreamDestroy(stream);
        if (host_error) cudaFreeHost(const_cast<uint32_t*>(host_error));
    }
};

static RRError map_exception()
{
    try { throw; }
    catch (const CudaError& e)
    {
        RR_ERR("%s", e.what());
        return e.code == cudaErrorMemoryAllocation ? RR_ERROR_OUT_OF_DEVICE_MEMORY : RR_ERROR_INTERNAL;
    }
    catch (const std::bad_alloc&) { return RR_ERROR_OUT_OF_HOST_MEMORY; }
    catch (const std::exception& e) { RR_ERR("%s", e.what()); return RR_ERROR_INTERNAL; }
    catch (...) { return RR_ERROR_INTERNAL; }
}

static RRError create_context(int device, void* stream, RRContext* out)
{
    int count = 0;
    RR_CUDA_CHECK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) throw std::runtime_error("no such CUDA device");
    RR_CUDA_CHECK(cudaSetDevice(device));
    std::unique_ptr<Context> ctx(new Context);
    ctx->dev.device = device;
    cudaDeviceProp prop;
    RR_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    ctx->dev.sm_count = prop.multiProcessorCount;
    ctx->dev.l2_bytes = (size_t)prop.l2CacheSize;
    ctx->dev.launches = &ctx->launches;
    {
        void* h = nullptr;
        RR_CUDA_CHECK(cudaHostAlloc(&h, 64, cudaHostAllocMapped));
        std::memset(h, 0, 64);
        ctx->host_error = static_cast<volatile uint32_t*>(h);
        void* d = nullptr;
        RR_CUDA_CHECK(cudaHostGetDevicePointer(&d, h, 0));
        ctx->dev.error_word = static_cast<uint32_t*>(d);
    }
    if (stream) ctx->stream = static_cast<cudaStream_t>(stream);
    else
    {
        RR_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    RR_INFO("CUDA context on device %d (%s, %d SMs)", device, prop.name, prop.multiProcessorCount);
    *out = reinterpret_cast<RRContext>(ctx.release());
    return RR_SUCCESS;
}

static inline Context*       C(RRContext c) { return reinterpret_cast<Context*>(c); }
static inline DevicePtr*     D(RRDevicePtr p) { return reinterpret_cast<DevicePtr*>(p); }
static inline CommandStream* S(RRCommandStream s) { return reinterpret_cast<CommandStream*>(s); }
static inline Event*         E(RREvent e) { return reinterpret_cast<Event*>(e); }

// The kernels use 16- and 32-byte vector accesses: node arrays (geometry / scene buffers) must be 64-byte aligned, ray, hit
// and temporary buffers 16-byte aligned (every cudaMalloc'ed buffer is; an interop pointer + offset may not be).
static bool aligned_to(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// A context may live on the client's own stream (rrCreateContextCuda): if the client is capturing that stream into a graph of
// its own, the library must not begin a second capture on it -- its commands are then simply recorded into the client's graph.
static bool client_is_capturing(cudaStream_t s)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone;
}

static bool wants_restructure(const RRBuildOptions* o)
{   // vlk/intersector.cpp:116-119,170
    return o && (o->build_flags & RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD) == 0;
}

// GetTriangleMeshBuildInfo, radeonrays.cpp:59-74.  The reference builds one mesh per geometry (vlk/intersector.cpp:110,146 assert
// primitive_count == 1); here up to kMaxMeshesPerGeometry meshes are accepted and concatenated on the device (rr_build.cu
// merge_meshes): prim_id = running triangle index over the meshes in input order.
static RRError meshes_from_input(const RRGeometryBuildInput* in, MeshGroup& g)
{
    if (in->primitive_type != RR_PRIMITIVE_TYPE_TRIANGLE_MESH) return RR_ERROR_NOT_IMPLEMENTED;
    if (in->primitive_count == 0 || !in->triangle_mesh_primitives) return RR_ERROR_INVALID_PARAMETER;
    if (in->primitive_count > (uint32_t)kMaxMeshesPerGeometry) return RR_ERROR_NOT_IMPLEMENTED;
    g.count = in->primitive_count;
    uint64_t tris = 0, verts = 0;
    for (uint32_t k = 0; k < g.count; ++k)
    {
        const RRTriangleMeshPrimitive& p = in->triangle_mesh_primitives[k];
        if (p.index_type != RR_INDEX_TYPE_UINT32 && p.index_type != RR_INDEX_TYPE_UINT16) return RR_ERROR_INVALID_PARAMETER;
        if (p.triangle_count == 0 || (p.vertex_stride & 3u) || p.vertex_stride < 12) return RR_ERROR_INVALID_PARAMETER;
        MeshDesc& m      = g.mesh[k];
        m.vertices       = p.vertices ? reinterpret_cast<const float*>(D(p.vertices)->ptr()) : nullptr;
        m.vertex_count   = p.vertex_count;
        m.stride_floats  = p.vertex_stride >> 2;
        m.indices        = p.triangle_indices ? reinterpret_cast<const uint32_t*>(D(p.triangle_indices)->ptr()) : nullptr;
        m.triangle_count = p.triangle_count;
        m.index16        = p.index_type == RR_INDEX_TYPE_UINT16 ? 1u : 0u;  // beyond the reference, which reads 32-bit indices whatever index_type says
        g.tri_first[k]   = (uint32_t)tris;
        g.vert_first[k]  = (uint32_t)verts;
        tris += p.triangle_count;
        verts += p.vertex_count;
    }
    if (tris > 0x3FFFFFFFull || verts > 0xFFFFFFFFull) return RR_ERROR_INVALID_PARAMETER;
    g.tri_first[g.count]  = (uint32_t)tris;
    g.vert_first[g.count] = (uint32_t)verts;
    return RR_SUCCESS;
}
}  // namespace rr

using namespace rr;

extern "C" {

RRError rrCreateContext(uint32_t api_version, RRApi api, RRContext* context)
{
    RR_INFO("rrCreateContext(%u)", api_version);
    if (!context) return RR_ERROR_INVALID_PARAMETER;
    if (api != RR_API_CUDA) return RR_ERROR_UNSUPPORTED_API;
    try { return create_context(0, nullptr, context); }
    catch (...) { return map_exception(); }
}

RRError rrCreateContextCuda(uint32_t api_version, int device_ordinal, void* cuda_stream, RRContext* context)
{
    RR_INFO("rrCreateContextCuda(%u, device %d)", api_version, device_ordinal);
    if (!context) return RR_ERROR_INVALID_PARAMETER;
    try { return create_context(device_ordinal, cuda_stream, context); }
    catch (...) { return map_exception(); }
}

RRError rrDestroyContext(RRContext context)
{
    RR_INFO("rrDestroyContext");
    if (!context) return RR_ERROR_INVALID_PARAMETER;
    delete C(context);
    return RR_SUCCESS;
}

RRError rrSetLogLevel(RRLogLevel log_level)
{
    if (log_level < RR_LOG_LEVEL_DEBUG || log_level > RR_LOG_LEVEL_OFF) return RR_ERROR_INVALID_PARAMETER;
    Logger::get().level = log_level;
    return RR_SUCCESS;
}

RRError rrSetLogFile(char const* filename)
{
    Logger& L = Logger::get();
    std::lock_guard<std::mutex> g(L.mu);
    if (L.file) { std::fclose(L.file); L.file = nullptr; }
    if (filename)
    {
        L.file = std::fopen(filename, "w");
        if (!L.file) return RR_ERROR_INTERNAL;
    }
    return RR_SUCCESS;
}

RRError rrGetGeometryBuildMemoryRequirements(RRContext context, const RRGeometryBuildInput* build_input,
                                             const RRBuildOptions* build_options, RRMemoryRequirements* memory_requirements)
{
    RR_INFO("rrGetGeometryBuildMemoryRequirements");
    if (!context || !build_input || !memory_requirements) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        MeshGroup g{};
        if (RRError e = meshes_from_input(build_input, g)) return e;
        const BlasLayout L = blas_layout(g.triangles(), wants_restructure(build_options), C(context)->dev.morton63);
        memory_requirements->result_buffer_size           = L.result_total;
        memory_requirements->temporary_build_buffer_size  = merge_scratch_size(g) + L.scratch_total;
        // work lists of the staged refit (the dx backend reports 4 N, dx/update_hlbvh.cpp:59-63; vlk reports 0).  An update
        // recorded without a temporary buffer still works, as one slower kernel (single-mesh geometries only).
        memory_requirements->temporary_update_buffer_size = merge_scratch_size(g) + update_scratch_size(g.triangles());
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrCmdBuildGeometry(RRContext context, RRBuildOperation build_operation, const RRGeometryBuildInput* build_input,
                           const RRBuildOptions* build_options, RRDevicePtr temporary_buffer, RRDevicePtr geometry_buffer,
                           RRCommandStream command_stream)
{
    RR_INFO("rrCmdBuildGeometry");
    if (!context || !command_stream || !build_input) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        Context* ctx = C(context);
        MeshGroup g{};
        if (RRError e = meshes_from_input(build_input, g)) return e;
        if (!geometry_buffer) return RR_ERROR_INVALID_PARAMETER;
        for (uint32_t k = 0; k < g.count; ++k)
            if (!g.mesh[k].vertices || !g.mesh[k].indices || !aligned_to(g.mesh[k].vertices, 4) || !aligned_to(g.mesh[k].indices, g.mesh[k].index16 ? 2 : 4))
                return RR_ERROR_INVALID_PARAMETER;
        Node*            nodes = reinterpret_cast<Node*>(D(geometry_buffer)->ptr());
        if (!aligned_to(nodes, 64) || (temporary_buffer && !aligned_to(D(temporary_buffer)->ptr(), 16))) return RR_ERROR_INVALID_PARAMETER;
        const DeviceInfo dev     = ctx->dev;
        const uint32_t   n       = g.triangles();
        const size_t     merged  = merge_scratch_size(g);   // 0 for a single mesh
        if (build_operation == RR_BUILD_OPERATION_BUILD)
        {
            if (!temporary_buffer) return RR_ERROR_INVALID_PARAMETER;
            const bool       restructure = wants_restructure(build_options);
            const BlasLayout L           = blas_layout(n, restructure, ctx->dev.morton63);
            if (D(temporary_buffer)->bytes_available() < merged + L.scratch_total || D(geometry_buffer)->bytes_available() < L.result_total)
                throw std::runtime_error("geometry build: buffer smaller than rrGetGeometryBuildMemoryRequirements reported");
            char* scratch = D(temporary_buffer)->ptr();
            S(command_stream)->commands.push_back([=](cudaStream_t s) {
                const MeshDesc m = merge_meshes(dev, s, g, scratch);
                build_blas(dev, s, m, L, scratch + merged, nodes, restructure);
            });
        }
        else
        {   // UPDATE: the temporary buffer is optional for a single mesh (the reference's Vulkan backend ignores it,
            // vlk/intersector.cpp:176-204); a multi-mesh geometry needs it for the merged arrays.
            // the update reads the tail of the geometry buffer (rr_build.cu update_blas): the whole result buffer must be there
            if (D(geometry_buffer)->bytes_available() < blas_layout(n, false).result_total)
                throw std::runtime_error("geometry update: buffer smaller than rrGetGeometryBuildMemoryRequirements reported");
            char*        scratch       = temporary_buffer ? D(temporary_buffer)->ptr() : nullptr;
            const size_t scratch_bytes = temporary_buffer ? D(temporary_buffer)->bytes_available() : 0;
            if (merged && scratch_bytes < merged) return RR_ERROR_INVALID_PARAMETER;
            S(command_stream)->commands.push_back([=](cudaStream_t s) {
                const MeshDesc m = merge_meshes(dev, s, g, scratch);
                update_blas(dev, s, m, nodes, scratch ? scratch + merged : nullptr, scratch_bytes - merged);
            });
        }
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrGetSceneBuildMemoryRequirements(RRContext context, const RRSceneBuildInput* build_input, const RRBuildOptions*,
                                          RRMemoryRequirements* memory_requirements)
{
    RR_INFO("rrGetSceneBuildMemoryRequirements");
    if (!context || !build_input || !memory_requirements) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        const SceneLayout L = scene_layout(build_input->instance_count);
        memory_requirements->result_buffer_size           = L.result_total;
        memory_requirements->temporary_build_buffer_size  = L.scratch_total;
        memory_requirements->temporary_update_buffer_size = 0;
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrCmdBuildScene(RRContext context, const RRSceneBuildInput* build_input, const RRBuildOptions*, RRDevicePtr temporary_buffer,
                        RRDevicePtr scene_buffer, RRCommandStream command_stream)
{
    RR_INFO("rrCmdBuildScene");
    if (!context || !command_stream || !build_input) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        Context* ctx = C(context);
        const uint32_t n = build_input->instance_count;
        if (n == 0 || !build_input->instances || !temporary_buffer || !scene_buffer) return RR_ERROR_INVALID_PARAMETER;
        // Host data is consumed at record time (vlk/intersector.cpp:222-247).
        auto descs = std::make_shared<std::vector<InstanceDesc>>(n);
        for (uint32_t i = 0; i < n; ++i)
        {
            if (!build_input->instances[i].geometry) return RR_ERROR_INVALID_PARAMETER;
            InstanceDesc& d = (*descs)[i];
            std::memcpy(d.m, &build_input->instances[i].transform[0][0], 12 * sizeof(float));
            d.blas  = reinterpret_cast<const Node*>(D(build_input->instances[i].geometry)->ptr());
            d.index = i;
            d.pad   = 0;
        }
        const SceneLayout L = scene_layout(n);
        if (D(temporary_buffer)->bytes_available() < L.scratch_total || D(scene_buffer)->bytes_available() < L.result_total)
            throw std::runtime_error("scene build: buffer smaller than rrGetSceneBuildMemoryRequirements reported");
        void*            scratch = D(temporary_buffer)->ptr();
        void*            scene   = D(scene_buffer)->ptr();
        if (!aligned_to(scene, 64) || !aligned_to(scratch, 16)) return RR_ERROR_INVALID_PARAMETER;
        const DeviceInfo dev     = ctx->dev;
        const bool       quirk   = ctx->reference_corner_quirk;
        S(command_stream)->graphable = false;  // the instance descriptors are copied from pageable host memory
        S(command_stream)->commands.push_back([=](cudaStream_t s) {
            build_scene(dev, s, descs->data(), L, scratch, scene, quirk);
            // the pageable host vector is staged synchronously by cudaMemcpyAsync; `descs` stays alive with the closure
        });
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrGetTraceMemoryRequirements(RRContext context, uint32_t ray_count, size_t* scratch_size)
{
    RR_INFO("rrGetTraceMemoryRequirements");
    if (!context || !scratch_size || !ray_count) return RR_ERROR_INVALID_PARAMETER;
    *scratch_size = trace_scratch_size(C(context)->dev, ray_count);
    return RR_SUCCESS;
}

RRError rrCmdIntersect(RRContext context, RRDevicePtr scene_buffer, RRIntersectQuery query, RRDevicePtr rays, uint32_t ray_count,
                       RRDevicePtr indirect_ray_count, RRIntersectQueryOutput query_output, RRDevicePtr hits, RRDevicePtr scratch,
                       RRCommandStream command_stream)
{
    RR_INFO("rrCmdIntersect");
    if (!context || !scene_buffer || !rays || !hits || !scratch || !command_stream) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        Context*  ctx = C(context);
        TraceArgs a{};
        // One level or two: the buffer says so itself (SceneHeader, rr_internal.h) and the kernels branch on it on the device, so
        // a scene built through another context, copied to another buffer or built later in the same stream is traced correctly
        // (the reference keys a host-side map by the buffer handle instead, vlk/intersector.cpp:289-324).
        a.scene          = D(scene_buffer)->ptr();
        a.rays           = reinterpret_cast<const RRRay*>(D(rays)->ptr());
        a.ray_count      = ray_count;
        a.indirect_count = indirect_ray_count ? reinterpret_cast<const uint32_t*>(D(indirect_ray_count)->ptr()) : nullptr;
        a.hits           = D(hits)->ptr();
        a.scratch        = reinterpret_cast<uint32_t*>(D(scratch)->ptr());
        a.scratch_bytes  = D(scratch)->bytes_available();
        a.query          = query;
        a.output         = query_output;
        a.first_found_tie_rule = ctx->first_found_tie_rule;
        if (!aligned_to(a.scene, 64) || !aligned_to(a.rays, 16) || !aligned_to(a.scratch, 16) ||
            !aligned_to(a.hits, query_output == RR_INTERSECT_QUERY_OUTPUT_FULL_HIT ? 16 : 4) || !aligned_to(a.indirect_count, 4))
            return RR_ERROR_INVALID_PARAMETER;
        if (ray_count && a.scratch_bytes < trace_scratch_size(ctx->dev, ray_count))
            throw std::runtime_error("intersect: scratch smaller than rrGetTraceMemoryRequirements reported");
        const DeviceInfo dev = ctx->dev;
        S(command_stream)->commands.push_back([=](cudaStream_t s) { trace(dev, s, a); });
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrAllocateCommandStream(RRContext context, RRCommandStream* command_stream)
{
    RR_INFO("rrAllocateCommandStream");
    if (!context || !command_stream) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        CommandStream* s = new CommandStream;
        s->stream        = C(context)->stream;
        *command_stream  = reinterpret_cast<RRCommandStream>(s);
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrReleaseCommandStream(RRContext context, RRCommandStream command_stream)
{
    RR_INFO("rrReleaseCommandStream");
    if (!context || !command_stream) return RR_ERROR_INVALID_PARAMETER;
    delete S(command_stream);
    return RR_SUCCESS;
}

RRError rrReleaseExternalCommandStream(RRContext context, RRCommandStream command_stream)
{
    RR_INFO("rrReleaseExternalCommandStream");
    if (!context || !command_stream) return RR_ERROR_INVALID_PARAMETER;
    delete S(command_stream);
    return RR_SUCCESS;
}

RRError rrSumbitCommandStream(RRContext context, RRCommandStream command_stream, RREvent wait_event, RREvent* out_event)
{
    RR_INFO("rrSumbitCommandStream");
    if (!context || !command_stream || !out_event) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        Context*       ctx = C(context);
        CommandStream* cs  = S(command_stream);
        RR_CUDA_CHECK(cudaSetDevice(ctx->dev.device));
        // GPU-side wait: satisfies both the Vulkan (CPU wait, vlk/device.cpp:229-232) and DX12 (queue wait) contracts.
        if (wait_event) RR_CUDA_CHECK(cudaStreamWaitEvent(cs->stream, E(wait_event)->ev, 0));
        static const bool graphs = [] { const char* e = std::getenv("RR_CUDA_GRAPHS"); return !e || std::atoi(e) != 0; }();
        ++cs->submits;
        if (graphs && cs->graphable && !cs->external && cs->exec && cs->captured_commands == cs->commands.size() && !client_is_capturing(cs->stream))
        {
            RR_CUDA_CHECK(cudaGraphLaunch(cs->exec, cs->stream));
            ctx->launches += cs->captured_launches;
        }
        else if (graphs && cs->graphable && !cs->external && cs->submits >= 2 && !cs->commands.empty() && !client_is_capturing(cs->stream))
        {   // (re)capture: the closures issue exactly the work a plain submit would
            if (cs->exec) { cudaGraphExecDestroy(cs->exec); cs->exec = nullptr; }
            const uint64_t l0 = ctx->launches;
            cudaGraph_t    graph = nullptr;
            bool           ok = cudaStreamBeginCapture(cs->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok)
            {
                try { for (auto& cmd : cs->commands) cmd(cs->stream); }
                catch (...) { ok = false; }
                if (cudaStreamEndCapture(cs->stream, &graph) != cudaSuccess || !graph) ok = false;
            }
            if (ok && cudaGraphInstantiate(&cs->exec, graph, 0) != cudaSuccess) { ok = false; cs->exec = nullptr; }
            if (graph) cudaGraphDestroy(graph);
            if (ok)
            {
                cs->captured_commands = cs->commands.size();
                cs->captured_launches = ctx->launches - l0;
                RR_CUDA_CHECK(cudaGraphLaunch(cs->exec, cs->stream));
            }
            else
            {   // not capturable: never try again, run the commands directly
                (void)cudaGetLastError();
                cs->graphable = false;
                ctx->launches = l0;
                for (auto& cmd : cs->commands) cmd(cs->stream);
            }
        }
        else
            for (auto& cmd : cs->commands) cmd(cs->stream);
        std::unique_ptr<Event> ev(new Event);
        RR_CUDA_CHECK(cudaEventCreateWithFlags(&ev->ev, cudaEventDisableTiming));
        RR_CUDA_CHECK(cudaEventRecord(ev->ev, cs->stream));
        *out_event = reinterpret_cast<RREvent>(ev.release());
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrReleaseEvent(RRContext context, RREvent event)
{
    RR_INFO("rrReleaseEvent");
    if (!context || !event) return RR_ERROR_INVALID_PARAMETER;
    delete E(event);
    return RR_SUCCESS;
}

RRError rrWaitEvent(RRContext context, RREvent event)
{
    RR_INFO("rrWaitEvent");
    if (!context || !event) return RR_ERROR_INVALID_PARAMETER;
    try { RR_CUDA_CHECK(cudaEventSynchronize(E(event)->ev)); }
    catch (...) { return map_exception(); }
    Context* ctx = C(context);
    if (ctx->host_error && *ctx->host_error)
    {
        const uint32_t bits = *ctx->host_error;
        *ctx->host_error    = 0;
        RR_ERR("device-side error bits 0x%x: %s%s", bits, (bits & kErrorTraceStackOverflow) ? "traversal stack deeper than the deep kernel's arena; " : "",
               (bits & kErrorEmitListOverflow) ? "hierarchy emission hand-over list overflow" : "");
        return RR_ERROR_INTERNAL;
    }
    return RR_SUCCESS;
}

RRError rrReleaseDevicePtr(RRContext context, RRDevicePtr ptr)
{
    RR_INFO("rrReleaseDevicePtr");
    if (!context || !ptr) return RR_ERROR_INVALID_PARAMETER;
    delete D(ptr);
    return RR_SUCCESS;
}

// ---- CUDA interop (radeonrays_cuda.h) ------------------------------------------------------------------
RRError rrGetDevicePtrFromCudaPtr(RRContext context, void* device_memory, size_t offset, RRDevicePtr* device_ptr)
{
    if (!context || !device_memory || !device_ptr) return RR_ERROR_INVALID_PARAMETER;
    if (C(context)->api != RR_API_CUDA) return RR_ERROR_UNSUPPORTED_INTEROP;
    try
    {
        DevicePtr* p = new DevicePtr;
        p->base      = static_cast<char*>(device_memory);
        p->offset    = offset;
        *device_ptr  = reinterpret_cast<RRDevicePtr>(p);
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrGetCommandStreamFromCudaStream(RRContext context, void* cuda_stream, RRCommandStream* command_stream)
{
    // NULL is the legacy default stream, a valid cudaStream_t, so only the out pointer is checked.
    if (!context || !command_stream) return RR_ERROR_INVALID_PARAMETER;
    if (C(context)->api != RR_API_CUDA) return RR_ERROR_UNSUPPORTED_INTEROP;
    try
    {
        CommandStream* s = new CommandStream;
        s->stream        = static_cast<cudaStream_t>(cuda_stream);
        s->external      = true;
        *command_stream  = reinterpret_cast<RRCommandStream>(s);
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrAllocateDeviceBuffer(RRContext context, size_t size, RRDevicePtr* device_ptr)
{
    if (!context || !size || !device_ptr) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        RR_CUDA_CHECK(cudaSetDevice(C(context)->dev.device));
        std::unique_ptr<DevicePtr> p(new DevicePtr);
        void* mem = nullptr;
        RR_CUDA_CHECK(cudaMalloc(&mem, size));
        p->base  = static_cast<char*>(mem);
        p->size  = size;
        p->owned = true;
        *device_ptr = reinterpret_cast<RRDevicePtr>(p.release());
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrMapDevicePtr(RRContext context, RRDevicePtr device_ptr, void** mapping_ptr)
{
    if (!context || !device_ptr || !mapping_ptr) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        DevicePtr* p = D(device_ptr);
        if (!p->size) throw std::runtime_error("rrMapDevicePtr: only buffers from rrAllocateDeviceBuffer can be mapped");
        RR_CUDA_CHECK(cudaSetDevice(C(context)->dev.device));
        if (!p->host_shadow) RR_CUDA_CHECK(cudaMallocHost(&p->host_shadow, p->size));
        // Everything previously submitted must be visible in the mapping.
        RR_CUDA_CHECK(cudaDeviceSynchronize());
        RR_CUDA_CHECK(cudaMemcpy(p->host_shadow, p->base, p->size, cudaMemcpyDeviceToHost));
        *mapping_ptr = static_cast<char*>(p->host_shadow) + p->offset;
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrUnmapDevicePtr(RRContext context, RRDevicePtr device_ptr, void** mapping_ptr)
{
    if (!context || !device_ptr || !mapping_ptr) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        DevicePtr* p = D(device_ptr);
        if (!p->host_shadow) throw std::runtime_error("rrUnmapDevicePtr: buffer is not mapped");
        RR_CUDA_CHECK(cudaSetDevice(C(context)->dev.device));
        RR_CUDA_CHECK(cudaMemcpy(p->base, p->host_shadow, p->size, cudaMemcpyHostToDevice));
        RR_CUDA_CHECK(cudaFreeHost(p->host_shadow));
        p->host_shadow = nullptr;
        *mapping_ptr   = nullptr;
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrGetCudaPtrFromDevicePtr(RRContext context, RRDevicePtr device_ptr, void** device_memory)
{
    if (!context || !device_ptr || !device_memory) return RR_ERROR_INVALID_PARAMETER;
    *device_memory = D(device_ptr)->ptr();
    return RR_SUCCESS;
}

RRError rrCudaSetOption(RRContext context, RRCudaOption option, int value)
{
    if (!context) return RR_ERROR_INVALID_PARAMETER;
    switch (option)
    {
    case RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND: C(context)->first_found_tie_rule = value != 0; return RR_SUCCESS;
    case RR_CUDA_OPTION_REFERENCE_TRANSFORM_AABB_QUIRK: C(context)->reference_corner_quirk = value != 0; return RR_SUCCESS;
    case RR_CUDA_OPTION_MORTON_BITS:
        if (value != 30 && value != 63) return RR_ERROR_INVALID_PARAMETER;
        C(context)->dev.morton63 = value == 63;
        return RR_SUCCESS;
    case RR_CUDA_OPTION_SORT_RAYS: C(context)->dev.sort_rays = value != 0; return RR_SUCCESS;
    case RR_CUDA_OPTION_RAY_GRID_WIDTH:
        if (value < 0 || (value > 1 && value < 64)) return RR_ERROR_INVALID_PARAMETER;
        C(context)->dev.ray_grid_width = (uint32_t)value;
        return RR_SUCCESS;
    case RR_CUDA_OPTION_DEBUG_REFIT_LIST_CAPACITY: C(context)->dev.refit_list_capacity = value > 0 ? (uint32_t)value : 0u; return RR_SUCCESS;
    default: return RR_ERROR_INVALID_PARAMETER;
    }
}

RRError rrCudaCmdRebindSceneGeometry(RRContext context, RRDevicePtr scene_buffer, void* old_geometry_address, RRDevicePtr new_geometry,
                                     RRCommandStream command_stream)
{
    if (!context || !scene_buffer || !old_geometry_address || !new_geometry || !command_stream) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        const DeviceInfo dev   = C(context)->dev;
        void*            scene = D(scene_buffer)->ptr();
        const void*      neu   = D(new_geometry)->ptr();
        if (!aligned_to(scene, 64) || !aligned_to(neu, 64)) return RR_ERROR_INVALID_PARAMETER;
        S(command_stream)->commands.push_back([=](cudaStream_t s) { rebind_scene(dev, s, scene, old_geometry_address, neu); });
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

static_assert(sizeof(cudaIpcMemHandle_t) == RR_CUDA_IPC_HANDLE_SIZE, "IPC handle size");

RRError rrCudaExportDeviceMemory(RRContext context, RRDevicePtr device_ptr, void* handle_out, size_t* offset_out)
{
    if (!context || !device_ptr || !handle_out || !offset_out) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        RR_CUDA_CHECK(cudaSetDevice(C(context)->dev.device));
        // the handle names the whole allocation: find its base through the driver (cuMemGetAddressRange)
        using RangeFn = int (*)(unsigned long long*, size_t*, unsigned long long);
        static RangeFn range = [] {
            void* f = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
            return reinterpret_cast<RangeFn>(f);
        }();
        char*              p    = D(device_ptr)->ptr();
        unsigned long long base = reinterpret_cast<unsigned long long>(p);
        size_t             size = 0;
        if (!range || range(&base, &size, reinterpret_cast<unsigned long long>(p)) != 0) throw std::runtime_error("cuMemGetAddressRange failed");
        cudaIpcMemHandle_t h;
        RR_CUDA_CHECK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
        std::memcpy(handle_out, &h, sizeof(h));
        *offset_out = static_cast<size_t>(reinterpret_cast<unsigned long long>(p) - base);
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrCudaImportDeviceMemory(RRContext context, const void* handle, size_t offset, RRDevicePtr* device_ptr)
{
    if (!context || !handle || !device_ptr) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        RR_CUDA_CHECK(cudaSetDevice(C(context)->dev.device));
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle, sizeof(h));
        void* mapped = nullptr;
        RR_CUDA_CHECK(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
        DevicePtr* p = new DevicePtr;
        p->base      = static_cast<char*>(mapped);
        p->offset    = offset;
        p->imported  = true;
        *device_ptr  = reinterpret_cast<RRDevicePtr>(p);
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrCudaGetLaunchCount(RRContext context, uint64_t* launches)
{
    if (!context || !launches) return RR_ERROR_INVALID_PARAMETER;
    *launches = C(context)->launches;
    return RR_SUCCESS;
}

// ---- debug / test hooks (radeonrays_cuda_debug.h) ------------------------------------------------------
RRError rrCudaDebugGetBuildScratchLayout(RRContext context, uint32_t triangle_count, RRCudaBuildScratchLayout* layout)
{
    if (!context || !layout || !triangle_count) return RR_ERROR_INVALID_PARAMETER;
    const BlasLayout L         = blas_layout(triangle_count, false, C(context)->dev.morton63);
    layout->scene_aabb_offset   = L.aabb_off;
    layout->morton_codes_offset = L.codes_off;
    layout->sorted_codes_offset = L.morton63 ? L.codes64_off : L.sorted_codes_off;   // u64[N] with RR_CUDA_OPTION_MORTON_BITS = 63
    layout->sorted_refs_offset  = L.tail_refs_off;  // inside the GEOMETRY buffer
    layout->sort_tmp_values_offset = L.sort_off + L.sort.tmp_vals_off;
    return RR_SUCCESS;
}

RRError rrCudaDebugSortPairs(RRContext context, void* keys_in, void* values_in, void* keys_out, void* values_out, uint32_t count)
{
    if (!context || !keys_in || !keys_out || !values_out) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        Context* ctx = C(context);
        RR_CUDA_CHECK(cudaSetDevice(ctx->dev.device));
        if (count == 0) return RR_SUCCESS;
        const SortLayout L = sort_layout(count);
        void* scratch = nullptr;
        RR_CUDA_CHECK(cudaMalloc(&scratch, L.total));
        try
        {
            sort_reset(ctx->dev, ctx->stream, L, scratch);
            sort_histogram(ctx->dev, ctx->stream, L, scratch, static_cast<const uint32_t*>(keys_in));
            sort_pairs(ctx->dev, ctx->stream, L, scratch, static_cast<uint32_t*>(keys_in), static_cast<const uint32_t*>(values_in),
                       static_cast<uint32_t*>(keys_out), static_cast<uint32_t*>(values_out));
            RR_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
        catch (...) { cudaFree(scratch); throw; }
        RR_CUDA_CHECK(cudaFree(scratch));
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrCudaDebugRestructure(RRContext context, RRDevicePtr geometry, uint32_t triangle_count, RRDevicePtr temporary_buffer)
{
    if (!context || !geometry || !temporary_buffer || !triangle_count) return RR_ERROR_INVALID_PARAMETER;
    try
    {
        Context* ctx = C(context);
        RR_CUDA_CHECK(cudaSetDevice(ctx->dev.device));
        if (D(temporary_buffer)->bytes_available() < treelet_scratch_size(triangle_count))
            throw std::runtime_error("restructure: temporary buffer too small");
        restructure_blas(ctx->dev, ctx->stream, reinterpret_cast<Node*>(D(geometry)->ptr()), triangle_count, D(temporary_buffer)->ptr());
        RR_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    catch (...) { return map_exception(); }
    return RR_SUCCESS;
}

RRError rrCudaDebugGetSceneLayout(RRContext context, uint32_t instance_count, RRCudaSceneLayout* layout)
{
    if (!context || !layout || !instance_count) return RR_ERROR_INVALID_PARAMETER;
    const SceneLayout L               = scene_layout(instance_count);
    layout->nodes_offset              = L.nodes_off;
    layout->records_offset            = L.records_off;
    layout->forward_transforms_offset = L.fwd_off;
    return RR_SUCCESS;
}

}  // extern "C"
