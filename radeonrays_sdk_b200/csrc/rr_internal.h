// rr_internal.h -- shared declarations of the CUDA backend (host runtime <-> kernel launchers).
//
// Layering (the analogue of the reference's base/ + vlk/ split):
//   rr_api.cpp      C ABI shim: null checks -> RR_ERROR_INVALID_PARAMETER, exceptions -> RR_ERROR_INTERNAL
//                   + context / command stream / event / device pointer objects over the CUDA runtime
//   rr_build.cu     HLBVH build + refit + TLAS kernels and their launch sequences
//   rr_sort.cu      onesweep LSD radix sort (stable key-value)
//   rr_treelet.cu   treelet restructuring
//   rr_trace.cu     closest / any-hit traversal, one and two level
// There is no CPU fallback anywhere: every entry point below launches CUDA kernels.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "radeonrays.h"

namespace rr
{
constexpr int      kMaxDevices = 64;  // per-device one-time initialisation flags
constexpr uint32_t kInvalid  = 0xFFFFFFFFu;
constexpr uint32_t kSentinel = 0xFFFFFFFEu;  // isect_2l.comp RR_TOP_LEVEL_SENTINEL
// bits of DeviceInfo::error_word
constexpr uint32_t kErrorTraceStackOverflow = 1u, kErrorEmitListOverflow = 2u;

// The `update` word of an internal node (VkBvhNode::update, refit scratch state in the reference and not part of the node parity)
// carries what the packet traversal wants to know before it fetches a child:
//   bit 0        the refit's rendezvous parity (toggled with atomicXor)
//   bits 1 / 2   child0 / child1 is a leaf
//   bit 3        order the children by a vote over the rays' entry distances instead of the table below (trees that were not
//                treelet-optimised: on the plain LBVH the static order visits 8 % more nodes, on the optimised tree as many)
//   bits 8..15   for each direction octant (bit a of the octant set = the direction is negative along axis a): enter child1 first.
//                The children are ordered along the axis on which their boxes' centres are furthest apart; in the CPU model of
//                the packet walk this static order visits as many nodes as a vote over the rays' entry distances.
//   bits 16..31  a tag that says "written by this builder" -- a node array that comes from elsewhere (a dump of the reference's)
//                has small counters there and is traced by the per-ray kernel.
constexpr uint32_t kNodeTag = 0x52A50000u, kNodeTagMask = 0xFFFF0000u, kNodeLeaf0 = 2u, kNodeLeaf1 = 4u, kNodeVoteOrder = 8u, kNodeOrderShift = 8u;
__host__ __device__ inline uint32_t node_order_bits(float3 lo0, float3 hi0, float3 lo1, float3 hi1)
{
    const float cx = (lo1.x + hi1.x) - (lo0.x + hi0.x), cy = (lo1.y + hi1.y) - (lo0.y + hi0.y), cz = (lo1.z + hi1.z) - (lo0.z + hi0.z);
    const float ax = cx < 0.f ? -cx : cx, ay = cy < 0.f ? -cy : cy, az = cz < 0.f ? -cz : cz;
    uint32_t octants_negative;  // the octants whose direction is negative along the chosen axis
    float    c;
    if (ax >= ay && ax >= az) { octants_negative = 0xAAu; c = cx; }
    else if (ay >= az)        { octants_negative = 0xCCu; c = cy; }
    else                      { octants_negative = 0xF0u; c = cz; }
    // child1 lies further along the axis (c > 0): it is the near one for the negative directions; else for the positive ones
    return (c < 0.f ? ~octants_negative & 0xFFu : octants_negative) << kNodeOrderShift;
}
__host__ __device__ inline uint32_t node_update_word(uint32_t child0, uint32_t child1, uint32_t first_leaf, uint32_t parity, uint32_t order_bits)
{
    return kNodeTag | order_bits | (child0 >= first_leaf ? kNodeLeaf0 : 0u) | (child1 >= first_leaf ? kNodeLeaf1 : 0u) | (parity & 1u);
}

// 64-byte BVH2 node, identical to the reference's layout (vlk/kernels/bvh2.h:25-35) so that a raw
// dump of a geometry buffer is a VkBvhNode[] that bvh_analyzer can load (bvh_analyzer/transform.h:31-41).
//   q0 = (aabb0_min | v0, child0)   q1 = (aabb0_max | v1, child1)
//   q2 = (aabb1_min | v2, parent)   q3 = (aabb1_max | 0 , update)
struct alignas(64) Node
{
    float4 q0, q1, q2, q3;
};
static_assert(sizeof(Node) == 64, "node must be 64 bytes");

// Per-instance record of a scene buffer (64 B, one cache line): inverse transform rows + BLAS base.
struct alignas(64) InstanceRecord
{
    float4          inv0, inv1, inv2;  // object<-world rows (m0, m1, m2 of common.h Transform)
    const Node*     blas;              // device address of the instance's BLAS node 0
    unsigned long long pad;
};
static_assert(sizeof(InstanceRecord) == 64, "instance record must be 64 bytes");

// First 64 bytes of a scene (TLAS) buffer.  The reference keeps "this buffer is a scene with n instances" in a per-context map
// keyed by the buffer handle (vlk/intersector.cpp:86,263,289-290); here the buffer says so itself, so that a scene built through
// one context can be traced through another, copied, or broadcast to another GPU (SURVEY.md section 8b "Ownership").
// rrCmdIntersect launches the one-level and the two-level kernels; each reads these words on the device and returns at once when
// the buffer is not of its kind.  The magic cannot be the start of a BLAS: word 3 would be node 0's child0, which is a node index
// or 0xFFFFFFFF, and words 0-2 would be box coordinates, which are never these NaN payloads.
struct alignas(64) SceneHeader
{
    uint32_t magic[4];
    uint32_t version;
    uint32_t instance_count;
    uint32_t pad0[2];
    uint64_t nodes_off, records_off, fwd_off;  // byte offsets from the start of the buffer
    uint64_t pad1;
};
static_assert(sizeof(SceneHeader) == 64, "scene header must be 64 bytes");
constexpr uint32_t kSceneMagic0 = 0x7FC05252u, kSceneMagic1 = 0x7FC05343u, kSceneMagic2 = 0x7FC04E45u, kSceneMagic3 = 0xFFFFFFFCu;
constexpr uint32_t kSceneVersion = 2;

// Host-side instance description consumed at record time (vlk/intersector.cpp:40-45, 222-247).
struct InstanceDesc
{
    float       m[12];
    const Node* blas;
    uint32_t    index;
    uint32_t    pad;
};
static_assert(sizeof(InstanceDesc) == 64, "instance desc must be 64 bytes");

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct CudaError : std::runtime_error
{
    cudaError_t code;
    CudaError(cudaError_t c, const char* what) : std::runtime_error(std::string(what) + ": " + cudaGetErrorString(c)), code(c) {}
};
#define RR_CUDA_CHECK(expr)                                                  \
    do                                                                       \
    {                                                                        \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) throw ::rr::CudaError(_e, #expr);             \
    } while (0)

// ---- launch context handed to every launcher -------------------------------------------------------
struct DeviceInfo
{
    int      device      = 0;
    int      sm_count    = 148;
    size_t   l2_bytes    = 0;
    uint64_t* launches   = nullptr;  // per-context launch counter (host side)
    bool      morton63   = false;    // RR_CUDA_OPTION_MORTON_BITS = 63: geometry builds use 21 bits per axis (extension)
    bool      sort_rays  = false;    // RR_CUDA_OPTION_SORT_RAYS: bin the rays of every rrCmdIntersect on the device before tracing
    uint32_t  ray_grid_width = 0;    // RR_CUDA_OPTION_RAY_GRID_WIDTH: 0 = look for the row length of image-ordered rays, 1 = never, W = told
    uint32_t  refit_list_capacity = 0;  // test hook (RR_CUDA_OPTION_DEBUG_REFIT_LIST_CAPACITY): 0 = sized by the library
    uint32_t* error_word = nullptr;  // device-visible (host-mapped) word the kernels OR error bits into; checked by rrWaitEvent
};

// ---- sort (rr_sort.cu) ------------------------------------------------------------------------------
struct SortLayout
{
    uint32_t n = 0, tiles = 0;
    size_t   hist_off = 0, counter_off = 0, status_off = 0, tmp_keys_off = 0, tmp_vals_off = 0, total = 0;
};
SortLayout sort_layout(uint32_t n);
// Zeroes histogram / tile counters / look-back status (must precede histogram accumulation).
void sort_reset(const DeviceInfo& dev, cudaStream_t s, const SortLayout& L, void* scratch);
// Accumulates the 4x256 digit histograms of keys (skipped when the producer kernel already did).
void sort_histogram(const DeviceInfo& dev, cudaStream_t s, const SortLayout& L, void* scratch, const uint32_t* keys);
// 4 onesweep passes. keys_in is clobbered only if it aliases nothing else; result lands in keys_out/vals_out.
// vals_in == nullptr means "values are 0..n-1" (generated on the fly in the first pass).
// `passes` least-significant 8-bit digits are sorted (4 = full 32-bit keys).
void sort_pairs(const DeviceInfo& dev, cudaStream_t s, const SortLayout& L, void* scratch, uint32_t* keys_in,
                const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, int passes = 4);
uint32_t* sort_hist_ptr(const SortLayout& L, void* scratch);

// ---- build (rr_build.cu) ------------------------------------------------------------------------------
struct MeshDesc
{
    const float*    vertices;       // device
    uint32_t        vertex_count;
    uint32_t        stride_floats;  // vertex_stride >> 2 (lbvh_calc_mesh_aabb.comp:128)
    const uint32_t* indices;        // device, 3 per triangle: u32, or u16 when index16 (RR_INDEX_TYPE_UINT16; the reference ignores index_type)
    uint32_t        triangle_count;
    uint32_t        index16;        // 1: indices are 16-bit
};
// A geometry made of several triangle meshes (RRGeometryBuildInput::primitive_count > 1; the reference asserts it out,
// vlk/intersector.cpp:110).  The meshes are concatenated on the device into one vertex / index array at the front of the temporary
// buffer (k_merge_meshes) and the single-mesh pipeline runs on that; a hit's prim_id is the triangle's running index over the meshes
// in input order.
constexpr int kMaxMeshesPerGeometry = 16;
struct MeshGroup
{
    uint32_t count = 1;                           // 1: `mesh[0]` is used as it is, nothing is merged
    MeshDesc mesh[kMaxMeshesPerGeometry];
    uint32_t tri_first[kMaxMeshesPerGeometry + 1];   // running triangle / vertex offsets
    uint32_t vert_first[kMaxMeshesPerGeometry + 1];
    uint32_t triangles() const { return count == 1 ? mesh[0].triangle_count : tri_first[count]; }
    uint32_t vertices() const { return count == 1 ? mesh[0].vertex_count : vert_first[count]; }
};
size_t   merge_scratch_size(const MeshGroup& g);   // bytes the merged arrays take at the front of the temporary buffer (0 for one mesh)
MeshDesc merge_meshes(const DeviceInfo& dev, cudaStream_t s, const MeshGroup& g, void* scratch);   // -> the mesh to build / refit

struct BlasLayout
{
    uint32_t   n = 0;
    size_t     aabb_off = 0, codes_off = 0, sorted_codes_off = 0, sorted_refs_off = 0, sort_off = 0, lists_off = 0;
    size_t     treelet_off = 0, treelet_size = 0;
    // 63-bit Morton builds: low / high code words, the two permutations of the two-stage sort, the sorted 64-bit codes
    bool       morton63 = false;
    size_t     hi_off = 0, sorted_lo_off = 0, perm1_off = 0, hi_gathered_off = 0, sorted_hi_off = 0, perm2_off = 0, codes64_off = 0;
    SortLayout sort;
    size_t     scratch_total = 0, result_total = 0;
    // private tail of the geometry buffer, after the VkBvhNode[2N-1] array: what a refit needs to re-run the closed-form
    // emission without the sorted codes -- [256 B header: word 0 = 1 while the tree is Karras-numbered | delta(j, j+1) per
    // sorted leaf, 1 B | primitive id per sorted leaf, 4 B]
    size_t     tail_off = 0, tail_deltas_off = 0, tail_refs_off = 0;
};
BlasLayout blas_layout(uint32_t triangle_count, bool restructure, bool morton63 = false);
void build_blas(const DeviceInfo& dev, cudaStream_t s, const MeshDesc& mesh, const BlasLayout& L, void* scratch, Node* nodes,
                bool restructure);
size_t update_scratch_size(uint32_t triangle_count);
// scratch may be null / too small: the refit then runs as one kernel without work lists (slower on large meshes).
void update_blas(const DeviceInfo& dev, cudaStream_t s, const MeshDesc& mesh, Node* nodes, void* scratch, size_t scratch_bytes);

struct SceneLayout
{
    uint32_t   n = 0;
    size_t     nodes_off = 0, records_off = 0, fwd_off = 0, result_total = 0;  // scene buffer (SceneHeader at offset 0)
    size_t     desc_off = 0, boxes_off = 0, aabb_off = 0, codes_off = 0, sorted_codes_off = 0, sorted_refs_off = 0,
               sort_off = 0, lists_off = 0, scratch_total = 0;  // temporary buffer
    SortLayout sort;
};
SceneLayout scene_layout(uint32_t instance_count);
// descs: host array already staged into pinned memory by the caller; copied to scratch on the stream.
void build_scene(const DeviceInfo& dev, cudaStream_t s, const InstanceDesc* host_descs, const SceneLayout& L, void* scratch,
                 void* scene, bool reference_corner_quirk);
// Rewrites every instance record of `scene` that points at `old_blas` to `new_blas` (after a scene buffer built on another GPU, or
// next to other BLAS addresses, has been copied here).  Reads the instance count from the scene header on the device.
void rebind_scene(const DeviceInfo& dev, cudaStream_t s, void* scene, const void* old_blas, const void* new_blas);

// ---- treelets (rr_treelet.cu) ---------------------------------------------------------------------------
size_t treelet_scratch_size(uint32_t n);
void   restructure_blas(const DeviceInfo& dev, cudaStream_t s, Node* nodes, uint32_t n, void* scratch);

// ---- trace (rr_trace.cu) ----------------------------------------------------------------------------------
struct TraceArgs
{
    const void*           scene;        // geometry buffer (BLAS: VkBvhNode[2N-1]) or scene buffer (SceneHeader first); told apart on the device
    const RRRay*          rays;
    uint32_t              ray_count;
    const uint32_t*       indirect_count;  // optional device counter
    void*                 hits;
    uint32_t*             scratch;      // spill arena
    size_t                scratch_bytes;
    RRIntersectQuery      query;
    RRIntersectQueryOutput output;
    bool                  first_found_tie_rule;
};
size_t trace_scratch_size(const DeviceInfo& dev, uint32_t ray_count);
void   trace(const DeviceInfo& dev, cudaStream_t s, const TraceArgs& a);

}  // namespace rr
