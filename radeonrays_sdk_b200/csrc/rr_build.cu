// rr_build.cu -- HLBVH build, refit and top-level build for sm_100a.
//
// What the reference does in 45+ dispatches (vlk/hlbvh_builder.cpp:159-351: init, scene AABB, Morton,
// 8 x (histogram, scan, scatter), emit hierarchy, fit) runs here as:
//   3 memset nodes -> k_scene_aabb -> k_morton (+ digit histograms) -> 4 onesweep passes ->
//   k_emit_leaves -> k_emit_window -> k_emit_upper x ceil(log16(N/512))
// The k_emit_* kernels fuse hierarchy emission and bottom-up fitting in closed form (group_merge): every node's 64
// bytes are written once, without atomics.  They produce the SAME tree, node numbering (Karras: internal i in [0,N-1),
// leaf j at N-1+j, root 0) and boxes as lbvh_emit_hierarchy_mesh.comp:170-217 + lbvh_fit_aabb_mesh.comp:122-205 -- see
// the comment above group_merge.  RR_BUILD_OPERATION_UPDATE re-runs them from the deltas and sorted ids kept in the
// geometry buffer's tail (update_blas), or, for treelet-restructured trees, runs the generic staged refit (k_refit*).
// All arithmetic that feeds Morton codes is plain IEEE binary32 (the library is compiled with --fmad=false; SURVEY.md
// App. B2).
//
// Algorithmic HBM bytes per triangle (u32 indices, 12-B vertices): aabb 48 + morton 48 + code write 4 +
// sort 60 (identity values: no ref write/read in pass 0) + emit/fit: codes 8, refs 4, gather 48, nodes 128.
#include <algorithm>
#include <cstdlib>

#include <cuda.h>  // CUtensorMap (encoded through the runtime's driver entry point: libcuda is not linked)

#include <mutex>

#include "rr_internal.h"

namespace rr
{
namespace
{
// ---- ordered-uint float encoding (reference: vlk/kernels/common.h:68-90) -----------------------------
__device__ __forceinline__ uint32_t float_to_ordered(float f)
{
    uint32_t v = __float_as_uint(f);
    return v ^ ((1u + ~(v >> 31)) | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t v) { return __uint_as_float(v ^ (((v >> 31) - 1u) | 0x80000000u)); }

__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(p[0], p[1], p[2]); }
// The three vertex indices of a triangle: 32-bit, or 16-bit when the mesh says so (RR_INDEX_TYPE_UINT16).
__device__ __forceinline__ void tri_indices(const MeshDesc& m, size_t prim, uint32_t& i0, uint32_t& i1, uint32_t& i2)
{
    if (m.index16)
    {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(m.indices) + 3 * prim;
        i0 = p[0]; i1 = p[1]; i2 = p[2];
    }
    else
    {
        const uint32_t* p = m.indices + 3 * prim;
        i0 = p[0]; i1 = p[1]; i2 = p[2];
    }
}
// A vertex as one 8-byte and one 4-byte load: its 12 bytes start at a multiple of 4, so either half may be the aligned
// one (base8: the vertex buffer itself is 8-byte aligned).  Two L1 requests instead of three; pays where lanes gather
// scattered vertices (k_emit_leaves), not in the sweeps over triangles in input order (measured 30 % slower there).
__device__ __forceinline__ float3 ld_vertex(const float* verts, size_t off, bool base8)
{
    if (!base8) return ld3(verts + off);
    const bool   even = (off & 1) == 0;
    const float2 w = *reinterpret_cast<const float2*>(verts + off + (even ? 0 : 1));
    const float  u = verts[off + (even ? 2 : 0)];
    return even ? make_float3(w.x, w.y, u) : make_float3(u, w.x, w.y);
}
__device__ __forceinline__ float3 min3(float3 a, float3 b) { return make_float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__device__ __forceinline__ float3 max3(float3 a, float3 b) { return make_float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
__device__ __forceinline__ float3 xyz(float4 q) { return make_float3(q.x, q.y, q.z); }
__device__ __forceinline__ float4 pack(float3 v, uint32_t w) { return make_float4(v.x, v.y, v.z, __uint_as_float(w)); }
__device__ __forceinline__ uint32_t wbits(float4 q) { return __float_as_uint(q.w); }

// common.h:251-268
__device__ __forceinline__ uint32_t expand_bits(uint32_t r)
{
    r = (r * 0x00010001u) & 0xFF0000FFu;
    r = (r * 0x00000101u) & 0x0F00F00Fu;
    r = (r * 0x00000011u) & 0xC30C30C3u;
    r = (r * 0x00000005u) & 0x49249249u;
    return r;
}
// lbvh_calc_morton_codes_mesh.comp:114-125 + common.h:261-268.  NaN (zero extent) clamps to 0.
__device__ __forceinline__ uint32_t morton_of_box(float3 bmin, float3 bmax, float3 smin, float3 smax)
{
    const float3 ext = make_float3(smax.x - smin.x, smax.y - smin.y, smax.z - smin.z);
    const float3 c   = make_float3(0.5f * (bmin.x + bmax.x), 0.5f * (bmin.y + bmax.y), 0.5f * (bmin.z + bmax.z));
    const float  px = (c.x - smin.x) / ext.x, py = (c.y - smin.y) / ext.y, pz = (c.z - smin.z) / ext.z;
    const float  x = fminf(fmaxf(px * 1024.0f, 0.0f), 1023.0f);
    const float  y = fminf(fmaxf(py * 1024.0f, 0.0f), 1023.0f);
    const float  z = fminf(fmaxf(pz * 1024.0f, 0.0f), 1023.0f);
    return (expand_bits((uint32_t)x) << 2) | (expand_bits((uint32_t)y) << 1) | expand_bits((uint32_t)z);
}

struct OrderedBox
{
    uint32_t lo[3], hi[3];
    __device__ void init()
    {
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = 0xFFFFFFFFu; hi[a] = 0u; }
    }
    __device__ void grow(float3 p)
    {
        const uint32_t e[3] = {float_to_ordered(p.x), float_to_ordered(p.y), float_to_ordered(p.z)};
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], e[a]); hi[a] = max(hi[a], e[a]); }
    }
};

// CTA-wide reduction of an OrderedBox followed by 6 global atomics (lbvh_calc_mesh_aabb.comp:163-177).
// g_aabb layout: [0..2] = min xyz, [4..6] = max xyz (ordered-uint encoding), like the reference's uint[8].
__device__ __forceinline__ void reduce_box_to_global(OrderedBox b, uint32_t* g_aabb)
{
    __shared__ uint32_t s_red[6][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        b.lo[a] = __reduce_min_sync(0xffffffffu, b.lo[a]);
        b.hi[a] = __reduce_max_sync(0xffffffffu, b.hi[a]);
    }
    if (lane == 0)
    {
#pragma unroll
        for (int a = 0; a < 3; ++a) { s_red[a][warp] = b.lo[a]; s_red[3 + a][warp] = b.hi[a]; }
    }
    __syncthreads();
    if (warp == 0)
    {
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            uint32_t lo = lane < nwarps ? s_red[a][lane] : 0xFFFFFFFFu;
            uint32_t hi = lane < nwarps ? s_red[3 + a][lane] : 0u;
            lo          = __reduce_min_sync(0xffffffffu, lo);
            hi          = __reduce_max_sync(0xffffffffu, hi);
            if (lane == 0)
            {
                atomicMin(&g_aabb[a], lo);
                atomicMax(&g_aabb[4 + a], hi);
            }
        }
    }
}

// ---- K1: scene AABB over all referenced vertices (lbvh_calc_mesh_aabb.comp:121-178) -----------------
__global__ void __launch_bounds__(256) k_scene_aabb(MeshDesc m, uint32_t* __restrict__ g_aabb)
{
    OrderedBox box;
    box.init();
    const uint32_t stride = gridDim.x * blockDim.x;
    // two triangles per iteration: twice the loads in flight per thread (the sweep is latency-, not bandwidth-bound)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m.triangle_count; i += 2 * stride)
    {
        const uint32_t k  = min(i + stride, m.triangle_count - 1);  // second triangle (clamped: min/max are idempotent)
        uint32_t i0, i1, i2, k0, k1, k2;
        tri_indices(m, i, i0, i1, i2);
        tri_indices(m, k, k0, k1, k2);
        const float3   a0 = ld3(m.vertices + (size_t)i0 * m.stride_floats), a1 = ld3(m.vertices + (size_t)i1 * m.stride_floats),
                       a2 = ld3(m.vertices + (size_t)i2 * m.stride_floats);
        const float3   b0 = ld3(m.vertices + (size_t)k0 * m.stride_floats), b1 = ld3(m.vertices + (size_t)k1 * m.stride_floats),
                       b2 = ld3(m.vertices + (size_t)k2 * m.stride_floats);
        box.grow(a0); box.grow(a1); box.grow(a2);
        box.grow(b0); box.grow(b1); box.grow(b2);
    }
    reduce_box_to_global(box, g_aabb);
}

__device__ __forceinline__ void load_scene_box(const uint32_t* g_aabb, float3& smin, float3& smax)
{
    smin = make_float3(ordered_to_float(g_aabb[0]), ordered_to_float(g_aabb[1]), ordered_to_float(g_aabb[2]));
    smax = make_float3(ordered_to_float(g_aabb[4]), ordered_to_float(g_aabb[5]), ordered_to_float(g_aabb[6]));
}

// ---- K2: Morton codes + the 4x256 digit histograms the onesweep passes need --------------------------
// kScene=false: triangles (lbvh_calc_morton_codes_mesh.comp:78-130); kScene=true: instance world boxes
// (lbvh_calc_morton_codes_scene.comp:84-112).  ref[i]=i is implicit (identity values in sort pass 0).
template <bool kScene>
__global__ void __launch_bounds__(256)
    k_morton(MeshDesc m, const float4* __restrict__ boxes, uint32_t count, const uint32_t* __restrict__ g_aabb,
             uint32_t* __restrict__ codes, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_hist[4 * 256];
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    float3 smin, smax;
    load_scene_box(g_aabb, smin, smax);
    const uint32_t stride = gridDim.x * blockDim.x;
    // (two primitives per iteration, as in k_scene_aabb, measured slower here: 371 against 282 us at 50 M triangles)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        float3 bmin, bmax;
        if (kScene)
        {
            bmin = xyz(boxes[2 * (size_t)i]);
            bmax = xyz(boxes[2 * (size_t)i + 1]);
        }
        else
        {
            uint32_t i0, i1, i2;
            tri_indices(m, i, i0, i1, i2);
            const float3   v0 = ld3(m.vertices + (size_t)i0 * m.stride_floats);
            const float3   v1 = ld3(m.vertices + (size_t)i1 * m.stride_floats);
            const float3   v2 = ld3(m.vertices + (size_t)i2 * m.stride_floats);
            bmin              = min3(min3(v0, v1), v2);
            bmax              = max3(max3(v0, v1), v2);
        }
        const uint32_t code = morton_of_box(bmin, bmax, smin, smax);
        codes[i]            = code;
#pragma unroll
        for (int p = 0; p < 4; ++p) atomicAdd(&s_hist[p * 256 + ((code >> (8 * p)) & 255u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// ---- K4+K5 fused: hierarchy emission + bottom-up fit --------------------------------------------------
// delta(a, a+1) of lbvh_emit_hierarchy_mesh.comp:85-103 up to a strictly monotone map (true clz instead
// of the shader's 32-findMSB; only comparisons between deltas are ever used).
__device__ __forceinline__ int delta_adjacent(const uint32_t* __restrict__ codes, uint32_t a)
{
    const uint32_t x = codes[a] ^ codes[a + 1];
    return x ? __clz(x) : 32 + __clz(a ^ (a + 1));
}

struct EmitParams
{
    const uint32_t* codes;   // sorted
    const uint32_t* refs;    // sorted primitive / instance ids
    uint32_t        n;
    Node*           nodes;
    uint32_t*       lists;   // per 512-leaf window: count + up to kListSlots left ends of the subtrees k_emit_window left over
    uint8_t*        deltas;  // [n] delta(j, j+1) per sorted leaf in the geometry buffer's tail: written by a build, read by a refit
    uint32_t*       karras;      // tail header word: 1 while the tree is Karras-numbered (cleared by the treelet restructuring)
    bool            from_tail;   // refit: codes == nullptr, deltas / refs come from the tail; run only if *karras == 1
    uint32_t*       error;     // DeviceInfo::error_word (kErrorEmitListOverflow)
    bool            static_order;  // quality builds: the traversal orders children by the table in the update word (node_update_word)
    bool            prefetch;  // k_emit_window pulls the node images of its first pass into L2 before it merges
    bool            tma;     // k_emit_leaves stages node images in shared memory and stores them with TMA tensor copies
    uint32_t*       masks;             // [ceil(n/32)] per 32 leaves: left ends of the subtrees k_emit_leaves left over
    // mesh leaves
    MeshDesc mesh;
    // scene leaves
    const float4*       boxes;    // [2n] world boxes per instance
    const InstanceDesc* descs;    // [n]
    InstanceRecord*     records;  // [n]
    float4*             fwd;      // [3n] forward transforms (kept for layout parity with the reference's transforms[2i+1])
};

__device__ __forceinline__ void node_box(const float4 q0, const float4 q1, const float4 q2, const float4 q3, bool tri_leaf,
                                         float3& lo, float3& hi)
{
    if (tri_leaf)
    {   // calculate_aabb_for_triangle, common.h:241-248
        lo = min3(min3(xyz(q0), xyz(q1)), xyz(q2));
        hi = max3(max3(xyz(q0), xyz(q1)), xyz(q2));
    }
    else
    {   // union of the two stored child boxes, lbvh_fit_aabb_mesh.comp:98-110
        lo = min3(xyz(q0), xyz(q2));
        hi = max3(xyz(q1), xyz(q3));
    }
}

// Affine inverse with a fixed expression tree (the oracle evaluates the same one; GLSL inverse(mat4),
// common.h:326-342, is implementation defined).
__device__ __forceinline__ void affine_inverse(const float* m, float4& r0, float4& r1, float4& r2)
{
    const float a00 = m[0], a01 = m[1], a02 = m[2], tx = m[3];
    const float a10 = m[4], a11 = m[5], a12 = m[6], ty = m[7];
    const float a20 = m[8], a21 = m[9], a22 = m[10], tz = m[11];
    const float c00 = a11 * a22 - a12 * a21, c01 = a02 * a21 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const float c10 = a12 * a20 - a10 * a22, c11 = a00 * a22 - a02 * a20, c12 = a02 * a10 - a00 * a12;
    const float c20 = a10 * a21 - a11 * a20, c21 = a01 * a20 - a00 * a21, c22 = a00 * a11 - a01 * a10;
    const float det = (a00 * c00 + a01 * c10) + a02 * c20;
    const float inv = 1.0f / det;
    const float r00 = c00 * inv, r01 = c01 * inv, r02 = c02 * inv;
    const float r10 = c10 * inv, r11 = c11 * inv, r12 = c12 * inv;
    const float r20 = c20 * inv, r21 = c21 * inv, r22 = c22 * inv;
    r0 = make_float4(r00, r01, r02, -((r00 * tx + r01 * ty) + r02 * tz));
    r1 = make_float4(r10, r11, r12, -((r10 * tx + r11 * ty) + r12 * tz));
    r2 = make_float4(r20, r21, r22, -((r20 * tx + r21 * ty) + r22 * tz));
}

// Hierarchy emission + box fitting fused; every node's 64 bytes are written once (plus its parent word when the parent
// is formed by a later step).
//
// Node numbering (Karras 2012, lbvh_emit_hierarchy_mesh.comp:105-217): internal node i covers a range of sorted leaves
// that has i at one end; a non-root internal node sits at the RIGHT end of its range if it is a left child (it is its
// parent's split position) and at the LEFT end if it is a right child (split+1); the root is 0; leaf j is node N-1+j.
// The hierarchy over any run of consecutive, already finished subtrees ("elements") is the Cartesian tree of the deltas
// between neighbours: the node split right of element s spans the elements between the nearest smaller delta on either
// side -- the range FindSpan/FindSplit find by binary search from the node's own index.  Every subtree that lies inside
// a run of leaves also has its internal node indices inside that run.
//
// B200 mapping (no rendezvous words, atomics or fences anywhere; no shared-memory node images):
//   k_emit_leaves -- one warp per 32 consecutive sorted leaves.  group_merge() forms, in closed form and with ballots and
//     shuffles only, EVERY node whose range lies inside the 32 leaves (~85 % of all internal nodes): range, Karras index,
//     child ids, parent index, and both child boxes as range min/max over the lanes.  Leaves and nodes go straight to HBM
//     (64 B per lane, full sectors); what cannot be decided inside the group -- its maximal subtrees -- is recorded as one
//     32-bit mask of left ends.  Warps are persistent and software-pipeline the gather (ref -> indices -> vertices).
//   k_emit_window -- one warp per 512-leaf window repeats group_merge() over the left-over subtrees of its 16 groups
//     (boxes read back from the node images, L2), 32 at a time, until one group holds them all (2-3 passes); what is
//     left (~10 subtrees, at most 124: two per tree level) goes to a per-window list.
//   k_emit_upper  -- one CTA per 32 windows of the level below, the groups of a pass spread over its warps; launched
//     level after level (512 -> 16 Ki -> 512 Ki -> ... leaves per window) until one window is the whole tree and the root
//     is formed.
// profiles/round1_summary.md lists the variants measured before this one (single kernel with global atomics; 512-leaf
// CTA windows with shared-memory images and an atomic climb; barrier-paced rounds; a final global atomic climb; ...).
constexpr int      kEmitWindow    = 512;  // leaves per k_emit_window warp
constexpr int      kEmitMaxPasses = 1100;  // a run of 512 elements needs at most 511 merges, one every other pass
// What a contiguous run of leaves cannot merge internally are its maximal subtrees: at most two per level of the tree
// (one on each border path), and a radix tree over 30-bit codes with 32-bit index tie-breaks is at most 62 levels deep.
constexpr int      kListSlots  = 128;
constexpr int      kListStride = kListSlots + 1;  // count, then the left ends (sorted leaf indices), ascending

// A whole 64-byte node as two 32-byte stores (STG.E.256): every lane writes full sectors.
__device__ __forceinline__ void st_node(Node* dst, float4 q0, float4 q1, float4 q2, float4 q3)
{
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(q0.x), "f"(q0.y), "f"(q0.z), "f"(q0.w), "f"(q1.x),
                 "f"(q1.y), "f"(q1.z), "f"(q1.w)
                 : "memory");
    asm volatile("st.global.v8.f32 [%0+32], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(q2.x), "f"(q2.y), "f"(q2.z), "f"(q2.w),
                 "f"(q3.x), "f"(q3.y), "f"(q3.z), "f"(q3.w)
                 : "memory");
}

__device__ __forceinline__ void st_node_direct(Node* dst, float4 q0, float4 q1, float4 q2, float4 q3) { st_node(dst, q0, q1, q2, q3); }

// Range min (kMin) / max over lanes [Ll, lane] -> outL and [lane+1, pr] -> outR of a per-lane value: four doubling
// steps build the windows "2^k lanes ending here", the two ranges are then covered by the binary digits of their lengths.
template <bool kMin>
__device__ __forceinline__ void lane_range_reduce(float x, int lane, int lenL, int pr, int lenR, float& outL, float& outR)
{
    const uint32_t full = 0xffffffffu;
    auto op = [](float p, float q) { return kMin ? fminf(p, q) : fmaxf(p, q); };
    float t[5];
    t[0] = x;
#pragma unroll
    for (int k = 1; k < 5; ++k) t[k] = op(t[k - 1], __shfl_up_sync(full, t[k - 1], 1 << (k - 1)));
    float accL = kMin ? __int_as_float(0x7f800000) : __int_as_float(0xff800000), accR = accL;
    int   pL = lane, pR = pr;
#pragma unroll
    for (int k = 0; k < 5; ++k)
    {
        const float u = __shfl_sync(full, t[k], pL & 31);
        const float v = __shfl_sync(full, t[k], pR & 31);
        if (lenL & (1 << k)) { accL = op(accL, u); pL -= 1 << k; }
        if (lenR & (1 << k)) { accR = op(accR, v); pR -= 1 << k; }
    }
    outL = accL;
    outR = accR;
}

struct MergeOut
{
    bool     e_over;      // this lane's element is still a top-level subtree (its parent is not formed here, it is not the root)
    uint32_t e_parent;    // its parent's index when that was formed here, else INVALID
    bool     n_over;      // the node this lane formed is a top-level subtree ...
    int      n_a;         // ... starting at this sorted leaf,
    int      n_lane;      // which is the first leaf of this lane's element
    bool     any_formed;  // warp-uniform
};

// One group of up to 32 consecutive top-level subtrees ("elements": sorted leaves [a,b], node id, box), one per lane;
// D = delta(b, b+1), DL0 = delta left of the first element.  Forms every node whose range lies inside the group and
// writes its 64 bytes (parent word included when the parent is formed here as well, else INVALID).
// kLeaves: elements are the single leaves a = b = a0 + lane, id = leaf0 + a (saves four shuffles).  store_node(idx, q0..q3)
// writes a formed node.
template <bool kLeaves, class StoreNode>
__device__ __forceinline__ MergeOut group_merge(StoreNode&& store_node, int n, int cnt, bool valid, int a, int b, uint32_t id, float3 lo,
                                                float3 hi, int D, int DL0, uint32_t order_flag)
{
    const uint32_t full  = 0xffffffffu;
    const int      lane  = threadIdx.x & 31;
    const uint32_t vmask = cnt >= 32 ? full : ((1u << cnt) - 1u);
    const uint32_t below = (1u << lane) - 1u, above = ~((2u << lane) - 1u);
    if (!valid) D = 0;
    // lanes whose delta is smaller than mine (deltas fit 7 bits -- 0..63 for 30-bit codes, 0..95 for the 63-bit extension:
    // radix compare over bit planes)
    uint32_t lt = 0, eq = full;
#pragma unroll
    for (int bit = 6; bit >= 0; --bit)
    {
        const bool     mine = (D >> bit) & 1;
        const uint32_t B    = __ballot_sync(full, mine);
        if (mine) { lt |= eq & ~B; eq &= B; }
        else eq &= ~B;
    }
    lt &= vmask;
    const uint32_t lmask = lt & below, rmask = lt & above;
    const int      pl = lmask ? 31 - __clz(lmask) : -1;          // nearest smaller delta to the left (-1: the group's left border)
    const int      pr = rmask ? __ffs(rmask) - 1 : lane;         // ... to the right
    const bool     formed = lane + 1 < cnt && (lmask != 0 || DL0 < D) && rmask != 0;
    const uint32_t IG = __ballot_sync(full, formed);
    const int      Ll = pl + 1;
    const int      aL = kLeaves ? a - lane + Ll : __shfl_sync(full, a, Ll & 31);
    const int      bR = kLeaves ? a - lane + pr : __shfl_sync(full, b, pr);
    int            Dpl = __shfl_sync(full, D, pl < 0 ? 0 : pl);
    if (pl < 0) Dpl = DL0;
    const int      Dpr = __shfl_sync(full, D, pr);
    const bool     is_root = aL == 0 && bR == n - 1;
    const bool     nleft   = Dpr > Dpl;                           // left child iff delta(R,R+1) > delta(L-1,L)
    const uint32_t idx = is_root ? 0u : (uint32_t)(nleft ? bR : aL);
    const int      q   = nleft ? pr : pl;                         // split lane of the parent
    const bool     parent_in = formed && !is_root && q >= 0 && ((IG >> (q & 31)) & 1u);
    const uint32_t pidx = __shfl_sync(full, idx, parent_in ? q : lane);
    const uint32_t nid  = kLeaves ? id + 1u : __shfl_down_sync(full, id, 1);
    const uint32_t c0 = (Ll == lane) ? id : (uint32_t)b;          // Karras: left child = split, right child = split + 1
    const uint32_t c1 = (pr == lane + 1) ? nid : (uint32_t)(b + 1);
    float3 loL, hiL, loR, hiR;
    {
        const int lenL = lane - Ll + 1, lenR = pr - lane;
        lane_range_reduce<true>(lo.x, lane, lenL, pr, lenR, loL.x, loR.x);
        lane_range_reduce<true>(lo.y, lane, lenL, pr, lenR, loL.y, loR.y);
        lane_range_reduce<true>(lo.z, lane, lenL, pr, lenR, loL.z, loR.z);
        lane_range_reduce<false>(hi.x, lane, lenL, pr, lenR, hiL.x, hiR.x);
        lane_range_reduce<false>(hi.y, lane, lenL, pr, lenR, hiL.y, hiR.y);
        lane_range_reduce<false>(hi.z, lane, lenL, pr, lenR, hiL.z, hiR.z);
    }
    if (formed)
        store_node(idx, pack(loL, c0), pack(hiL, c1), pack(loR, parent_in ? pidx : kInvalid), pack(hiR, node_update_word(c0, c1, (uint32_t)n - 1u, 0u, order_flag ? order_flag : node_order_bits(loL, hiL, loR, hiR))));  // (no table where the traversal votes)

    // the elements themselves: parent = the node split at the larger of the two neighbouring deltas
    int Dprev = __shfl_up_sync(full, D, 1);
    if (lane == 0) Dprev = DL0;
    const int      qe = D > Dprev ? lane : lane - 1;
    const bool     e_parent_in = valid && qe >= 0 && ((IG >> (qe & 31)) & 1u);
    const uint32_t pidx_e = __shfl_sync(full, idx, e_parent_in ? qe : lane);
    MergeOut o;
    o.e_parent   = e_parent_in ? pidx_e : kInvalid;
    o.e_over     = valid && !e_parent_in && !(a == 0 && b == n - 1);
    o.n_over     = formed && !parent_in && !is_root;
    o.n_a        = aL;
    o.n_lane     = Ll;
    o.any_formed = IG != 0;
    return o;
}

__device__ __forceinline__ int delta_of(uint32_t ca, uint32_t cb, int a)  // delta(a, a+1) from the two codes
{
    const uint32_t x = ca ^ cb;
    return x ? __clz(x) : 32 + __clz((uint32_t)(a ^ (a + 1)));
}

// ---- level 1: one warp per 32 sorted leaves ---------------------------------------------------------------------------
constexpr int kLeavesCtasPerSm = 3;  // 80 registers; 4 / 5 / 6 CTAs per SM measured slower (spills), profiles/round1_summary.md
// Mesh path: persistent warps with a three-deep software pipeline over their groups -- while group g is merged, the
// vertex loads of g+1, the index loads of g+2 and the ref/code loads of g+3 are in flight, so the dependent chain
// ref -> indices -> vertices (three DRAM latencies) never stalls the warp (it was 72 % of the stall samples).
template <bool kScene, bool kFromTail>  // kFromTail: refit of a Karras-numbered tree from the deltas / ids kept in the geometry buffer's tail
__global__ void __launch_bounds__(256, kLeavesCtasPerSm)
    k_emit_leaves(EmitParams p, const __grid_constant__ CUtensorMap tm_leaf, const __grid_constant__ CUtensorMap tm_node)
{
    const uint32_t full = 0xffffffffu;
    const int n = (int)p.n, leaf0 = n - 1, lane = threadIdx.x & 31;
    const int ngroups = (n + 31) >> 5, nwarps = gridDim.x * (blockDim.x >> 5);
    const int g0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    auto direct_store = [&](uint32_t idx, float4 q0, float4 q1, float4 q2, float4 q3) { st_node(p.nodes + idx, q0, q1, q2, q3); };
    if (kScene)
    {
        for (int g = g0; g < ngroups; g += nwarps)
        {
            const int  wb = g << 5, j = wb + lane, cnt = min(32, n - wb);
            const bool valid = j < n;
            // deltas (0 outside the array: smaller than every real delta, 30-bit codes have clz >= 2)
            const uint32_t c  = valid ? p.codes[j] : 0u;
            uint32_t       cn = __shfl_down_sync(full, c, 1);
            if (lane == 31 && j + 1 < n) cn = p.codes[j + 1];
            const int D = (valid && j + 1 < n) ? delta_of(c, cn, j) : 0;
            int       DL0 = (lane == 0 && wb > 0) ? delta_of(p.codes[wb - 1], c, wb - 1) : 0;
            DL0 = __shfl_sync(full, DL0, 0);
            float3 lo = make_float3(0.f, 0.f, 0.f), hi = lo;
            uint32_t ref = 0;
            if (valid)
            {   // lbvh_fit_aabb_scene.comp:113-130
                ref = p.refs[j];
                const float4 bmin = p.boxes[2 * (size_t)ref], bmax = p.boxes[2 * (size_t)ref + 1];
                lo = xyz(bmin);
                hi = xyz(bmax);
                const InstanceDesc d = p.descs[ref];
                InstanceRecord rec;
                affine_inverse(d.m, rec.inv0, rec.inv1, rec.inv2);
                rec.blas = d.blas;
                rec.pad  = 0;
                p.records[ref]             = rec;
                p.fwd[3 * (size_t)ref + 0] = make_float4(d.m[0], d.m[1], d.m[2], d.m[3]);
                p.fwd[3 * (size_t)ref + 1] = make_float4(d.m[4], d.m[5], d.m[6], d.m[7]);
                p.fwd[3 * (size_t)ref + 2] = make_float4(d.m[8], d.m[9], d.m[10], d.m[11]);
            }
            const MergeOut o = group_merge<true>(direct_store, n, cnt, valid, j, j, (uint32_t)(leaf0 + j), lo, hi, D, DL0, p.static_order ? 0u : kNodeVoteOrder);
            if (valid) st_node(p.nodes + leaf0 + j, pack(lo, kInvalid), pack(hi, ref), pack(lo, o.e_parent), pack(hi, 0u));
            const uint32_t mask = __ballot_sync(full, o.e_over) | __reduce_or_sync(full, o.n_over ? 1u << ((o.n_a - wb) & 31) : 0u);
            if (lane == 0) p.masks[g] = mask;
        }
        return;
    }
    if (g0 >= ngroups) return;
    constexpr bool from_tail = kFromTail;
    if (from_tail && *p.karras != 1u) return;
    if (!from_tail && g0 == 0 && lane == 0) *p.karras = 1u;  // a fresh build is Karras-numbered (the treelet pass clears this)
    const float*    verts   = p.mesh.vertices;
    const size_t    vstride = p.mesh.stride_floats;
    // loads of a group that may lie past the end are clamped to the last leaf (results unused)
    auto leaf_of = [&](int g) -> int { return min((g << 5) + lane, n - 1); };
    struct Codes { uint32_t c, edge; };  // own code; lane 0: code left of the group, lane 31: code right of it
    auto load_codes = [&](int g) -> Codes {
        const int wb = g << 5;
        Codes k;
        k.edge = 0u;
        if (from_tail)
        {   // refit: delta(j, j+1) itself comes from the tail of the geometry buffer
            k.c = p.deltas[leaf_of(g)];
            if (lane == 0 && wb > 0 && wb < n) k.edge = p.deltas[wb - 1];
            return k;
        }
        k.c = p.codes[leaf_of(g)];
        if (lane == 0 && wb > 0 && wb < n) k.edge = p.codes[wb - 1];
        if (lane == 31 && wb + 32 < n) k.edge = p.codes[wb + 32];
        return k;
    };
    struct Idx { uint32_t i0, i1, i2; };
    auto load_idx = [&](uint32_t ref) -> Idx {
        Idx i;
        tri_indices(p.mesh, ref, i.i0, i.i1, i.i2);
        return i;
    };
    struct Tri { float3 v0, v1, v2; };
    const bool base8 = (reinterpret_cast<uintptr_t>(verts) & 7u) == 0;
    auto load_tri = [&](const Idx& i) -> Tri {
        return Tri{ld_vertex(verts, (size_t)i.i0 * vstride, base8), ld_vertex(verts, (size_t)i.i1 * vstride, base8),
                   ld_vertex(verts, (size_t)i.i2 * vstride, base8)};
    };
    // Node images leave through the TMA when p.tma: a 64-byte-strided STG costs the L1 data pipe one wavefront per 32-byte
    // sector (106 of this kernel's 276 wavefronts per group).  Each warp stages its 32 leaves and the 32 internal slots of its
    // group as two dense 2 KB blocks in shared memory -- 16-byte quads XOR-swizzled exactly as CU_TENSOR_MAP_SWIZZLE_64B
    // expects (chunk ^= (row >> 1) & 3), which makes the per-lane 16-byte stores conflict free -- and one lane issues two
    // tensor copies.  Internal slots the group did not form are written with stale bytes: they belong to nodes that a later
    // kernel of the emission forms (and writes whole), never to another group; rows past the arrays are clipped by the maps.
    __shared__ __align__(1024) unsigned char s_stage[8][2][2048];
    unsigned char* st_leaf = s_stage[threadIdx.x >> 5][0];
    unsigned char* st_node = s_stage[threadIdx.x >> 5][1];
    const bool     tma     = p.tma;
    auto quad_at = [](unsigned char* base, int row, int k) -> float4* {
        return reinterpret_cast<float4*>(base + row * 64 + ((k ^ ((row >> 1) & 3)) << 4));
    };
    int  wb_now = 0;
    auto staged_store = [&](uint32_t idx, float4 q0, float4 q1, float4 q2, float4 q3) {
        if (!tma) { st_node_direct(p.nodes + idx, q0, q1, q2, q3); return; }
        const int row = (int)idx - wb_now;
        *quad_at(st_node, row, 0) = q0; *quad_at(st_node, row, 1) = q1; *quad_at(st_node, row, 2) = q2; *quad_at(st_node, row, 3) = q3;
    };
    // prologue
    uint32_t ref_c = p.refs[leaf_of(g0)];
    Codes    cod_c = load_codes(g0);
    uint32_t ref_1 = p.refs[leaf_of(g0 + nwarps)];
    Codes    cod_1 = load_codes(g0 + nwarps);
    uint32_t ref_2 = p.refs[leaf_of(g0 + 2 * nwarps)];
    Tri      tri_c = load_tri(load_idx(ref_c));
    Idx      idx_1 = load_idx(ref_1);
    for (int g = g0; g < ngroups; g += nwarps)
    {
        // prefetch: vertices of the next group, indices of the one after, ref / codes of the third
        const Tri      tri_1 = load_tri(idx_1);
        const Idx      idx_2 = load_idx(ref_2);
        const uint32_t ref_3 = p.refs[leaf_of(g + 3 * nwarps)];
        const Codes    cod_2 = load_codes(g + 2 * nwarps);

        const int  wb = g << 5, j = wb + lane, cnt = min(32, n - wb);
        const bool valid = j < n;
        // deltas (0 outside the array: smaller than every real delta, 30-bit codes have clz >= 2)
        uint32_t cn = __shfl_down_sync(full, cod_c.c, 1);
        if (lane == 31) cn = cod_c.edge;
        int D, DL0;
        if (from_tail)
        {
            D   = (valid && j + 1 < n) ? (int)cod_c.c : 0;
            DL0 = (lane == 0 && wb > 0) ? (int)cod_c.edge : 0;
        }
        else
        {
            D   = (valid && j + 1 < n) ? delta_of(cod_c.c, cn, j) : 0;
            DL0 = (lane == 0 && wb > 0) ? delta_of(cod_c.edge, cod_c.c, wb - 1) : 0;
            if (valid) p.deltas[j] = (uint8_t)D;  // with the sorted ids (already in the tail) all a refit needs to repeat this
        }
        DL0 = __shfl_sync(full, DL0, 0);
        // lbvh_fit_aabb_mesh.comp:139-163
        const float3 lo = min3(min3(tri_c.v0, tri_c.v1), tri_c.v2), hi = max3(max3(tri_c.v0, tri_c.v1), tri_c.v2);
        const float3 v2 = tri_c.v2;
        wb_now = wb;
        if (tma)
        {   // the staging blocks are free again once the previous group's copies have read them
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            *quad_at(st_leaf, lane, 0) = pack(tri_c.v0, kInvalid);
            *quad_at(st_leaf, lane, 1) = pack(tri_c.v1, ref_c);
        }
        else if (valid)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p.nodes + leaf0 + j), "f"(tri_c.v0.x),
                         "f"(tri_c.v0.y), "f"(tri_c.v0.z), "f"(__uint_as_float(kInvalid)), "f"(tri_c.v1.x), "f"(tri_c.v1.y), "f"(tri_c.v1.z),
                         "f"(__uint_as_float(ref_c))
                         : "memory");
        const MergeOut o = group_merge<true>(staged_store, n, cnt, valid, j, j, (uint32_t)(leaf0 + j), lo, hi, D, DL0, p.static_order ? 0u : kNodeVoteOrder);
        if (tma)
        {
            *quad_at(st_leaf, lane, 2) = pack(v2, o.e_parent);
            *quad_at(st_leaf, lane, 3) = make_float4(0.f, 0.f, 0.f, 0.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0)
            {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tm_leaf)),
                             "r"((uint32_t)__cvta_generic_to_shared(st_leaf)), "r"(0), "r"(wb)
                             : "memory");
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tm_node)),
                             "r"((uint32_t)__cvta_generic_to_shared(st_node)), "r"(0), "r"(wb)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        else if (valid)
            asm volatile("st.global.v8.f32 [%0+32], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p.nodes + leaf0 + j), "f"(v2.x), "f"(v2.y),
                         "f"(v2.z), "f"(__uint_as_float(o.e_parent)), "f"(0.f), "f"(0.f), "f"(0.f), "f"(0.f)
                         : "memory");
        const uint32_t mask = __ballot_sync(full, o.e_over) | __reduce_or_sync(full, o.n_over ? 1u << ((o.n_a - wb) & 31) : 0u);
        if (lane == 0) p.masks[g] = mask;
        // rotate the pipeline
        tri_c = tri_1; ref_c = ref_1; cod_c = cod_1;
        idx_1 = idx_2; ref_1 = ref_2; cod_1 = cod_2;
        ref_2 = ref_3;
    }
    if (tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the copies are complete before the warp retires
}

// Box of a finished node from its image in memory (L2).
template <bool kScene>
__device__ __forceinline__ void node_box_ldcg(const Node* nodes, uint32_t node, int leaf0, float3& lo, float3& hi)
{
    const float4* np = reinterpret_cast<const float4*>(nodes + node);
    const float4  q0 = __ldcg(np), q1 = __ldcg(np + 1), q2 = __ldcg(np + 2);
    if (node >= (uint32_t)leaf0)
    {
        if (kScene) { lo = xyz(q0); hi = xyz(q1); }
        else
        {
            lo = min3(min3(xyz(q0), xyz(q1)), xyz(q2));
            hi = max3(max3(xyz(q0), xyz(q1)), xyz(q2));
        }
    }
    else
    {
        const float4 q3 = __ldcg(np + 3);
        lo = min3(xyz(q0), xyz(q2));
        hi = max3(xyz(q1), xyz(q3));
    }
}

// ---- level 2: one warp per 512-leaf window, over what its 16 groups left over --------------------------------------------
constexpr int kWindowWarps = 8;  // per CTA
template <bool kScene>
__global__ void __launch_bounds__(32 * kWindowWarps, 5) k_emit_window(EmitParams p)
{
    __shared__ __align__(16) uint8_t s_delta[kWindowWarps][kEmitWindow + 16];  // delta(a, a+1) for a = b0-4 .. b0+515: byte a - (b0 - 4)
    __shared__ uint32_t s_mask[kWindowWarps][2][kEmitWindow / 32];
    __shared__ uint16_t s_list[kWindowWarps][kEmitWindow + 2];  // left ends (relative to b0) of the current elements, ascending
    const uint32_t full = 0xffffffffu;
    const int n = (int)p.n, leaf0 = n - 1, lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int w = blockIdx.x * kWindowWarps + wi;
    const int b0 = w * kEmitWindow;
    if (b0 >= n) return;
    if (p.from_tail && *p.karras != 1u) return;  // restructured tree: the generic refit kernels do the work
    const int cnt = min(kEmitWindow, n - b0), b1 = b0 + cnt - 1;
    uint8_t*  delta = s_delta[wi];
    if (p.deltas)
    {   // a mesh build has just written delta(j, j+1) of every leaf into the geometry buffer's tail (k_emit_leaves), a refit finds
        // them there: the window's 513 bytes come in as 130 aligned words (b0 is a multiple of 512), staged at the same alignment.
        // (Reading two codes per delta here was 19 % of the kernel's stall samples; delta(n-1, n) is stored as 0, the padding of the
        // tail array is never looked at: delta_adj is only asked for a <= b1.)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.deltas + b0) - 1;   // word holding delta(b0-4 .. b0-1)
        uint32_t*       dst = reinterpret_cast<uint32_t*>(delta);
        const int       words = (cnt + 4 + 3) / 4;                                     // bytes b0-4 .. b1
        for (int k = lane; k < words; k += 32) dst[k] = (b0 == 0 && k == 0) ? 0u : __ldg(src + k);
    }
    else
        for (int k = lane; k < cnt + 1; k += 32)
        {
            const int a = b0 - 1 + k;
            delta[k + 3] = (uint8_t)((a >= 0 && a + 1 < n) ? delta_of(p.codes[a], p.codes[a + 1], a) : 0);
        }
    auto delta_adj = [&](int a) -> int { return (int)delta[a + 4 - b0]; };   // a in [b0-1, b1]
    const int ngrp = (cnt + 31) >> 5;
    if (lane < kEmitWindow / 32) s_mask[wi][0][lane] = lane < ngrp ? p.masks[(b0 >> 5) + lane] : 0u;
    __syncwarp();

    int       cur = 0, M = 0;
    uint16_t* list = s_list[wi];
    // expands the current mask words into the ordered list of left ends; list[M] = one past the window
    auto load_masks = [&]() {
        uint32_t m    = lane < kEmitWindow / 32 ? s_mask[wi][cur][lane] : 0u;
        uint32_t incl = __popc(m);
#pragma unroll
        for (int d = 1; d < kEmitWindow / 32; d <<= 1)
        {
            const uint32_t t = __shfl_up_sync(full, incl, d);
            if (lane >= d) incl += t;
        }
        int k = (int)(incl - __popc(m));
        M     = (int)__shfl_sync(full, incl, kEmitWindow / 32 - 1);
        while (m)
        {
            list[k++] = (uint16_t)(32 * lane + __ffs(m) - 1);
            m &= m - 1;
        }
        if (lane == 0) list[M] = (uint16_t)cnt;
        __syncwarp();
    };
    auto left_end = [&](int r) -> int { return b0 + (int)list[min(r, M)]; };
    load_masks();
    if (p.prefetch)
    {   // the first pass reads back ~100 node images k_emit_leaves wrote (long out of L2 on a large mesh), group after group with a
        // dependent merge in between: ask for all of them now
        for (int e = lane; e < M; e += 32)
        {
            const int a = left_end(e), b = left_end(e + 1) - 1;
            const uint32_t id = a == b ? (uint32_t)(leaf0 + a) : (uint32_t)(delta_adj(b) > delta_adj(a - 1) ? b : a);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.nodes + id));
        }
    }
    auto direct_store = [&](uint32_t idx, float4 q0, float4 q1, float4 q2, float4 q3) { st_node(p.nodes + idx, q0, q1, q2, q3); };
    bool prev_formed = true;
    for (int pass = 1; M > 1 && pass < kEmitMaxPasses; ++pass)
    {
        // every other pass shifts the grouping by half a group, so that siblings on either side of a border meet
        const int off = (M > 32 && (pass & 1)) ? 16 : 0;
        const int ngroups = off ? 1 + (M - 16 + 31) / 32 : (M + 31) / 32;
        if (lane < kEmitWindow / 32) s_mask[wi][cur ^ 1][lane] = 0u;
        __syncwarp();
        bool any = false;
        for (int g = 0; g < ngroups; ++g)
        {
            const int  e0 = off ? (g == 0 ? 0 : 16 + (g - 1) * 32) : g * 32;
            const int  e1 = min(M, off && g == 0 ? 16 : e0 + 32);
            const int  e  = e0 + lane;
            const bool v  = e < e1;
            const int  a = left_end(e);                        // (b1 + 1 for lanes past the end)
            const int  b = v ? left_end(e + 1) - 1 : a;
            uint32_t   id = 0;
            float3     elo = make_float3(0.f, 0.f, 0.f), ehi = elo;
            if (v)
            {
                id = a == b ? (uint32_t)(leaf0 + a) : (uint32_t)(delta_adj(b) > delta_adj(a - 1) ? b : a);
                node_box_ldcg<kScene>(p.nodes, id, leaf0, elo, ehi);
            }
            const int a_first = __shfl_sync(full, a, 0);
            const MergeOut o = group_merge<false>(direct_store, n, e1 - e0, v, a, b, id, elo, ehi, v ? delta_adj(b) : 0, delta_adj(a_first - 1), p.static_order ? 0u : kNodeVoteOrder);
            if (v && o.e_parent != kInvalid) reinterpret_cast<uint32_t*>(p.nodes + id)[11] = o.e_parent;  // q2.w
            if (o.e_over) atomicOr(&s_mask[wi][cur ^ 1][(a - b0) >> 5], 1u << ((a - b0) & 31));
            if (o.n_over) atomicOr(&s_mask[wi][cur ^ 1][(o.n_a - b0) >> 5], 1u << ((o.n_a - b0) & 31));
            any |= o.any_formed;
        }
        __syncwarp();
        cur ^= 1;
        load_masks();
        if (ngroups == 1 || (!any && !prev_formed)) break;
        prev_formed = any;
    }
    if (n <= kEmitWindow) return;  // the whole tree was local

    // ---- what is left has its sibling outside the window: the next level (k_emit_upper) continues from the list of left ends
    uint32_t* out = p.lists + (size_t)w * kListStride;
    if (lane == 0)
    {
        out[0] = (uint32_t)min(M, kListSlots);
        if (M > kListSlots) atomicOr(p.error, kErrorEmitListOverflow);  // (at most 2 per tree level can be left: never seen)
    }
    for (int e = lane; e < min(M, kListSlots); e += 32) out[1 + e] = (uint32_t)left_end(e);
}

// ---- level 3 and above: one CTA per 16 windows of the level below, until one window is the whole tree ----------------------
// Same passes as k_emit_window, with the groups of a pass spread over the warps of the CTA (a barrier between passes) and
// the deltas taken from the sorted codes: a few hundred elements per CTA, a few thousand CTAs at the first of these levels
// for 50 M triangles, one CTA at the last, which forms the root.  No rendezvous words, atomics or fences anywhere in the
// emission: the build is deterministic down to the order of its memory writes within a node.
constexpr int kUpperFan      = 16;  // windows of the level below per CTA (<= 32: one warp scans their counts); 16 halves the
                                    // shared memory of a CTA against 32 (8 instead of 4 resident): 4.18 -> 4.12 ms at 50 M triangles
constexpr int kUpperWarps    = 8;
constexpr int kUpperMaxElems = kUpperFan * kListSlots;
struct UpperSmem
{
    uint32_t list[2][kUpperMaxElems + 1];          // left ends of the current elements, ascending; ping-pong
    uint32_t stage[kUpperMaxElems + 64];           // what each group of a pass leaves over, before compaction
    uint32_t group_off[kUpperMaxElems / 32 + 4];   // per group: count, then (after the scan) offset
    uint32_t child_off[33];
};

template <bool kScene>
__global__ void __launch_bounds__(32 * kUpperWarps) k_emit_upper(EmitParams p, const uint32_t* __restrict__ in_lists, uint32_t num_in,
                                                                 uint32_t leaves_per_in, uint32_t* __restrict__ out_lists)
{
    extern __shared__ __align__(16) unsigned char upper_smem_raw[];
    UpperSmem& S = *reinterpret_cast<UpperSmem*>(upper_smem_raw);
    const uint32_t full = 0xffffffffu;
    const int n = (int)p.n, leaf0 = n - 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t c0 = blockIdx.x * kUpperFan, nc = min((uint32_t)kUpperFan, num_in - c0);
    const int b1 = (int)min((uint64_t)n, (uint64_t)(c0 + nc) * leaves_per_in) - 1;  // last leaf of this window
    if (p.from_tail && *p.karras != 1u) return;  // (uniform over the grid)

    // concatenate the lists of the children
    if (warp == 0)
    {
        const uint32_t c = lane < (int)nc ? in_lists[(size_t)(c0 + lane) * kListStride] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t t = __shfl_up_sync(full, incl, d);
            if (lane >= d) incl += t;
        }
        S.child_off[lane] = incl - c;
        if (lane == 31) S.child_off[32] = incl;
    }
    __syncthreads();
    int M = (int)S.child_off[32], cur = 0;
    for (uint32_t c = warp; c < nc; c += kUpperWarps)
    {
        const uint32_t base = S.child_off[c], cn = S.child_off[c + 1] - base;
        for (uint32_t k = lane; k < cn; k += 32) S.list[0][base + k] = in_lists[(size_t)(c0 + c) * kListStride + 1 + k];
    }
    __syncthreads();

    auto delta_at = [&](int a) -> int {
        return (a >= 0 && a + 1 < n) ? (p.deltas ? (int)p.deltas[a] : delta_of(p.codes[a], p.codes[a + 1], a)) : 0;
    };
    auto direct_store = [&](uint32_t idx, float4 q0, float4 q1, float4 q2, float4 q3) { st_node(p.nodes + idx, q0, q1, q2, q3); };
    bool prev_formed = true;
    for (int pass = 1; M > 1 && pass < kEmitMaxPasses; ++pass)
    {
        const int off = (M > 32 && (pass & 1)) ? 16 : 0;
        const int ngroups = off ? 1 + (M - 16 + 31) / 32 : (M + 31) / 32;
        bool any = false;
        for (int g = warp; g < ngroups; g += kUpperWarps)
        {
            const int  e0 = off ? (g == 0 ? 0 : 16 + (g - 1) * 32) : g * 32;
            const int  e1 = min(M, off && g == 0 ? 16 : e0 + 32);
            const int  e  = e0 + lane;
            const bool v  = e < e1;
            int        a = b1 + 1, b = b1 + 1, D = 0;
            uint32_t   id = 0;
            float3     elo = make_float3(0.f, 0.f, 0.f), ehi = elo;
            if (v)
            {
                a = (int)S.list[cur][e];
                b = e + 1 < M ? (int)S.list[cur][e + 1] - 1 : b1;
                D = delta_at(b);
                id = a == b ? (uint32_t)(leaf0 + a) : (uint32_t)(D > delta_at(a - 1) ? b : a);
                node_box_ldcg<kScene>(p.nodes, id, leaf0, elo, ehi);
            }
            const int a_first = __shfl_sync(full, a, 0);
            const MergeOut o = group_merge<false>(direct_store, n, e1 - e0, v, a, b, id, elo, ehi, D, delta_at(a_first - 1), p.static_order ? 0u : kNodeVoteOrder);
            if (v && o.e_parent != kInvalid) reinterpret_cast<uint32_t*>(p.nodes + id)[11] = o.e_parent;  // q2.w
            const uint32_t EL = __ballot_sync(full, o.e_over), NL = __ballot_sync(full, o.n_over);
            const uint32_t below = (1u << lane) - 1u;
            if (o.e_over) S.stage[32 * g + __popc(EL & below) + __popc(NL & below)] = (uint32_t)a;
            if (o.n_over)
            {
                const uint32_t bl = (1u << o.n_lane) - 1u;
                S.stage[32 * g + __popc(EL & bl) + __popc(NL & bl)] = (uint32_t)o.n_a;
            }
            if (lane == 0) S.group_off[g] = (uint32_t)(__popc(EL) + __popc(NL));
            any |= o.any_formed;
        }
        const bool formed = __syncthreads_or(any) != 0;
        // compaction: scan the group counts (warp 0), then every thread moves its slot
        if (warp == 0)
        {
            uint32_t run = 0;
            for (int g0 = 0; g0 < ngroups; g0 += 32)
            {
                const uint32_t c = g0 + lane < ngroups ? S.group_off[g0 + lane] : 0u;
                uint32_t incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1)
                {
                    const uint32_t t = __shfl_up_sync(full, incl, d);
                    if (lane >= d) incl += t;
                }
                if (g0 + lane < ngroups) S.group_off[g0 + lane] = run + incl - c;
                run += __shfl_sync(full, incl, 31);
            }
            if (lane == 0) S.group_off[ngroups] = run;
        }
        __syncthreads();
        for (int t = tid; t < 32 * ngroups; t += 32 * kUpperWarps)
        {
            const int g = t >> 5, r = t & 31;
            const uint32_t base = S.group_off[g];
            if ((uint32_t)r < S.group_off[g + 1] - base) S.list[cur ^ 1][base + r] = S.stage[t];
        }
        M = (int)S.group_off[ngroups];
        cur ^= 1;
        __syncthreads();
        if (ngroups == 1 || (!formed && !prev_formed)) break;
        prev_formed = formed;
    }
    uint32_t* out = out_lists + (size_t)blockIdx.x * kListStride;
    if (tid == 0) out[0] = (uint32_t)min(M, kListSlots);
    for (int e = tid; e < min(M, kListSlots); e += 32 * kUpperWarps) out[1 + e] = S.list[cur][e];
}

// ---- K6: refit / update (lbvh_fit_aabb_mesh.comp with UPDATE_KERNEL, vlk/update_hlbvh.cpp:118-185) ----
// Topology untouched (it may be a treelet-restructured tree, so nothing is assumed about node numbering).  The
// rendezvous counter is bit 0 of the node's own `update` word used as a parity bit (atomicXor; even = first arrival; the
// other bits are the builder's leaf flags, rr_internal.h node_update_word), so no
// reset pass is needed -- the reference's reset kernel covers only the first 1024 primitives.
//
// Three stages.  A thread that is still climbing after `max_levels` parents appends the node it just finished to a
// work list and retires; the next stage resumes from the list.  Stage 1 (one thread per leaf, 3 levels) does 7/8 of
// the node writes with short-lived CTAs, stage 2 (6 levels) most of the rest, stage 3 the O(n/512) long climbs to the
// root.  In a single kernel every CTA stayed resident until its one long climber reached the root and the SMs ran
// mostly empty (8.2 ms for 50 M triangles, 16 % of the HBM roofline).
// Finishes the refit of the ancestors of the complete subtree `me` on the spot, with the parity rendezvous of k_refit<false>
// (the node's `update` word counts arrivals; the second one writes the parent).  The staged refit kernels fall back to this when
// a hand-over list is full, so that a list sized for the common case can never lose a subtree (stale boxes would mean silently
// missed hits).
__device__ __forceinline__ void refit_climb_in_place(Node* __restrict__ nodes, int leaf0, uint32_t me)
{
    __threadfence();
    const float4* np = reinterpret_cast<const float4*>(nodes + me);
    const float4  q0 = __ldcg(np), q1 = __ldcg(np + 1), q2 = __ldcg(np + 2), q3 = __ldcg(np + 3);
    float3 lo, hi;
    node_box(q0, q1, q2, q3, me >= (uint32_t)leaf0, lo, hi);
    uint32_t parent = wbits(q2);
    while (parent != kInvalid)
    {
        __threadfence();
        const uint32_t old = atomicXor(reinterpret_cast<uint32_t*>(nodes + parent) + 15, 1u);
        if ((old & 1u) == 0) break;
        __threadfence();
        float4* pp = reinterpret_cast<float4*>(nodes + parent);
        const float4 p0 = __ldcg(pp), p1 = __ldcg(pp + 1), p2 = __ldcg(pp + 2);
        const uint32_t c0 = wbits(p0), c1 = wbits(p1), up = wbits(p2);
        const bool     is_left = (c0 == me);
        const uint32_t sib = is_left ? c1 : c0;
        const float4*  sp = reinterpret_cast<const float4*>(nodes + sib);
        const float4   s0 = __ldcg(sp), s1 = __ldcg(sp + 1), s2 = __ldcg(sp + 2), s3 = __ldcg(sp + 3);
        float3 slo, shi;
        node_box(s0, s1, s2, s3, sib >= (uint32_t)leaf0, slo, shi);
        // (the update word: parity as the atomicXor left it, leaf flags as they were, child order for the new boxes)
        if (is_left) { pp[0] = pack(lo, c0); pp[1] = pack(hi, c1); pp[2] = pack(slo, up); pp[3] = pack(shi, node_update_word(c0, c1, (uint32_t)leaf0, old ^ 1u, ((old & kNodeVoteOrder) ? kNodeVoteOrder : node_order_bits(lo, hi, slo, shi)))); }
        else { pp[0] = pack(slo, c0); pp[1] = pack(shi, c1); pp[2] = pack(lo, up); pp[3] = pack(hi, node_update_word(c0, c1, (uint32_t)leaf0, old ^ 1u, ((old & kNodeVoteOrder) ? kNodeVoteOrder : node_order_bits(slo, shi, lo, hi)))); }
        lo = min3(lo, slo);
        hi = max3(hi, shi);
        me = parent;
        parent = up;
    }
}

struct RefitLists
{
    uint32_t* count_a;   // stage 1 -> 2
    uint32_t* items_a;
    uint32_t* count_b;   // stage 2 -> 3
    uint32_t* items_b;
    uint32_t  capacity;  // of each list
};

template <bool kFromLeaves>
__global__ void __launch_bounds__(256)
    k_refit(MeshDesc m, Node* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ in_count, const uint32_t* __restrict__ in_items,
            uint32_t max_levels, uint32_t* __restrict__ out_count, uint32_t* __restrict__ out_items, uint32_t capacity,
            const uint32_t* __restrict__ karras)
{
    if (karras && *karras == 1u) return;  // Karras-numbered tree: the emission kernels have refitted it already
    __shared__ uint32_t s_list[256];
    __shared__ uint32_t s_n, s_base;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint32_t leaf0 = n - 1;
    const uint32_t total = kFromLeaves ? n : min(*in_count, capacity);
    // every thread runs the same number of rounds so that the hand-over barriers below are uniform
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x)
    {
        const uint32_t j = base + threadIdx.x;
        if (j < total)
        {
            uint32_t me, parent;
            float3   lo, hi;
            if (kFromLeaves)
            {
                me = leaf0 + j;
                float4* np = reinterpret_cast<float4*>(nodes + me);
                const float4 q1 = np[1], q2 = np[2];
                const uint32_t prim = wbits(q1);
                parent = wbits(q2);
                uint32_t i0, i1, i2;
                tri_indices(m, prim, i0, i1, i2);
                const float3 v0 = ld3(m.vertices + (size_t)i0 * m.stride_floats);
                const float3 v1 = ld3(m.vertices + (size_t)i1 * m.stride_floats);
                const float3 v2 = ld3(m.vertices + (size_t)i2 * m.stride_floats);
                np[0] = pack(v0, kInvalid);
                np[1] = pack(v1, prim);
                np[2] = pack(v2, parent);
                lo = min3(min3(v0, v1), v2);
                hi = max3(max3(v0, v1), v2);
            }
            else
            {
                me = in_items[j];
                const float4* np = reinterpret_cast<const float4*>(nodes + me);
                const float4 q0 = __ldcg(np), q1 = __ldcg(np + 1), q2 = __ldcg(np + 2), q3 = __ldcg(np + 3);
                node_box(q0, q1, q2, q3, me >= leaf0, lo, hi);  // k_refit_leaves may hand over a single leaf
                parent = wbits(q2);
            }
            uint32_t level = 0;
            while (parent != kInvalid)
            {
                if (level == max_levels)
                {   // hand over: `me` is complete, its rendezvous at `parent` is the next stage's first step
                    s_list[atomicAdd(&s_n, 1u)] = me;
                    break;
                }
                __threadfence();
                const uint32_t old = atomicXor(reinterpret_cast<uint32_t*>(nodes + parent) + 15, 1u);
                if ((old & 1u) == 0) break;
                __threadfence();
                float4* pp = reinterpret_cast<float4*>(nodes + parent);
                const float4 p0 = __ldcg(pp), p1 = __ldcg(pp + 1), p2 = __ldcg(pp + 2);
                const uint32_t c0 = wbits(p0), c1 = wbits(p1), up = wbits(p2);
                const bool     is_left = (c0 == me);
                const uint32_t sib     = is_left ? c1 : c0;
                const float4*  sp      = reinterpret_cast<const float4*>(nodes + sib);
                const float4   s0 = __ldcg(sp), s1 = __ldcg(sp + 1), s2 = __ldcg(sp + 2), s3 = __ldcg(sp + 3);
                float3 slo, shi;
                node_box(s0, s1, s2, s3, sib >= leaf0, slo, shi);
                if (is_left)
                {
                    pp[0] = pack(lo, c0); pp[1] = pack(hi, c1); pp[2] = pack(slo, up);
                    pp[3] = pack(shi, node_update_word(c0, c1, leaf0, old ^ 1u, ((old & kNodeVoteOrder) ? kNodeVoteOrder : node_order_bits(lo, hi, slo, shi))));
                }
                else
                {
                    pp[0] = pack(slo, c0); pp[1] = pack(shi, c1); pp[2] = pack(lo, up);
                    pp[3] = pack(hi, node_update_word(c0, c1, leaf0, old ^ 1u, ((old & kNodeVoteOrder) ? kNodeVoteOrder : node_order_bits(slo, shi, lo, hi))));
                }
                lo = min3(lo, slo);
                hi = max3(hi, shi);
                me = parent;
                parent = up;
                ++level;
            }
        }
        if (out_items)
        {
            __syncthreads();
            if (threadIdx.x == 0)
            {
                s_base = s_n ? atomicAdd(out_count, s_n) : 0u;
            }
            __syncthreads();
            if (threadIdx.x < s_n)
            {
                if (s_base + threadIdx.x < capacity) out_items[s_base + threadIdx.x] = s_list[threadIdx.x];
                else refit_climb_in_place(nodes, (int)leaf0, s_list[threadIdx.x]);  // full list: finish here, never drop
            }
            __syncthreads();
            if (threadIdx.x == 0) s_n = 0;
            __syncthreads();
        }
    }
}

constexpr int kRefitSlots = 16;  // words per 32-leaf group handed from k_refit_leaves to k_refit_window: count + up to 15 node ids

// Stage 1 of the staged refit, warp-cooperative: one persistent warp per 32 consecutive leaves (node order = sorted leaf
// order), gather software-pipelined like k_emit_leaves.  Subtrees of an LBVH cover contiguous leaf runs, so the sibling of
// a finished subtree is the neighbouring finished subtree of the warp whenever their parent lies inside the 32 leaves:
// the left lane takes the right lane's box by shuffles and writes the parent -- no atomic, no fence, no read-back of the
// sibling -- and keeps climbing, round by round; ~85 % of the internal nodes are refitted this way.  What is left (sibling
// outside the warp, or a treelet-restructured node whose children are not neighbours) is handed to stage 2, which
// continues with the parity rendezvous of k_refit.  The update word of a node fitted here keeps its parity: neither child
// announced itself.  (Re-deriving the hierarchy in closed form from deltas kept in the leaves, as the build does, was
// measured too: 3.8 ms against 3.1 ms at 50 M triangles -- the extra scattered word loads cost more than the rounds.)
__global__ void __launch_bounds__(256, 4)  // 64 registers; 3 / 5 / 6 CTAs per SM: 3.84 / 3.88 / 4.58 ms refit at 50 M triangles against 3.60
    k_refit_leaves(MeshDesc m, Node* __restrict__ nodes, uint32_t n, uint32_t* __restrict__ slots, uint32_t* __restrict__ out_count,
                   uint32_t* __restrict__ out_items, uint32_t capacity, const __grid_constant__ CUtensorMap tm_leaf, int tma,
                   const uint32_t* __restrict__ karras)
{
    if (karras && *karras == 1u) return;
    // leaves leave through the TMA like in k_emit_leaves (one swizzled 2 KB block per group) when `tma`
    __shared__ __align__(1024) unsigned char s_stage[8][2048];
    unsigned char* st_leaf = s_stage[threadIdx.x >> 5];
    auto quad_at = [](unsigned char* base, int row, int k) -> float4* {
        return reinterpret_cast<float4*>(base + row * 64 + ((k ^ ((row >> 1) & 3)) << 4));
    };
    const uint32_t full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int leaf0 = (int)n - 1, ngroups = ((int)n + 31) >> 5, nwarps = gridDim.x * (blockDim.x >> 5);
    const int g0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g0 >= ngroups) return;
    const float*    verts   = m.vertices;
    const size_t    vstride = m.stride_floats;
    const bool      base8   = (reinterpret_cast<uintptr_t>(verts) & 7u) == 0;
    auto leaf_of = [&](int g) -> int { return min((g << 5) + lane, (int)n - 1); };
    struct Words { uint32_t prim, parent; };
    auto load_words = [&](int g) -> Words {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(nodes + leaf0 + leaf_of(g));
        return Words{w[7], w[11]};  // q1.w, q2.w
    };
    struct Idx { uint32_t i0, i1, i2; };
    auto load_idx = [&](uint32_t prim) -> Idx {
        Idx i;
        tri_indices(m, prim, i.i0, i.i1, i.i2);
        return i;
    };
    struct Tri { float3 v0, v1, v2; };
    auto load_tri = [&](const Idx& i) -> Tri {
        return Tri{ld_vertex(verts, (size_t)i.i0 * vstride, base8), ld_vertex(verts, (size_t)i.i1 * vstride, base8),
                   ld_vertex(verts, (size_t)i.i2 * vstride, base8)};
    };
    // appends this warp's finished-but-waiting subtrees (up to two per lane) to the stage-2 list with one atomic; a full list
    // (adversarial topologies only) makes the lane finish the climb here with the parity rendezvous
    auto climb_here = [&](uint32_t me) { refit_climb_in_place(nodes, leaf0, me); };
    auto hand_over = [&](bool over_a, uint32_t id_a, bool over_b, uint32_t id_b) {
        const uint32_t ma = __ballot_sync(full, over_a), mb = __ballot_sync(full, over_b);
        if ((ma | mb) == 0) return;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(out_count, (uint32_t)(__popc(ma) + __popc(mb)));
        base = __shfl_sync(full, base, 0);
        const uint32_t lt = (1u << lane) - 1u;
        if (over_a)
        {
            const uint32_t slot = base + __popc(ma & lt);
            if (slot < capacity) out_items[slot] = id_a;
            else climb_here(id_a);
        }
        if (over_b)
        {
            const uint32_t slot = base + __popc(ma) + __popc(mb & lt);
            if (slot < capacity) out_items[slot] = id_b;
            else climb_here(id_b);
        }
    };
    Words w_c = load_words(g0), w_1 = load_words(g0 + nwarps), w_2 = load_words(g0 + 2 * nwarps);
    Tri   tri_c = load_tri(load_idx(w_c.prim));
    Idx   idx_1 = load_idx(w_1.prim);
    for (int g = g0; g < ngroups; g += nwarps)
    {
        const Tri   tri_1 = load_tri(idx_1);
        const Idx   idx_2 = load_idx(w_2.prim);
        const Words w_3   = load_words(g + 3 * nwarps);

        const int  j = (g << 5) + lane;
        const bool valid = j < (int)n;
        float3     lo = min3(min3(tri_c.v0, tri_c.v1), tri_c.v2), hi = max3(max3(tri_c.v0, tri_c.v1), tri_c.v2);
        if (tma)
        {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            *quad_at(st_leaf, lane, 0) = pack(tri_c.v0, kInvalid);
            *quad_at(st_leaf, lane, 1) = pack(tri_c.v1, w_c.prim);
            *quad_at(st_leaf, lane, 2) = pack(tri_c.v2, w_c.parent);
            *quad_at(st_leaf, lane, 3) = make_float4(0.f, 0.f, 0.f, 0.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0)
            {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tm_leaf)),
                             "r"((uint32_t)__cvta_generic_to_shared(st_leaf)), "r"(0), "r"(g << 5)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        else if (valid)
            st_node(nodes + leaf0 + j, pack(tri_c.v0, kInvalid), pack(tri_c.v1, w_c.prim), pack(tri_c.v2, w_c.parent),
                    make_float4(0.f, 0.f, 0.f, 0.f));
        {
            bool     owner = valid;
            uint32_t cur = (uint32_t)(leaf0 + j), par = w_c.parent;
            uint32_t own = __ballot_sync(full, owner);
            while (true)
            {
                const uint32_t above = own & ~((2u << lane) - 1u);
                const int      nxt   = above ? __ffs(above) - 1 : lane;
                const uint32_t npar  = __shfl_sync(full, par, nxt);
                const bool     can   = owner && above != 0 && npar == par && par != kInvalid;
                if (__ballot_sync(full, can) == 0) break;
                const int src = can ? nxt : lane;
                float3 nlo, nhi;
                nlo.x = __shfl_sync(full, lo.x, src); nlo.y = __shfl_sync(full, lo.y, src); nlo.z = __shfl_sync(full, lo.z, src);
                nhi.x = __shfl_sync(full, hi.x, src); nhi.y = __shfl_sync(full, hi.y, src); nhi.z = __shfl_sync(full, hi.z, src);
                if (can)
                {
                    float4 h0, h1, h2, h3;  // the parent's four words, fetched as its two 32-byte halves (L2)
                    asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=f"(h0.x), "=f"(h0.y), "=f"(h0.z), "=f"(h0.w), "=f"(h1.x), "=f"(h1.y), "=f"(h1.z), "=f"(h1.w)
                                 : "l"(nodes + par));
                    asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8+32];"
                                 : "=f"(h2.x), "=f"(h2.y), "=f"(h2.z), "=f"(h2.w), "=f"(h3.x), "=f"(h3.y), "=f"(h3.z), "=f"(h3.w)
                                 : "l"(nodes + par));
                    const uint32_t c0 = wbits(h0), c1 = wbits(h1), up = wbits(h2), upd = wbits(h3);
                    const bool     first = c0 == cur;  // which of the two is child0
                    st_node(nodes + par, pack(first ? lo : nlo, c0), pack(first ? hi : nhi, c1), pack(first ? nlo : lo, up),
                            pack(first ? nhi : hi, (upd & kNodeVoteOrder) ? upd : (upd & ~(0xFFu << kNodeOrderShift)) |
                                                       (first ? node_order_bits(lo, hi, nlo, nhi) : node_order_bits(nlo, nhi, lo, hi))));
                    lo = min3(lo, nlo);
                    hi = max3(hi, nhi);
                    cur = par;
                    par = up;
                }
                own &= ~__reduce_or_sync(full, can ? (1u << nxt) : 0u);
                owner = (own >> lane) & 1u;
            }
            // what is still waiting for a sibling goes, in leaf order, to this group's slots for k_refit_window (the root has no
            // parent and is done); more than kRefitSlots - 1 of them (never seen) overflow into the stage-3 list
            const bool     over = owner && par != kInvalid;
            const uint32_t om   = __ballot_sync(full, over);
            const int      rank = __popc(om & ((1u << lane) - 1u));
            uint32_t*      gs   = slots + (size_t)g * kRefitSlots;
            if (lane == 0) gs[0] = (uint32_t)min(__popc(om), kRefitSlots - 1);
            if (over && rank < kRefitSlots - 1) gs[1 + rank] = cur;
            hand_over(over && rank >= kRefitSlots - 1, cur, false, 0u);
        }
        tri_c = tri_1; w_c = w_1;
        idx_1 = idx_2; w_1 = w_2;
        w_2 = w_3;
    }
    if (tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Stage 2 of the staged refit for large meshes: one warp per 512-leaf window over what its 16 groups handed over (~100 finished
// subtrees, in leaf order).  Same sibling pairing by shuffles as k_refit_leaves, 32 subtrees at a time with their boxes read
// back from the node images, in passes (the grouping shifted by half a group every other pass) until nothing pairs any more;
// the rest (~2 % of the leaves: sibling outside the window) goes to the parity-rendezvous climb of k_refit<false>.
__global__ void __launch_bounds__(256)
    k_refit_window(Node* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ slots, uint32_t* __restrict__ out_count,
                   uint32_t* __restrict__ out_items, uint32_t capacity, const uint32_t* __restrict__ karras)
{
    if (karras && *karras == 1u) return;
    __shared__ uint32_t s_list[8][2][16 * (kRefitSlots - 1) + 16];
    const uint32_t full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int leaf0 = (int)n - 1, ngroups = ((int)n + 31) >> 5;
    const int w = blockIdx.x * 8 + wi, gbase = w * 16;
    if (gbase >= ngroups) return;
    const int ngrp = min(16, ngroups - gbase);
    // concatenate the slots of the groups
    int M;
    {
        const uint32_t c = lane < ngrp ? slots[(size_t)(gbase + lane) * kRefitSlots] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1)
        {
            const uint32_t t = __shfl_up_sync(full, incl, d);
            if (lane >= d) incl += t;
        }
        M = (int)__shfl_sync(full, incl, 15);
        const uint32_t* gs = slots + (size_t)(gbase + lane) * kRefitSlots + 1;
        for (uint32_t r = 0; r < c; ++r) s_list[wi][0][incl - c + r] = gs[r];
    }
    __syncwarp();
    int  cur = 0;
    bool prev_formed = true;
    for (int pass = 1; M > 1 && pass < 64; ++pass)
    {
        const int off = (M > 32 && (pass & 1)) ? 16 : 0;
        const int ngr = off ? 1 + (M - 16 + 31) / 32 : (M + 31) / 32;
        int  Mn = 0;
        bool any = false;
        for (int q = 0; q < ngr; ++q)
        {
            const int e0 = off ? (q == 0 ? 0 : 16 + (q - 1) * 32) : q * 32;
            const int e1 = min(M, off && q == 0 ? 16 : e0 + 32);
            const int e  = e0 + lane;
            bool      owner = e < e1;
            uint32_t  id = 0, par = kInvalid;
            float3    lo = make_float3(0.f, 0.f, 0.f), hi = lo;
            if (owner)
            {
                id = s_list[wi][cur][e];
                const float4* np = reinterpret_cast<const float4*>(nodes + id);
                const float4  q0 = __ldcg(np), q1 = __ldcg(np + 1), q2 = __ldcg(np + 2), q3 = __ldcg(np + 3);
                node_box(q0, q1, q2, q3, id >= (uint32_t)leaf0, lo, hi);
                par = wbits(q2);
            }
            uint32_t own = __ballot_sync(full, owner);
            while (true)
            {
                const uint32_t above = own & ~((2u << lane) - 1u);
                const int      nxt   = above ? __ffs(above) - 1 : lane;
                const uint32_t npar  = __shfl_sync(full, par, nxt);
                const bool     can   = owner && above != 0 && npar == par && par != kInvalid;
                if (__ballot_sync(full, can) == 0) break;
                any = true;
                const int src = can ? nxt : lane;
                float3 nlo, nhi;
                nlo.x = __shfl_sync(full, lo.x, src); nlo.y = __shfl_sync(full, lo.y, src); nlo.z = __shfl_sync(full, lo.z, src);
                nhi.x = __shfl_sync(full, hi.x, src); nhi.y = __shfl_sync(full, hi.y, src); nhi.z = __shfl_sync(full, hi.z, src);
                if (can)
                {
                    const uint32_t* pw = reinterpret_cast<const uint32_t*>(nodes + par);
                    const uint32_t  c0 = __ldcg(pw + 3), c1 = __ldcg(pw + 7), up = __ldcg(pw + 11), upd = __ldcg(pw + 15);
                    const bool      first = c0 == id;  // which of the two is child0
                    st_node(nodes + par, pack(first ? lo : nlo, c0), pack(first ? hi : nhi, c1), pack(first ? nlo : lo, up),
                            pack(first ? nhi : hi, (upd & kNodeVoteOrder) ? upd : (upd & ~(0xFFu << kNodeOrderShift)) |
                                                       (first ? node_order_bits(lo, hi, nlo, nhi) : node_order_bits(nlo, nhi, lo, hi))));
                    lo = min3(lo, nlo);
                    hi = max3(hi, nhi);
                    id = par;
                    par = up;
                }
                own &= ~__reduce_or_sync(full, can ? (1u << nxt) : 0u);
                owner = (own >> lane) & 1u;
            }
            const bool     over = owner && par != kInvalid;
            const uint32_t om   = __ballot_sync(full, over);
            if (over) s_list[wi][cur ^ 1][Mn + __popc(om & ((1u << lane) - 1u))] = id;
            Mn += __popc(om);
            __syncwarp();
        }
        M = Mn;
        cur ^= 1;
        if (ngr == 1 || (!any && !prev_formed)) break;
        prev_formed = any;
    }
    // the rest continues with the parity rendezvous (stage 3)
    uint32_t base = 0;
    if (lane == 0 && M > 0) base = atomicAdd(out_count, (uint32_t)M);
    base = __shfl_sync(full, base, 0);
    for (int e = lane; e < M; e += 32)
    {
        if (base + e < capacity) out_items[base + e] = s_list[wi][cur][e];
        else refit_climb_in_place(nodes, leaf0, s_list[wi][cur][e]);  // full list: finish here, never drop
    }
}

// ---- K8: instance world boxes + scene AABB (lbvh_calc_scene_aabb.comp:131-163, common.h:270-308) -------
__device__ __forceinline__ float3 transform_point(const float* m, float3 p)
{
    return make_float3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
                       ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}

__global__ void __launch_bounds__(256)
    k_instance_boxes(const InstanceDesc* __restrict__ descs, uint32_t n, int corner_quirk, float4* __restrict__ boxes,
                     uint32_t* __restrict__ g_aabb, SceneHeader* __restrict__ header, SceneHeader header_value)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *header = header_value;  // the scene buffer describes itself (rr_internal.h)
    OrderedBox sb;
    sb.init();
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const InstanceDesc d  = descs[i];
        const float4*      rp = reinterpret_cast<const float4*>(d.blas);
        const float4       q0 = rp[0], q1 = rp[1], q2 = rp[2], q3 = rp[3];
        float3 a, b;
        // A BLAS whose root is a leaf (one triangle) is bounded by its vertices.
        node_box(q0, q1, q2, q3, wbits(q0) == kInvalid, a, b);
        float3 c[8];
        c[0] = a;
        c[1] = make_float3(a.x, a.y, b.z);
        c[2] = make_float3(a.x, b.y, a.z);
        c[3] = make_float3(a.x, b.y, b.z);
        c[4] = make_float3(b.x, a.y, b.z);
        c[5] = make_float3(b.x, b.y, a.z);
        c[6] = corner_quirk ? b : make_float3(b.x, a.y, a.z);  // reference enumerates pmax twice (common.h:288-289)
        c[7] = b;
        float3 lo = transform_point(d.m, c[0]), hi = lo;
#pragma unroll
        for (int k = 1; k < 8; ++k)
        {
            const float3 t = transform_point(d.m, c[k]);
            lo = min3(lo, t);
            hi = max3(hi, t);
        }
        boxes[2 * (size_t)i]     = make_float4(lo.x, lo.y, lo.z, 0.f);
        boxes[2 * (size_t)i + 1] = make_float4(hi.x, hi.y, hi.z, 0.f);
        sb.grow(lo);
        sb.grow(hi);
    }
    reduce_box_to_global(sb, g_aabb);
}

// Tensor map over `rows` consecutive 64-byte nodes (16 floats per row, 32 rows per box, 64-byte swizzle): the destination of the
// TMA stores of k_emit_leaves.  cuTensorMapEncodeTiled comes from the runtime's driver entry point; false if unavailable.
static bool node_tensor_map(CUtensorMap* tm, Node* base, uint64_t rows)
{
    using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<EncodeFn>(f);
    }();
    if (!fn || rows == 0 || rows > 0xFFFFFFFFull || (reinterpret_cast<uintptr_t>(base) & 15u)) return false;
    const cuuint64_t dims[2]    = {16, rows};
    const cuuint64_t strides[1] = {sizeof(Node)};
    const cuuint32_t box[2]     = {16, 32};
    const cuuint32_t estr[2]    = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Words of the list area one level with `windows` windows needs.
inline size_t emit_list_words(uint32_t windows) { return align_up((size_t)windows * kListStride, 64); }

template <bool kScene>
int launch_emit_fit(const DeviceInfo& dev, cudaStream_t s, const EmitParams& p_in)
{
    const uint32_t groups = (p_in.n + 31) / 32, windows = (p_in.n + kEmitWindow - 1) / kEmitWindow;
    // persistent warps: exactly the resident CTAs, so that every warp streams through a long run of groups
    const uint32_t ctas = std::min<uint32_t>((groups + 7) / 8, (uint32_t)(dev.sm_count * kLeavesCtasPerSm));
    EmitParams  p = p_in;
    CUtensorMap tm_leaf{}, tm_node{};
    static const bool tma_allowed = [] { const char* e = std::getenv("RR_CUDA_EMIT_TMA"); return !e || std::atoi(e) != 0; }();
    static const bool prefetch_allowed = [] { const char* e = std::getenv("RR_CUDA_EMIT_PREFETCH"); return !e || std::atoi(e) != 0; }();
    p.prefetch = prefetch_allowed;
    p.error    = dev.error_word;
    p.tma = !kScene && tma_allowed && p.n >= 2 && node_tensor_map(&tm_leaf, p.nodes + (p.n - 1), p.n) && node_tensor_map(&tm_node, p.nodes, p.n - 1);
    if (p.from_tail) k_emit_leaves<kScene, true><<<ctas, 256, 0, s>>>(p, tm_leaf, tm_node);
    else k_emit_leaves<kScene, false><<<ctas, 256, 0, s>>>(p, tm_leaf, tm_node);
    if (p.n <= 32) return 1;  // the whole tree was inside one group
    k_emit_window<kScene><<<(windows + kWindowWarps - 1) / kWindowWarps, 32 * kWindowWarps, 0, s>>>(p);
    int launches = 2;
    // once per device and thread-safe: contexts on different devices / threads share this code (radeonrays.h:267-270)
    static std::once_flag attr_once[kMaxDevices];
    std::call_once(attr_once[dev.device % kMaxDevices], [] {
        RR_CUDA_CHECK(cudaFuncSetAttribute(k_emit_upper<kScene>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UpperSmem)));
    });
    // upper levels: kUpperFan windows of the level below per CTA, lists ping-pong between the two halves of the list area
    uint32_t  num = windows, leaves_per = (uint32_t)kEmitWindow;
    uint32_t* in  = p.lists;
    uint32_t* out = p.lists + emit_list_words(windows);
    while (num > 1)
    {
        const uint32_t upper_ctas = (num + kUpperFan - 1) / kUpperFan;
        k_emit_upper<kScene><<<upper_ctas, 32 * kUpperWarps, sizeof(UpperSmem), s>>>(p, in, num, leaves_per, out);
        ++launches;
        num = upper_ctas;
        leaves_per = (uint32_t)std::min<uint64_t>((uint64_t)leaves_per * kUpperFan, 0x80000000ull);
        std::swap(in, out);
    }
    return launches;
}

inline int grid_for(const DeviceInfo& dev, uint32_t n, int threads, int ctas_per_sm)
{
    size_t need = ((size_t)n + threads - 1) / threads;
    return (int)std::max<size_t>(1, std::min<size_t>(need, (size_t)dev.sm_count * ctas_per_sm));
}
}  // namespace

// ---- layouts -------------------------------------------------------------------------------------------
// Build scratch: [aabb_max 16 | sort hist/tickets/status]  <- one memset(0)
//                [aabb_min 16]                             <- one memset(0xFF)
//                codes 4N | sorted codes 4N | sorted refs 4N | sort tmp keys/vals 8N
// The scene-AABB words are split so that each memset node also initialises the min (all ones) / max (zero)
// identities of the ordered encoding; g_aabb[0..2]=min, [4..6]=max as in the reference's uint[8].
// List area of the emission levels: one list per 512-leaf window, and one per window of the next level (ping-pong).
static size_t emit_list_bytes(uint32_t n)
{
    const uint32_t windows = (n + kEmitWindow - 1) / kEmitWindow;
    return n <= (uint32_t)kEmitWindow ? 0 : sizeof(uint32_t) * (emit_list_words(windows) + emit_list_words((windows + kUpperFan - 1) / kUpperFan));
}

BlasLayout blas_layout(uint32_t n, bool restructure, bool morton63)
{
    BlasLayout L;
    L.n    = n;
    L.morton63 = morton63;
    L.sort = sort_layout(n);
    size_t off = 0;
    // scene AABB: min words preset to 0xFF.., max words to 0
    L.aabb_off  = off;              // 32 B: min[4] then max[4]
    off += 32;
    off += 256 - 32;  // the scene AABB keeps a 256-byte line of its own
    L.codes_off = off; off += align_up(sizeof(uint32_t) * (size_t)n, 256);
    L.sorted_codes_off = off; off += align_up(sizeof(uint32_t) * (size_t)n, 256);
    L.sorted_refs_off  = 0;  // the sorted primitive ids go straight into the geometry buffer's tail (tail_refs_off)
    L.sort_off = off; off += L.sort.total;
    L.lists_off = off; off += align_up(emit_list_bytes(n), 256);
    if (morton63)
    {   // codes_off holds the low words, sorted_codes_off is unused
        const size_t w = align_up(sizeof(uint32_t) * (size_t)n, 256);
        L.hi_off = off; off += w;
        L.sorted_lo_off = off; off += w;
        L.perm1_off = off; off += w;
        L.hi_gathered_off = off; off += w;
        L.sorted_hi_off = off; off += w;
        L.perm2_off = off; off += w;
        L.codes64_off = off; off += 2 * w;
    }
    L.treelet_off  = 0;
    L.treelet_size = restructure ? treelet_scratch_size(n) : 0;
    L.scratch_total = std::max(off, L.treelet_size);
    // geometry buffer: VkBvhNode[2N-1], then the private tail (header | deltas | sorted primitive ids) a refit re-emits from
    size_t roff = align_up(sizeof(Node) * (2 * (size_t)n - 1), 256);
    L.tail_off        = roff; roff += 256;
    L.tail_deltas_off = roff; roff += align_up((size_t)n, 256);
    L.tail_refs_off   = roff; roff += align_up(sizeof(uint32_t) * (size_t)n, 256);
    L.result_total    = roff;
    return L;
}

// ---- 63-bit Morton extension (BASELINE north_star "30/63-bit Morton codes"; RR_CUDA_OPTION_MORTON_BITS = 63) ----------------------
// The reference ships 30-bit codes only (its DX fallback keeps a compiled-out 64-bit delta(), build_hlbvh_fallback.hlsl:16,95-108),
// so this is an extension, defined in oracle/rr_oracle.c (rro_build_blas63) and bit-exact against it: the 30-bit pipeline with 21
// bits per axis.  Built from the pieces that exist: the code is kept as two 32-bit words, the stable 64-bit sort is two stable
// 32-bit onesweep sorts (low word, then high word of the low-sorted sequence), and the hierarchy is emitted by the refit's
// re-emission path, which needs only delta(j, j+1) and the sorted primitive ids in the geometry buffer's tail.
__device__ __forceinline__ uint64_t expand_bits21(uint32_t v)
{
    uint64_t x = v & 0x1FFFFFu;
    x = (x | (x << 32)) & 0x001F00000000FFFFull;
    x = (x | (x << 16)) & 0x001F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
__global__ void __launch_bounds__(256)
    k_morton63(MeshDesc m, const uint32_t* __restrict__ g_aabb, uint32_t* __restrict__ lo, uint32_t* __restrict__ hi)
{
    float3 smin, smax;
    load_scene_box(g_aabb, smin, smax);
    const float3   ext = make_float3(smax.x - smin.x, smax.y - smin.y, smax.z - smin.z);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m.triangle_count; i += stride)
    {
        uint32_t i0, i1, i2;
        tri_indices(m, i, i0, i1, i2);
        const float3 v0 = ld3(m.vertices + (size_t)i0 * m.stride_floats), v1 = ld3(m.vertices + (size_t)i1 * m.stride_floats),
                     v2 = ld3(m.vertices + (size_t)i2 * m.stride_floats);
        const float3 bmin = min3(min3(v0, v1), v2), bmax = max3(max3(v0, v1), v2);
        const float3 c  = make_float3(0.5f * (bmin.x + bmax.x), 0.5f * (bmin.y + bmax.y), 0.5f * (bmin.z + bmax.z));
        const float  px = (c.x - smin.x) / ext.x, py = (c.y - smin.y) / ext.y, pz = (c.z - smin.z) / ext.z;
        const float  x = fminf(fmaxf(px * 2097152.0f, 0.0f), 2097151.0f);
        const float  y = fminf(fmaxf(py * 2097152.0f, 0.0f), 2097151.0f);
        const float  z = fminf(fmaxf(pz * 2097152.0f, 0.0f), 2097151.0f);
        const uint64_t code = (expand_bits21((uint32_t)x) << 2) | (expand_bits21((uint32_t)y) << 1) | expand_bits21((uint32_t)z);
        lo[i] = (uint32_t)code;
        hi[i] = (uint32_t)(code >> 32);
    }
}
__global__ void __launch_bounds__(256) k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t* __restrict__ dst, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[perm[i]];
}
// Final order k: position perm2[k] of the low-sorted sequence, i.e. primitive perm1[perm2[k]]; writes the sorted 64-bit codes, the
// sorted primitive ids and delta(k, k+1) (dx/kernels/build_hlbvh_fallback.hlsl:95-108 with true clz) into the geometry buffer's tail.
__global__ void __launch_bounds__(256)
    k_finish63(uint32_t n, const uint32_t* __restrict__ perm1, const uint32_t* __restrict__ perm2, const uint32_t* __restrict__ sorted_lo,
               const uint32_t* __restrict__ sorted_hi, uint32_t* __restrict__ refs, uint8_t* __restrict__ deltas, uint64_t* __restrict__ codes64,
               uint32_t* __restrict__ karras)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *karras = 1u;  // the tree about to be emitted is the Karras tree of these deltas
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    {
        const uint32_t p2 = perm2[k];
        const uint64_t c  = ((uint64_t)sorted_hi[k] << 32) | sorted_lo[p2];
        refs[k]    = perm1[p2];
        codes64[k] = c;
        int d = 0;
        if (k + 1 < n)
        {
            const uint64_t cn = ((uint64_t)sorted_hi[k + 1] << 32) | sorted_lo[perm2[k + 1]];
            const uint64_t x  = c ^ cn;
            d = x ? __clzll((long long)x) : 64 + __clz((int)(k ^ (k + 1)));
        }
        deltas[k] = (uint8_t)d;
    }
}

// ---- multi-mesh geometries: concatenate on the device, then build / refit as one mesh -----------------------------------------
__global__ void __launch_bounds__(256) k_merge_meshes(MeshGroup g, float* __restrict__ out_v, uint32_t* __restrict__ out_i)
{
    const uint32_t tris = g.tri_first[g.count], verts = g.vert_first[g.count];
    const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t t = t0; t < tris; t += stride)
    {
        uint32_t k = 0;
        while (t >= g.tri_first[k + 1]) ++k;
        uint32_t i0, i1, i2;
        tri_indices(g.mesh[k], t - g.tri_first[k], i0, i1, i2);
        out_i[3 * (size_t)t]     = i0 + g.vert_first[k];
        out_i[3 * (size_t)t + 1] = i1 + g.vert_first[k];
        out_i[3 * (size_t)t + 2] = i2 + g.vert_first[k];
    }
    for (uint32_t v = t0; v < verts; v += stride)
    {
        uint32_t k = 0;
        while (v >= g.vert_first[k + 1]) ++k;
        const float3 p = ld3(g.mesh[k].vertices + (size_t)(v - g.vert_first[k]) * g.mesh[k].stride_floats);
        out_v[3 * (size_t)v] = p.x; out_v[3 * (size_t)v + 1] = p.y; out_v[3 * (size_t)v + 2] = p.z;
    }
}
size_t merge_scratch_size(const MeshGroup& g)
{
    return g.count <= 1 ? 0 : align_up(12 * (size_t)g.vertices(), 256) + align_up(12 * (size_t)g.triangles(), 256);
}
MeshDesc merge_meshes(const DeviceInfo& dev, cudaStream_t s, const MeshGroup& g, void* scratch)
{
    if (g.count <= 1) return g.mesh[0];
    float*    out_v = reinterpret_cast<float*>(scratch);
    uint32_t* out_i = reinterpret_cast<uint32_t*>(static_cast<char*>(scratch) + align_up(12 * (size_t)g.vertices(), 256));
    k_merge_meshes<<<grid_for(dev, std::max(g.triangles(), g.vertices()), 256, 8), 256, 0, s>>>(g, out_v, out_i);
    *dev.launches += 1;
    RR_CUDA_CHECK(cudaGetLastError());
    MeshDesc m{};
    m.vertices = out_v; m.vertex_count = g.vertices(); m.stride_floats = 3; m.indices = out_i; m.triangle_count = g.triangles(); m.index16 = 0;
    return m;
}

static void reset_build_scratch(cudaStream_t s, char* sc, size_t aabb_off, const SortLayout& sl, size_t sort_off)
{
    // min words -> 0xFF ; max words -> 0 ; sort bookkeeping -> 0
    RR_CUDA_CHECK(cudaMemsetAsync(sc + aabb_off, 0xFF, 16, s));
    RR_CUDA_CHECK(cudaMemsetAsync(sc + aabb_off + 16, 0x00, 16, s));
    RR_CUDA_CHECK(cudaMemsetAsync(sc + sort_off + sl.hist_off, 0, sl.tmp_keys_off - sl.hist_off, s));
}

void build_blas(const DeviceInfo& dev, cudaStream_t s, const MeshDesc& mesh, const BlasLayout& L, void* scratch, Node* nodes,
                bool restructure)
{
    const uint32_t n = mesh.triangle_count;
    if (n == 0) return;
    char*     sc           = (char*)scratch;
    uint32_t* g_aabb       = reinterpret_cast<uint32_t*>(sc + L.aabb_off);
    uint32_t* codes        = reinterpret_cast<uint32_t*>(sc + L.codes_off);
    uint32_t* sorted_codes = reinterpret_cast<uint32_t*>(sc + L.sorted_codes_off);
    // the last sort pass writes the sorted primitive ids where a refit will look for them: the geometry buffer's tail
    uint32_t* sorted_refs  = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(nodes) + L.tail_refs_off);
    void*     sort_scratch = sc + L.sort_off;

    reset_build_scratch(s, sc, L.aabb_off, L.sort, L.sort_off);
    k_scene_aabb<<<grid_for(dev, n, 256, 8), 256, 0, s>>>(mesh, g_aabb);
    EmitParams p{};
    if (L.morton63)
    {
        uint32_t* hi = reinterpret_cast<uint32_t*>(sc + L.hi_off), *sorted_lo = reinterpret_cast<uint32_t*>(sc + L.sorted_lo_off);
        uint32_t* perm1 = reinterpret_cast<uint32_t*>(sc + L.perm1_off), *hi_g = reinterpret_cast<uint32_t*>(sc + L.hi_gathered_off);
        uint32_t* sorted_hi = reinterpret_cast<uint32_t*>(sc + L.sorted_hi_off), *perm2 = reinterpret_cast<uint32_t*>(sc + L.perm2_off);
        char*     geom = reinterpret_cast<char*>(nodes);
        const int grid = grid_for(dev, n, 256, 8);
        k_morton63<<<grid, 256, 0, s>>>(mesh, g_aabb, codes, hi);
        sort_histogram(dev, s, L.sort, sort_scratch, codes);
        sort_pairs(dev, s, L.sort, sort_scratch, codes, nullptr, sorted_lo, perm1);            // stable by the low word
        k_gather_u32<<<grid, 256, 0, s>>>(hi, perm1, hi_g, n);
        sort_reset(dev, s, L.sort, sort_scratch);
        sort_histogram(dev, s, L.sort, sort_scratch, hi_g);
        sort_pairs(dev, s, L.sort, sort_scratch, hi_g, nullptr, sorted_hi, perm2);             // then stable by the high word
        k_finish63<<<grid, 256, 0, s>>>(n, perm1, perm2, sorted_lo, sorted_hi, sorted_refs, reinterpret_cast<uint8_t*>(geom + L.tail_deltas_off),
                                        reinterpret_cast<uint64_t*>(sc + L.codes64_off), reinterpret_cast<uint32_t*>(geom + L.tail_off));
        *dev.launches += 4;
        p.codes = nullptr; p.from_tail = true;   // emission from the tail's deltas and ids, as a refit does
    }
    else
    {
        k_morton<false><<<grid_for(dev, n, 256, 8), 256, 0, s>>>(mesh, nullptr, n, g_aabb, codes, sort_hist_ptr(L.sort, sort_scratch));
        *dev.launches += 1;
        sort_pairs(dev, s, L.sort, sort_scratch, codes, nullptr, sorted_codes, sorted_refs);
        p.codes = sorted_codes;
    }
    *dev.launches += 1;
    p.refs = sorted_refs; p.n = n; p.nodes = nodes; p.mesh = mesh;
    p.static_order = restructure && n >= 64;   // (restructure_blas leaves smaller trees alone, and so does a later refit's re-emission)
    p.lists = reinterpret_cast<uint32_t*>(sc + L.lists_off);
    p.masks = reinterpret_cast<uint32_t*>(sc + L.sort_off + L.sort.tmp_vals_off);  // the sort's ping-pong buffer is dead now
    char* geom  = reinterpret_cast<char*>(nodes);
    p.karras    = reinterpret_cast<uint32_t*>(geom + L.tail_off);
    p.deltas    = reinterpret_cast<uint8_t*>(geom + L.tail_deltas_off);
    *dev.launches += launch_emit_fit<false>(dev, s, p);
    RR_CUDA_CHECK(cudaGetLastError());
    if (restructure) restructure_blas(dev, s, nodes, n, scratch);  // (clears the tail's header word: no longer the Karras tree)
}

// Update scratch: [256 B: the two list counters | list A 4 x capacity | list B 4 x capacity], capacity = n/4 + 256:
// a hand-over after k levels implies a finished subtree of >= k+1 leaves, so stage 1 (3 levels) emits <= n/4 entries.
constexpr uint32_t kRefitWarpPathMin = 500000u;  // triangles from which the generic stages 1 / 2 are the warp-cooperative kernels
// Stage-3 list of the warp-cooperative path: a 32-leaf group hands at most kRefitSlots - 1 subtrees to its window, and a window
// passes on at most what it was handed, so (kRefitSlots - 1) entries per group can never overflow, whatever the topology
// (a treelet-restructured tree pairs far fewer siblings inside a window than a Karras tree does).
static uint32_t refit_warp_capacity(uint32_t n) { return (kRefitSlots - 1) * ((n + 31) / 32) + 256; }
// Scratch of the re-emission path: [256 B | one mask per 32 leaves | per-window lists], as in the build.
static size_t reemit_scratch_size(uint32_t n) { return 256 + align_up(sizeof(uint32_t) * (((size_t)n + 31) / 32), 256) + align_up(emit_list_bytes(n), 256); }
size_t update_scratch_size(uint32_t n)
{
    const size_t cap = (size_t)n / 4 + 256;
    size_t generic;
    if (n >= kRefitWarpPathMin)  // counters | stage-3 list | 16 words per 32-leaf group between k_refit_leaves and k_refit_window
        generic = 256 + align_up(sizeof(uint32_t) * (size_t)refit_warp_capacity(n), 256) + align_up(sizeof(uint32_t) * kRefitSlots * (((size_t)n + 31) / 32), 256);
    else
        generic = 256 + 2 * align_up(sizeof(uint32_t) * cap, 256);
    return std::max(generic, reemit_scratch_size(n));
}

// RR_BUILD_OPERATION_UPDATE.  Two paths, chosen ON THE DEVICE by the header word of the geometry buffer's tail, so that a BLAS
// stays self-describing when it is copied or broadcast:
//  * Karras-numbered tree (fast build): the closed-form emission kernels of the build run again from the deltas and sorted
//    primitive ids kept in the tail -- same topology by construction, every node rewritten whole, no atomics;
//  * restructured tree (quality build): the generic staged refit (sibling pairing by shuffles, then the parity rendezvous).
// Both sets of kernels are launched; the set that does not apply returns at once.
void update_blas(const DeviceInfo& dev, cudaStream_t s, const MeshDesc& mesh, Node* nodes, void* scratch, size_t scratch_bytes)
{
    const uint32_t n = mesh.triangle_count;
    if (n == 0) return;
    const uint32_t   kUnbounded = 0xFFFFFFFFu;
    const BlasLayout BL = blas_layout(n, false);
    char*            geom = reinterpret_cast<char*>(nodes);
    uint32_t*        karras = reinterpret_cast<uint32_t*>(geom + BL.tail_off);
    if (!scratch || scratch_bytes < update_scratch_size(n))
    {   // a client that passes no temporary buffer for updates (the Vulkan backend reports 0 bytes): one generic kernel
        k_refit<true><<<(n + 255) / 256, 256, 0, s>>>(mesh, nodes, n, nullptr, nullptr, kUnbounded, nullptr, nullptr, 0, nullptr);
        ++*dev.launches;
        RR_CUDA_CHECK(cudaGetLastError());
        return;
    }
    char* sc = (char*)scratch;
    {   // re-emission (runs only while *karras == 1)
        EmitParams p{};
        p.codes = nullptr; p.n = n; p.nodes = nodes; p.mesh = mesh;
        p.karras    = karras;
        p.deltas    = reinterpret_cast<uint8_t*>(geom + BL.tail_deltas_off);
        p.refs      = reinterpret_cast<const uint32_t*>(geom + BL.tail_refs_off);  // written there by the build's last sort pass
        p.from_tail = true;
        p.masks = reinterpret_cast<uint32_t*>(sc + 256);
        p.lists = reinterpret_cast<uint32_t*>(sc + 256 + align_up(sizeof(uint32_t) * (((size_t)n + 31) / 32), 256));
        *dev.launches += launch_emit_fit<false>(dev, s, p);
    }
    // (a debug option shrinks the lists so that tests can drive the hand-overs into their in-place fallback)
    const uint32_t cap = dev.refit_list_capacity ? std::min<uint32_t>(dev.refit_list_capacity, n / 4 + 256) : n / 4 + 256;
    RefitLists L;
    L.count_a = reinterpret_cast<uint32_t*>(sc);
    L.count_b = L.count_a + 1;
    L.items_a = reinterpret_cast<uint32_t*>(sc + 256);
    L.items_b = reinterpret_cast<uint32_t*>(sc + 256 + align_up(sizeof(uint32_t) * (size_t)cap, 256));
    L.capacity = cap;
    // (the generic path reuses the scratch of the re-emission: only one of the two is active)
    RR_CUDA_CHECK(cudaMemsetAsync(sc, 0, 8, s));
    if (n < 4096)
    {
        k_refit<true><<<(n + 255) / 256, 256, 0, s>>>(mesh, nodes, n, nullptr, nullptr, kUnbounded, nullptr, nullptr, 0, karras);
        ++*dev.launches;
    }
    else if (n >= kRefitWarpPathMin)
    {   // large meshes: warp-cooperative stages 1 (per 32 leaves) and 2 (per 512-leaf window), then the parity-rendezvous climb
        // for the ~2 % that is left.  For small ones the pipeline prologue of the persistent kernel costs more than it hides
        // (Sponza, 262 k triangles: 0.150 against 0.133 ms).
        const uint32_t wcap  = dev.refit_list_capacity ? std::min<uint32_t>(dev.refit_list_capacity, refit_warp_capacity(n)) : refit_warp_capacity(n);
        uint32_t*      slots = reinterpret_cast<uint32_t*>(sc + 256 + align_up(sizeof(uint32_t) * (size_t)refit_warp_capacity(n), 256));
        CUtensorMap tm_leaf{};
        const int   tma = node_tensor_map(&tm_leaf, nodes + (n - 1), n) ? 1 : 0;
        k_refit_leaves<<<std::min<uint32_t>((n + 255) / 256, (uint32_t)dev.sm_count * 4u), 256, 0, s>>>(mesh, nodes, n, slots, L.count_a, L.items_a,
                                                                                                   wcap, tm_leaf, tma, karras);
        const uint32_t windows = (n + kEmitWindow - 1) / kEmitWindow;
        k_refit_window<<<(windows + 7) / 8, 256, 0, s>>>(nodes, n, slots, L.count_a, L.items_a, wcap, karras);
        const uint32_t grid3 = std::min<uint32_t>((n / 4 + 511) / 256, (uint32_t)dev.sm_count * 32u);
        k_refit<false><<<grid3, 256, 0, s>>>(mesh, nodes, n, L.count_a, L.items_a, kUnbounded, nullptr, nullptr, wcap, karras);
        *dev.launches += 3;
    }
    else
    {
        k_refit<true><<<(n + 255) / 256, 256, 0, s>>>(mesh, nodes, n, nullptr, nullptr, 3u, L.count_a, L.items_a, cap, karras);
        const uint32_t grid2 = std::min<uint32_t>((cap + 255) / 256, (uint32_t)dev.sm_count * 64u);
        k_refit<false><<<grid2, 256, 0, s>>>(mesh, nodes, n, L.count_a, L.items_a, 6u, L.count_b, L.items_b, cap, karras);
        const uint32_t grid3 = std::min<uint32_t>((cap / 64 + 255) / 256 + 1, (uint32_t)dev.sm_count * 16u);
        k_refit<false><<<grid3, 256, 0, s>>>(mesh, nodes, n, L.count_b, L.items_b, kUnbounded, nullptr, nullptr, cap, karras);
        *dev.launches += 3;
    }
    RR_CUDA_CHECK(cudaGetLastError());
}

// Scene buffer: [SceneHeader | nodes (2n-1) x 64 | instance records n x 64 | forward transforms n x 48]
// (reference: [(2n-1) x 64 | 2n x 48], vlk/hlbvh_top_level_builder.cpp:378-381; the BLAS addresses that the
// reference binds as a descriptor array live in the records here, so there is no 2048-instance cap).
SceneLayout scene_layout(uint32_t n)
{
    SceneLayout L;
    L.n    = n;
    L.sort = sort_layout(n);
    size_t off = 0;
    off += 256;  // SceneHeader (64 B) in a line of its own
    L.nodes_off   = off; off += align_up(sizeof(Node) * (2 * (size_t)(n ? n : 1) - 1), 256);
    L.records_off = off; off += align_up(sizeof(InstanceRecord) * (size_t)n, 256);
    L.fwd_off     = off; off += align_up(48 * (size_t)n, 256);
    L.result_total = off;
    off = 0;
    L.aabb_off  = off; off += 32;
    off += 256 - 32;  // the scene AABB keeps a 256-byte line of its own
    L.desc_off  = off; off += align_up(sizeof(InstanceDesc) * (size_t)n, 256);
    L.boxes_off = off; off += align_up(32 * (size_t)n, 256);
    L.codes_off = off; off += align_up(sizeof(uint32_t) * (size_t)n, 256);
    L.sorted_codes_off = off; off += align_up(sizeof(uint32_t) * (size_t)n, 256);
    L.sorted_refs_off  = off; off += align_up(sizeof(uint32_t) * (size_t)n, 256);
    L.sort_off = off; off += L.sort.total;
    L.lists_off = off; off += align_up(emit_list_bytes(n), 256);
    L.scratch_total = off;
    return L;
}

void build_scene(const DeviceInfo& dev, cudaStream_t s, const InstanceDesc* host_descs, const SceneLayout& L, void* scratch, void* scene,
                 bool corner_quirk)
{
    const uint32_t n = L.n;
    if (n == 0) return;
    char* sc = (char*)scratch;
    char* out = (char*)scene;
    uint32_t*     g_aabb       = reinterpret_cast<uint32_t*>(sc + L.aabb_off);
    InstanceDesc* descs        = reinterpret_cast<InstanceDesc*>(sc + L.desc_off);
    float4*       boxes        = reinterpret_cast<float4*>(sc + L.boxes_off);
    uint32_t*     codes        = reinterpret_cast<uint32_t*>(sc + L.codes_off);
    uint32_t*     sorted_codes = reinterpret_cast<uint32_t*>(sc + L.sorted_codes_off);
    uint32_t*     sorted_refs  = reinterpret_cast<uint32_t*>(sc + L.sorted_refs_off);
    void*         sort_scratch = sc + L.sort_off;

    RR_CUDA_CHECK(cudaMemcpyAsync(descs, host_descs, sizeof(InstanceDesc) * (size_t)n, cudaMemcpyHostToDevice, s));
    reset_build_scratch(s, sc, L.aabb_off, L.sort, L.sort_off);
    SceneHeader h{};
    h.magic[0] = kSceneMagic0; h.magic[1] = kSceneMagic1; h.magic[2] = kSceneMagic2; h.magic[3] = kSceneMagic3;
    h.version = kSceneVersion; h.instance_count = n;
    h.nodes_off = L.nodes_off; h.records_off = L.records_off; h.fwd_off = L.fwd_off;
    k_instance_boxes<<<grid_for(dev, n, 256, 4), 256, 0, s>>>(descs, n, corner_quirk ? 1 : 0, boxes, g_aabb, reinterpret_cast<SceneHeader*>(out), h);
    k_morton<true><<<grid_for(dev, n, 256, 4), 256, 0, s>>>(MeshDesc{}, boxes, n, g_aabb, codes, sort_hist_ptr(L.sort, sort_scratch));
    *dev.launches += 2;
    sort_pairs(dev, s, L.sort, sort_scratch, codes, nullptr, sorted_codes, sorted_refs);
    EmitParams p{};
    p.codes = sorted_codes; p.refs = sorted_refs; p.n = n;
    p.nodes   = reinterpret_cast<Node*>(out + L.nodes_off);
    p.boxes   = boxes;
    p.descs   = descs;
    p.records = reinterpret_cast<InstanceRecord*>(out + L.records_off);
    p.fwd     = reinterpret_cast<float4*>(out + L.fwd_off);
    p.lists = reinterpret_cast<uint32_t*>(sc + L.lists_off);
    p.masks = reinterpret_cast<uint32_t*>(sc + L.sort_off + L.sort.tmp_vals_off);  // the sort's ping-pong buffer is dead now
    *dev.launches += launch_emit_fit<true>(dev, s, p);
    RR_CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(256) k_rebind_scene(void* scene, const Node* old_blas, const Node* new_blas)
{
    const SceneHeader* h = reinterpret_cast<const SceneHeader*>(scene);
    if (h->magic[0] != kSceneMagic0 || h->magic[1] != kSceneMagic1 || h->magic[2] != kSceneMagic2 || h->magic[3] != kSceneMagic3) return;
    InstanceRecord* rec = reinterpret_cast<InstanceRecord*>(static_cast<char*>(scene) + h->records_off);
    const uint32_t  n   = h->instance_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (rec[i].blas == old_blas) rec[i].blas = new_blas;
}

void rebind_scene(const DeviceInfo& dev, cudaStream_t s, void* scene, const void* old_blas, const void* new_blas)
{
    k_rebind_scene<<<dev.sm_count, 256, 0, s>>>(scene, static_cast<const Node*>(old_blas), static_cast<const Node*>(new_blas));
    *dev.launches += 1;
    RR_CUDA_CHECK(cudaGetLastError());
}

}  // namespace rr
