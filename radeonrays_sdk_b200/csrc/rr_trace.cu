// rr_trace.cu -- closest-hit / any-hit BVH2 traversal, one level (geometry) and two level (scene).
//
// Two traversal kernels share the reference's arithmetic (slab test with explicit fma, Moller-Trumbore with the shader's
// evaluation order; the library is built with --fmad=false so nothing else is contracted):
//
//  * k_trace -- restates vlk/kernels/isect.comp:88-246 and isect_2l.comp:105-323 (+ common.h:103-209): one ray per thread,
//    near child first (`c1first = hit1 && t0_0 > t0_1`), the other child deferred on a LIFO stack.  The visit order is the
//    reference's, which is what makes ANY-hit ids and first-found closest hits reproducible: it serves ANY queries, everything
//    two-level, RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, and whatever the packet kernel declines.
//  * k_trace_packet -- closest hits under the default (t, prim) rule for coherent rays: 64 rays per warp share ONE walk
//    (see the comment above it); the 64 rays are an 8 x 8 tile of the image when k_detect_grid finds the batch to be one in
//    row order.  6.05 -> 9.4 Grays/s on the C2 batch.
//
// B200 mapping of k_trace:
//  * persistent CTAs (a multiple of the SM count); each warp pulls 32-ray chunks from a ticket, so per-thread state
//    is bounded by resident threads, not by ray_count (the reference asks for 256 B of global stack per ray: 4 GiB
//    for a 16 Mi batch, vlk/geometry_trace.cpp:169);
//  * traversal stack: kSmemStack entries per thread in shared memory laid out [level][thread] (bank = lane, conflict
//    free) and nothing else in the hot loop; a ray that would need more is appended to an overflow list and traced by
//    a second, normally empty, launch (k_trace_deep) with a global-memory stack;
//  * a node is fetched with two 32-byte read-only loads (LDG.E.256) of the 64-byte aligned node;
//  * the two slab tests of a node visit use packed fma.rn.f32x2 (FFMA2) and, when the 32 rays of a chunk share a
//    direction octant, a loop specialised for that octant in which min/max(far,near) per axis is a compile-time
//    choice (bit-identical, see slab<>): 8 FMA-pipe + 8 min/max instructions instead of 12 + 20.
// Roofline: compulsory HBM traffic is 32 B/ray in + 16 B (or 4 B) out and the BVH is L2 resident, so HBM is not the
// bound.  In k_trace every lane pulls the 64 bytes of each node it visits through the SM's L1 return path (4 bytes per lane
// per clock per SM whatever the address pattern, tools/ubench/l1_broadcast.cu): 91 % busy on the C2 batch; k_trace_packet halves
// the bytes per ray by sharing the visit and is bound by issue slots / the ALU pipe (profiles/).
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "rr_internal.h"

namespace rr
{
namespace
{
constexpr int kTraceThreads = 128;
// entries per thread kept in shared memory: 24 for one-level traces (measured 16 / 20 / 24 / 28 / 32: 5187 / 5470 / 5714 /
// 5655 / 5545 Mrays/s on Sponza primary rays -- a smaller stack leaves more of the SM's 256 KB to L1, a too small one sends
// rays to the deep kernel), 32 for two-level traces (TLAS + sentinel + BLAS entries)
constexpr int kSmemStack1   = 24;
constexpr int kSmemStack2   = 32;
constexpr int kDeepStack    = 192;  // entries per thread of the deep kernel's global-memory stack

struct Vec3 { float x, y, z; };
__device__ __forceinline__ Vec3 v3(float x, float y, float z) { Vec3 r{x, y, z}; return r; }
__device__ __forceinline__ Vec3 v3(float4 q) { Vec3 r{q.x, q.y, q.z}; return r; }
__device__ __forceinline__ Vec3 sub(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
// GLSL dot()/cross() evaluation order
__device__ __forceinline__ float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b) { return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
__device__ __forceinline__ uint32_t wbits(float4 q) { return __float_as_uint(q.w); }

// safe_invdir, common.h:166-183
__device__ __forceinline__ float safe_inv(float d)
{
    const float e = 1e-5f;
    return 1.0f / (fabsf(d) > e ? d : (d < 0.0f ? -e : e));
}

// ---- packed binary32 pairs: fma.rn.f32x2 (FFMA2 on sm_100) does two IEEE fused multiply-adds in one issue slot ----
__device__ __forceinline__ uint64_t pack2(float a, float b)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// Read-only 32-byte half of a 64-byte node (LDG.E.256 on sm_100): two loads fetch a node.  volatile keeps both at the
// top of the iteration (ptxas otherwise sinks the second below the leaf test, which puts another dependent-load
// latency on every internal-node visit).
__device__ __forceinline__ void ldg_half_node(const float4* p, float4& a, float4& b)
{
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

struct RayState
{
    Vec3     o, d, inv, oxinv;
    uint64_t inv_xy, oxinv_xy;
    __device__ __forceinline__ void set(Vec3 oo, Vec3 dd)
    {
        o = oo; d = dd;
        inv      = v3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
        oxinv    = v3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
        inv_xy   = pack2(inv.x, inv.y);
        oxinv_xy = pack2(oxinv.x, oxinv.y);
    }
    // Octant of the direction (bit a set: axis a is traversed from max to min), or 8 when the specialised slab test
    // must not be used: it is bit-identical to the generic one only while every plane distance is a number.
    __device__ __forceinline__ int octant() const
    {
        const float big = 3.402823466e+38f;
        const bool  ok  = fabsf(inv.x) <= big && fabsf(inv.y) <= big && fabsf(inv.z) <= big && fabsf(oxinv.x) <= big &&
                        fabsf(oxinv.y) <= big && fabsf(oxinv.z) <= big;
        return ok ? (int)((__float_as_uint(inv.x) >> 31) | ((__float_as_uint(inv.y) >> 31) << 1) | ((__float_as_uint(inv.z) >> 31) << 2)) : 8;
    }
};

// fast_intersect_aabb, common.h:150-164.  kOct == 8: the shader's expression, min/max of both plane distances per axis.
// kOct in 0..7 (all lanes of the warp share the direction octant): fma is monotone in the plane coordinate, so for
// bmin <= bmax the larger of the two distances is known from the sign of inv alone and max(f,n) / min(f,n) -- 12 of the
// 32 arithmetic instructions of a node visit -- reduce to a compile-time choice.  Same bits in t0/t1 either way.
template <int kOct>
__device__ __forceinline__ void slab(float4 bmin, float4 bmax, const RayState& r, float t_max, float t_min, float& t0, float& t1)
{
    float lx, ly, hx, hy;
    unpack2(fma2(pack2(bmin.x, bmin.y), r.inv_xy, r.oxinv_xy), lx, ly);
    unpack2(fma2(pack2(bmax.x, bmax.y), r.inv_xy, r.oxinv_xy), hx, hy);
    const float lz = __fmaf_rn(bmin.z, r.inv.z, r.oxinv.z), hz = __fmaf_rn(bmax.z, r.inv.z, r.oxinv.z);
    float ax, ay, az, ix, iy, iz;
    if (kOct == 8)
    {
        ax = fmaxf(hx, lx); ay = fmaxf(hy, ly); az = fmaxf(hz, lz);
        ix = fminf(hx, lx); iy = fminf(hy, ly); iz = fminf(hz, lz);
    }
    else
    {
        ax = (kOct & 1) ? lx : hx; ix = (kOct & 1) ? hx : lx;
        ay = (kOct & 2) ? ly : hy; iy = (kOct & 2) ? hy : ly;
        az = (kOct & 4) ? lz : hz; iz = (kOct & 4) ? hz : lz;
    }
    t1 = fminf(fminf(az, fminf(ax, ay)), t_max);
    t0 = fmaxf(fmaxf(iz, fmaxf(ix, iy)), t_min);
}

// fast_intersect_triangle, common.h:103-137 (returns acceptance by the shader's range test)
__device__ __forceinline__ bool tri_test(const RayState& r, float min_t, float4 q0, float4 q1, float4 q2, float t_max, float& t)
{
    const Vec3  v0 = v3(q0), e1 = sub(v3(q1), v0), e2 = sub(v3(q2), v0);
    const Vec3  s1 = cross(r.d, e2);
    const float denom = dot(s1, e1);
    if (denom == 0.0f) return false;
    const float invd = 1.0f / denom;
    const Vec3  dd = sub(r.o, v0);
    const float b1 = dot(dd, s1) * invd;
    const Vec3  s2 = cross(dd, e1);
    const float b2 = dot(r.d, s2) * invd;
    const float tt = dot(e2, s2) * invd;
    if (b1 < 0.0f || b1 > 1.0f || b2 < 0.0f || b1 + b2 > 1.0f || tt < min_t || tt > t_max) return false;
    t = tt;
    return true;
}

// calculate_barycentrics, common.h:187-209, at P = o + t*d
__device__ __forceinline__ float2 barycentrics(const RayState& r, float t, float4 q0, float4 q1, float4 q2)
{
    const Vec3  p  = v3(r.o.x + t * r.d.x, r.o.y + t * r.d.y, r.o.z + t * r.d.z);
    const Vec3  v0 = v3(q0), e1 = sub(v3(q1), v0), e2 = sub(v3(q2), v0), e = sub(p, v0);
    const float d00 = dot(e1, e1), d01 = dot(e1, e2), d11 = dot(e2, e2), d20 = dot(e, e1), d21 = dot(e, e2);
    const float denom = d00 * d11 - d01 * d01;
    if (denom == 0.0f) return make_float2(0.f, 0.f);
    const float inv = 1.0f / (d00 * d11 - d01 * d01);
    return make_float2((d11 * d20 - d01 * d21) * inv, (d00 * d21 - d01 * d20) * inv);
}

// transform_ray, common.h:310-324 (direction is NOT renormalised, so t is shared between spaces)
__device__ __forceinline__ void transform_ray(const InstanceRecord* rec, Vec3 o, Vec3 d, Vec3& oo, Vec3& od)
{
    const float4 m0 = __ldg(&rec->inv0), m1 = __ldg(&rec->inv1), m2 = __ldg(&rec->inv2);
    oo = v3(dot(v3(m0), o) + m0.w, dot(v3(m1), o) + m1.w, dot(v3(m2), o) + m2.w);
    od = v3(dot(v3(m0), d), dot(v3(m1), d), dot(v3(m2), d));
}

// Traversal stack of the main kernel: kSmemStack entries per thread in shared memory, laid out [level][thread]
// (bank = lane, conflict free).  push() refuses the entry that does not fit; the ray is then handed to the deep kernel.
template <int kEntries>
struct SmemStack
{
    uint32_t* base;  // &s_stack[tid]
    int       sp;
    __device__ __forceinline__ bool push(uint32_t v)
    {
        if (sp >= kEntries) return false;
        base[sp * kTraceThreads] = v;
        ++sp;
        return true;
    }
    __device__ __forceinline__ uint32_t pop() { --sp; return base[sp * kTraceThreads]; }
};
// Stack of the deep kernel: kDeepStack entries per thread in the scratch arena, same [level][thread] layout (coalesced,
// L1/L2 resident).  A radix tree over distinct (30-bit code, 32-bit index) keys is at most 62 levels deep and treelet
// restructuring only permutes 7-leaf treelets, so kDeepStack = 192 covers TLAS + sentinel + BLAS with margin; a deeper
// entry is not stored (memory stays intact) and sets kErrorTraceStackOverflow, which the next rrWaitEvent reports.
struct DeepStack
{
    uint32_t* base;  // &arena[thread slot]
    uint32_t  stride;
    int       sp;
    uint32_t* error;  // DeviceInfo::error_word: an entry that does not fit is reported (rrWaitEvent -> RR_ERROR_INTERNAL), never silent
    __device__ __forceinline__ bool push(uint32_t v)
    {
        if (sp < kDeepStack) base[(size_t)sp * stride] = v;
        else atomicOr(error, kErrorTraceStackOverflow);
        ++sp;
        return true;
    }
    __device__ __forceinline__ uint32_t pop() { --sp; return sp < kDeepStack ? base[(size_t)sp * stride] : kInvalid; }
};

struct TraceParams
{
    const Node*           bvh;
    const InstanceRecord* instances;
    const float4*         rays;
    uint32_t              ray_count;
    const uint32_t*       indirect;
    void*                 hits;
    uint32_t*             arena;          // deep-kernel stacks
    uint32_t*             ticket;         // scratch header word 0: chunk counter of the main kernel
    uint32_t*             overflow_count; // word 1: rays whose stack did not fit in shared memory
    uint32_t*             deep_ticket;    // word 2: chunk counter of the deep kernel
    uint32_t*             overflow_list;  // [ray_count] indices of those rays
    uint32_t*             packet_ticket;  // word 3: packet counter of k_trace_packet
    uint32_t*             chunk_count;    // word 4: 32-ray chunks k_trace_packet handed to the per-ray kernel
    uint32_t*             chunk_list;     // [ceil(ray_count / 32)] their chunk indices; nullptr: k_trace walks all chunks itself
    const uint32_t*       grid;           // words 5, 6, 7: row length of the ray grid k_detect_grid found (0: none) and the (signed) ray index
                                          // of the first ray of its row 0; "the batch is incoherent"; nullptr: packets / chunks of consecutive rays
    const uint32_t*       perm;           // RR_CUDA_OPTION_SORT_RAYS: position -> ray index in binned order; nullptr: client order
    int                   first_found;
    int                   force_generic;
    uint32_t*             error;          // DeviceInfo::error_word
};

// A geometry buffer starts with node 0, a scene buffer with a SceneHeader (rr_internal.h): every kernel checks on the device that
// the buffer is of the kind it traverses and returns at once otherwise (rrCmdIntersect launches both kinds; the reference keeps a
// host-side map keyed by the buffer instead, vlk/intersector.cpp:289-290).  For a scene, bvh / instances are taken from the header.
template <bool kTwoLevel>
__device__ __forceinline__ bool resolve_scene(TraceParams& P)
{
    const uint4 m     = __ldg(reinterpret_cast<const uint4*>(P.bvh));
    const bool  scene = m.x == kSceneMagic0 && m.y == kSceneMagic1 && m.z == kSceneMagic2 && m.w == kSceneMagic3;
    if (scene != kTwoLevel) return false;
    if (kTwoLevel)
    {
        const char*        base = reinterpret_cast<const char*>(P.bvh);
        const SceneHeader* h    = reinterpret_cast<const SceneHeader*>(base);
        P.instances = reinterpret_cast<const InstanceRecord*>(base + h->records_off);
        P.bvh       = reinterpret_cast<const Node*>(base + h->nodes_off);
    }
    return true;
}

// Ray grids (k_detect_grid below): the mapping from (tile, half, lane) to a ray index, shared by k_trace_packet and k_trace.
constexpr uint32_t kGridMinWidth = 64, kGridMinRows = 8, kGridSearch = 32768;
constexpr uint32_t kAllChunks = 0xFFFFFFFFu;   // chunk count that means "no list: every 32-ray chunk of the batch"
struct RayGrid
{
    uint32_t w;         // 0: no grid
    int32_t  base;      // ray index of the first ray of row 0 (<= 0: the batch may start inside row 0)
    uint32_t tiles_x, tiles_y;
};
__device__ __forceinline__ RayGrid load_grid(const uint32_t* words, uint32_t count)
{
    RayGrid g;
    g.w = words ? __ldg(words) : 0u;
    g.base = 0; g.tiles_x = g.tiles_y = 0;
    if (g.w)
    {
        g.base = (int32_t)__ldg(words + 1);
        const uint64_t span = (uint64_t)((int64_t)count - g.base);   // rays of the virtual grid, from (0, 0) to the last ray
        g.tiles_x = (g.w + 7) / 8;
        g.tiles_y = (uint32_t)(((span + g.w - 1) / g.w + 7) / 8);
    }
    return g;
}
// The ray of lane `lane` in half `slot` (0: rows 0-3 of the tile, 1: rows 4-7) of packet `packet`; false when the tile sticks out.
__device__ __forceinline__ bool grid_ray(const RayGrid& g, uint32_t packet, uint32_t slot, uint32_t lane, uint32_t count, uint32_t& index)
{
    const uint32_t ty = packet / g.tiles_x, tx = packet - ty * g.tiles_x;
    const uint32_t x = tx * 8 + (lane & 7), y = ty * 8 + slot * 4 + (lane >> 3);
    const int64_t  i = (int64_t)y * g.w + x + g.base;
    index = (uint32_t)i;
    return x < g.w && i >= 0 && i < (int64_t)count;
}

// Leaving an instance: restore the world-space ray (isect_2l.comp:279-287) and pop again.
#define RR_POP_NEXT()                                                         \
    do                                                                        \
    {                                                                         \
        addr = st.pop();                                                      \
        if (kTwoLevel && addr == kSentinel)                                   \
        {                                                                     \
            cur_inst = kInvalid;                                              \
            cur_bvh  = P.bvh;                                                 \
            ray.set(v3(r0), v3(r1));                                          \
            addr = st.pop();                                                  \
        }                                                                     \
    } while (0)

// One ray, start to finish: the reference's loop (isect.comp:121-216 / isect_2l.comp:173-287) in its "if-if" shape --
// every iteration fetches the current node and then either tests its two child boxes and descends / defers, enters
// an instance, or tests the triangle.  Measured alternatives that were NOT kept (profiles/round1_trace_modes.md):
// a while-while restructuring (descend until a leaf, then test) was 25-40 % slower on B200, and per-lane ray refill
// (persistent threads with batched replacement) 20-60 % slower, on coherent and incoherent rays alike.
template <bool kAny, bool kFullHit, bool kTwoLevel, int kOct, class StackT>
__device__ __forceinline__ void trace_ray(const TraceParams& P, StackT& st, uint32_t gidx, bool valid, float4 r0, float4 r1, RayState& ray)
{
    const float min_t        = r0.w;
    float       closest      = r1.w;
    uint32_t    closest_addr = kInvalid, closest_prim = kInvalid, closest_inst = kInvalid;
    uint32_t    cur_inst     = kInvalid;
    const Node* cur_bvh      = P.bvh;
    st.sp = 0;
    st.push(kInvalid);
    uint32_t addr = valid ? 0u : kInvalid;
    while (addr != kInvalid)
    {
        const float4* np = reinterpret_cast<const float4*>(cur_bvh + addr);
        float4 q0, q1, q2, q3;
        ldg_half_node(np, q0, q1);
        ldg_half_node(np + 2, q2, q3);
        if (wbits(q0) != kInvalid)
        {
            float a0, a1, b0, b1;
            slab<kOct>(q0, q1, ray, closest, min_t, a0, a1);
            slab<kOct>(q2, q3, ray, closest, min_t, b0, b1);
            const bool t0 = a0 <= a1, t1 = b0 <= b1;
            const bool c1first = t1 && (a0 > b0);
            const bool take1   = c1first || !t0;
            const uint32_t c0 = wbits(q0), c1 = wbits(q1);
            const uint32_t near_child = take1 ? c1 : c0, far_child = take1 ? c0 : c1;
            if (t0 && t1 && !st.push(far_child)) goto overflow;
            if (t0 || t1) { addr = near_child; continue; }
        }
        else if (kTwoLevel && cur_inst == kInvalid)
        {   // top-level leaf: enter the instance (isect_2l.comp:231-245)
            cur_inst = wbits(q1);
            const InstanceRecord* rec = P.instances + cur_inst;
            Vec3 oo, od;
            transform_ray(rec, ray.o, ray.d, oo, od);
            ray.set(oo, od);
            cur_bvh = rec->blas;
            if (!st.push(kSentinel)) goto overflow;
            addr = 0;
            continue;
        }
        else
        {
            float t;
            if (tri_test(ray, min_t, q0, q1, q2, closest, t))
            {
                const uint32_t prim = wbits(q1);
                bool accept = t < closest;
                if (!kAny && !accept && !P.first_found && t == closest && closest_addr != kInvalid)
                    accept = kTwoLevel ? (cur_inst < closest_inst || (cur_inst == closest_inst && prim < closest_prim))
                                       : (prim < closest_prim);
                if (accept)
                {
                    if (kAny)
                    {   // first accepted triangle in traversal order (isect.comp:181-200)
                        if (kFullHit)
                        {
                            const float2 uv = barycentrics(ray, t, q0, q1, q2);
                            __stcs(reinterpret_cast<float4*>(P.hits) + gidx,
                                   make_float4(uv.x, uv.y, __uint_as_float(kTwoLevel ? cur_inst : 0u), __uint_as_float(prim)));
                        }
                        else
                            __stcs(reinterpret_cast<uint32_t*>(P.hits) + gidx, kTwoLevel ? cur_inst : prim);
                        return;
                    }
                    closest      = t;
                    closest_addr = addr;
                    closest_prim = prim;
                    closest_inst = cur_inst;
                }
            }
        }
        RR_POP_NEXT();
    }
    if (!valid) return;
    if (closest_addr != kInvalid)
    {
        if (kFullHit)
        {
            const Node* hb = P.bvh;
            if (kTwoLevel)
            {
                const InstanceRecord* rec = P.instances + closest_inst;
                Vec3 oo, od;
                transform_ray(rec, v3(r0), v3(r1), oo, od);
                ray.o = oo; ray.d = od;
                hb = rec->blas;
            }
            const float4* np = reinterpret_cast<const float4*>(hb + closest_addr);
            const float2  uv = barycentrics(ray, closest, __ldg(np), __ldg(np + 1), __ldg(np + 2));
            __stcs(reinterpret_cast<float4*>(P.hits) + gidx,
                   make_float4(uv.x, uv.y, __uint_as_float(kTwoLevel ? closest_inst : 0u), __uint_as_float(closest_prim)));
        }
        else
            __stcs(reinterpret_cast<uint32_t*>(P.hits) + gidx, kTwoLevel ? closest_inst : closest_prim);
    }
    else
    {   // miss: only the id word is written (isect.comp:238-245)
        if (kFullHit) reinterpret_cast<uint32_t*>(P.hits)[4 * (size_t)gidx + 2] = kInvalid;
        else reinterpret_cast<uint32_t*>(P.hits)[gidx] = kInvalid;
    }
    return;
overflow:
    // nothing has been written for this ray yet: the deep kernel traces it again from the start
    P.overflow_list[atomicAdd(P.overflow_count, 1u)] = gidx;
}
#undef RR_POP_NEXT

// Persistent warps pull one 32-ray chunk at a time from a global ticket (P.ticket, zeroed by a memset node before
// the launch) and trace it to completion.  Against a static warp-stride assignment this removes the end-of-kernel
// imbalance: +29 % on coherent primary rays, +39 % on shadow rays (profiles/round1_trace_modes.md).
// One-level kernels dispatch each chunk to the loop specialised for its direction octant when all 32 rays agree
// (coherent batches almost always do); mixed chunks, and every two-level trace (the ray changes octant per instance),
// run the generic loop.
// kList: the chunks come from the list the packet kernel left behind (its own instantiation, so that the loops of the other
// queries do not carry the list / ray-grid code: they sit at their register limit).
template <bool kAny, bool kFullHit, bool kTwoLevel, bool kList>
__global__ void __launch_bounds__(kTraceThreads, kTwoLevel ? 8 : 10) k_trace(TraceParams P)
{
    constexpr int kEntries = kTwoLevel ? kSmemStack2 : kSmemStack1;
    __shared__ uint32_t s_stack[kEntries * kTraceThreads];
    if (!resolve_scene<kTwoLevel>(P)) return;
    uint32_t count = P.ray_count;
    if (P.indirect) count = min(count, __ldg(P.indirect));  // isect.comp:98-103
    const uint32_t lane = threadIdx.x & 31;
    SmemStack<kEntries> st;
    st.base = s_stack + threadIdx.x;
    const uint32_t listed = kList ? *P.chunk_count : 0u;
    if (kList && listed == 0) return;   // (coherent batches: nothing was declined; no ticket traffic for nothing)
    const bool    all = kList && listed == kAllChunks;   // k_trace_packet stood aside: every chunk, in order
    const RayGrid G = load_grid(kList && !all ? P.grid : nullptr, count);
    while (true)
    {
        uint32_t chunk = 0;
        if (lane == 0)
        {
            chunk = atomicAdd(P.ticket, 1u);
            // list mode: only the chunks the packet kernel declined (incoherent rays, deep trees)
            if (kList && !all) chunk = chunk < listed ? P.chunk_list[chunk] : kInvalid;
        }
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        uint32_t gidx;
        bool     valid;
        if (kList && G.w)
        {   // a declined half of an 8 x 8 tile of the ray grid (k_trace_packet): chunk = 2 * tile + half
            if (chunk == kInvalid) break;
            valid = grid_ray(G, chunk >> 1, chunk & 1u, lane, count, gidx);
        }
        else if (!kList && P.grid && __ldg(P.grid))
        {   // image-ordered rays: a warp takes an 8 x 4 half tile instead of 32 consecutive rays (its slowest lane is 5 % closer to
            // the average).  The grid is re-read per chunk (L1) rather than kept in registers across the traversal.
            const RayGrid g = load_grid(P.grid, count);
            if (chunk >= 2 * (uint64_t)g.tiles_x * g.tiles_y) break;
            valid = grid_ray(g, chunk >> 1, chunk & 1u, lane, count, gidx);
        }
        else
        {
            if ((uint64_t)chunk * 32 >= count) break;
            gidx  = chunk * 32 + lane;
            valid = gidx < count;
        }
        // binned order: the warp takes 32 neighbours of the sorted sequence; everything below (ray load, hit store, overflow
        // list) uses the ray's own index, so the client sees its own order (rays past a device-side count sort last)
        if (!kList && P.perm) gidx = __ldg(P.perm + (valid ? gidx : count - 1));
        const uint32_t ridx  = valid ? gidx : (P.perm ? gidx : count - 1);  // tail lanes shadow a valid ray and write nothing
        // rays are read once: streaming loads (evict-first) keep them from displacing BVH nodes in L1 / L2
        const float4 r0 = __ldcs(P.rays + 2 * (size_t)ridx), r1 = __ldcs(P.rays + 2 * (size_t)ridx + 1);
        RayState ray;
        ray.set(v3(r0), v3(r1));
        int oct = 8;
        if (!kTwoLevel && !P.force_generic)
        {
            oct = ray.octant();
            if (!__all_sync(0xffffffffu, oct == __shfl_sync(0xffffffffu, oct, 0))) oct = 8;
        }
        switch (oct)
        {
        case 0: trace_ray<kAny, kFullHit, kTwoLevel, 0>(P, st, gidx, valid, r0, r1, ray); break;
        case 1: trace_ray<kAny, kFullHit, kTwoLevel, 1>(P, st, gidx, valid, r0, r1, ray); break;
        case 2: trace_ray<kAny, kFullHit, kTwoLevel, 2>(P, st, gidx, valid, r0, r1, ray); break;
        case 3: trace_ray<kAny, kFullHit, kTwoLevel, 3>(P, st, gidx, valid, r0, r1, ray); break;
        case 4: trace_ray<kAny, kFullHit, kTwoLevel, 4>(P, st, gidx, valid, r0, r1, ray); break;
        case 5: trace_ray<kAny, kFullHit, kTwoLevel, 5>(P, st, gidx, valid, r0, r1, ray); break;
        case 6: trace_ray<kAny, kFullHit, kTwoLevel, 6>(P, st, gidx, valid, r0, r1, ray); break;
        case 7: trace_ray<kAny, kFullHit, kTwoLevel, 7>(P, st, gidx, valid, r0, r1, ray); break;
        default: trace_ray<kAny, kFullHit, kTwoLevel, 8>(P, st, gidx, valid, r0, r1, ray); break;
        }
    }
}


// ---- packet traversal of coherent closest-hit rays ------------------------------------------------------------------------------
// k_trace above is bound by the SM's L1 return path: every lane receives the 64 bytes of each node it visits, 4 bytes per lane per
// clock per SM whatever the address pattern (tools/ubench/l1_broadcast.cu: a warp-uniform LDG.128 costs the same 4 clk as a
// divergent one that hits one line; sm_100a has no global load into uniform registers).  What can be shared is the VISIT: when the
// rays of a warp are coherent they walk almost the same nodes (C2: 56 internal nodes per ray, 73 in the union of 64 neighbouring
// rays), so here a warp owns 64 consecutive rays -- two per lane, "slots" of 32 -- and ONE traversal: one node fetch, one stack
// (in shared memory, warp uniform) and one descend / defer decision per node for all 64 rays; a lane still tests its own two rays
// against both child boxes with the reference's arithmetic (slab<kOct>) and against each triangle (tri_test), and a child is
// entered when any ray wants it.  Per 64 rays that is half the node bytes of two 32-ray warps, no divergence in the triangle test, and the
// control flow amortised over two rays per lane.
// Closest-hit only, and only under the default tie rule: a ray's result is the (t, prim) minimum over the triangles it accepts,
// which does not depend on the visit order; every triangle a ray would reach on its own walk is reached here too because the
// lane votes with the same box test against its own running closest t.  (Exact-t ties at the ulp level of the slab test can
// resolve differently from k_trace; they equal the brute-force (t, prim) minimum more often, tests/test_gpu_trace.py counts them.)
// ANY queries and RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND depend on the reference's visit order and stay on k_trace.
// A packet whose 64 rays do not share a direction octant, or whose shared stack would overflow, is handed to k_trace through
// the chunk list (two 32-ray chunks), which then runs in list mode right after this kernel.
constexpr int      kPacketThreads = 128;
constexpr int      kPacketStack   = 64;            // entries per warp; LBVH depth beyond that goes to k_trace / k_trace_deep

template <bool kFullHit>
__device__ __forceinline__ void write_closest(const TraceParams& P, uint32_t gidx, const float4& r0, const float4& r1, float closest,
                                              uint32_t closest_addr)
{   // the primitive id is word 7 of the leaf (the traversal keeps only the leaf's address: registers)
    if (closest_addr != kInvalid)
    {
        const float4* np = reinterpret_cast<const float4*>(P.bvh + closest_addr);
        if (kFullHit)
        {
            RayState ray;
            ray.o = v3(r0); ray.d = v3(r1);
            const float4 l1 = __ldg(np + 1);
            const float2 uv = barycentrics(ray, closest, __ldg(np), l1, __ldg(np + 2));
            __stcs(reinterpret_cast<float4*>(P.hits) + gidx, make_float4(uv.x, uv.y, __uint_as_float(0u), l1.w));
        }
        else
            __stcs(reinterpret_cast<uint32_t*>(P.hits) + gidx, __ldg(reinterpret_cast<const uint32_t*>(np) + 7));
    }
    else
    {   // miss: only the id word is written (isect.comp:238-245)
        if (kFullHit) reinterpret_cast<uint32_t*>(P.hits)[4 * (size_t)gidx + 2] = kInvalid;
        else reinterpret_cast<uint32_t*>(P.hits)[gidx] = kInvalid;
    }
}

// Both slots of a lane against one box: the six plane distances of the two rays come out of six FFMA2, each with the box coordinate
// as a broadcast scalar operand and the two rays' (inv, -o * inv) packed per axis -- the same fma per ray as slab<kOct>, so the same
// bits -- instead of 2 FFMA2 + 2 FFMA per ray.
// One lane of the (converged) warp: ELECT instead of reading the lane id and comparing.
__device__ __forceinline__ bool elect_one()
{
    uint32_t p;
    asm volatile("{ .reg .pred e; elect.sync _|e, 0xffffffff; selp.u32 %0, 1, 0, e; }" : "=r"(p));
    return p != 0;
}

struct PairConsts
{
    uint64_t inv[3], ox[3];   // per axis: (slot 0, slot 1)
};
template <int kOct>
__device__ __forceinline__ void slab_pair(float4 bmin, float4 bmax, const PairConsts& c, const float (&t_max)[2], const float (&t_min)[2],
                                          float (&t0)[2], float (&t1)[2])
{
    float lx[2], ly[2], lz[2], hx[2], hy[2], hz[2];
    unpack2(fma2(pack2(bmin.x, bmin.x), c.inv[0], c.ox[0]), lx[0], lx[1]);
    unpack2(fma2(pack2(bmax.x, bmax.x), c.inv[0], c.ox[0]), hx[0], hx[1]);
    unpack2(fma2(pack2(bmin.y, bmin.y), c.inv[1], c.ox[1]), ly[0], ly[1]);
    unpack2(fma2(pack2(bmax.y, bmax.y), c.inv[1], c.ox[1]), hy[0], hy[1]);
    unpack2(fma2(pack2(bmin.z, bmin.z), c.inv[2], c.ox[2]), lz[0], lz[1]);
    unpack2(fma2(pack2(bmax.z, bmax.z), c.inv[2], c.ox[2]), hz[0], hz[1]);
#pragma unroll
    for (int k = 0; k < 2; ++k)
    {
        const float ax = (kOct & 1) ? lx[k] : hx[k], ix = (kOct & 1) ? hx[k] : lx[k];
        const float ay = (kOct & 2) ? ly[k] : hy[k], iy = (kOct & 2) ? hy[k] : ly[k];
        const float az = (kOct & 4) ? lz[k] : hz[k], iz = (kOct & 4) ? hz[k] : lz[k];
        t1[k] = fminf(fminf(az, fminf(ax, ay)), t_max[k]);
        t0[k] = fmaxf(fmaxf(iz, fmaxf(ix, iy)), t_min[k]);
    }
}

// ---- ray grids: 8 x 8 tiles instead of 64 x 1 strips -------------------------------------------------------------------------------
// The coherent batches this kernel is for are camera rays in image order (the reference's tests generate them so: for y, for x).
// 64 consecutive rays are then a 64 x 1 strip of pixels; an 8 x 8 tile of the same image walks 16 % fewer internal nodes and
// 43 % fewer leaves as a packet (CPU model of the packet walk on the C2 batch: 72.6 + 12.1 -> 61.3 + 6.9 visits per packet).
// Which rays share a packet never changes a result -- every ray keeps its own (t, prim) minimum and writes its hit at its own
// index -- so the row length may be found by looking at the rays: k_detect_grid takes d(1) - d(0) as the step along a row (the
// origins' step when all directions are equal) and calls ray i a row end when the step from i to i + 1 points backwards; the
// first two row ends give the row length W and the phase of row 0 (a batch may start in mid-row: a shard of a frame), and a
// few more rows are checked.  No grid found (W < 64, fewer than 8 rows, jittered or unordered rays): strips, as before.
// RR_CUDA_OPTION_RAY_GRID_WIDTH: 0 = detect (default), 1 = strips only, W >= 64 = the client says the rows are W rays long.
__global__ void __launch_bounds__(1024) k_detect_grid(const float4* __restrict__ rays, uint32_t ray_count, const uint32_t* __restrict__ indirect,
                                                       uint32_t* __restrict__ words, uint32_t explicit_w, uint32_t chunk_capacity)
{
    __shared__ uint32_t s_first, s_second, s_bad;
    uint32_t count = ray_count;
    if (indirect) count = min(count, __ldg(indirect));
    uint32_t w = 0;
    int32_t  base = 0;
    // An incoherent batch (diffuse bounces) would only be read once more by the packet kernel to be declined 64 rays at a time:
    // 4 096 pairs of consecutive rays spread over the batch say so beforehand (word 2: the packet kernel hands ALL chunks to the
    // per-ray kernel at once).  Camera rays change octant on a few lines of the image; a quarter of the pairs is far from that.
    uint32_t mixed = 0;
    if (explicit_w == 0 && count >= 8192)
    {
        float4 a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const uint32_t i = (uint32_t)(((uint64_t)(threadIdx.x * 4 + k) * (count - 1)) >> 12);
            a[k] = __ldg(rays + 2 * (size_t)i + 1); b[k] = __ldg(rays + 2 * (size_t)i + 3);
        }
        int differ = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            differ += ((__float_as_uint(a[k].x) ^ __float_as_uint(b[k].x)) | (__float_as_uint(a[k].y) ^ __float_as_uint(b[k].y)) |
                       (__float_as_uint(a[k].z) ^ __float_as_uint(b[k].z))) >> 31;
        int total = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) total += __syncthreads_count(differ > k);   // sum over the block of `differ` (0..4)
        mixed = total > 1024 ? 1u : 0u;
    }
    if (mixed) { /* nothing to tile */ }
    else if (explicit_w >= kGridMinWidth) w = explicit_w;
    else if (count >= kGridMinWidth * kGridMinRows)
    {
        // the step along a row: directions, or origins when the first two directions are equal (parallel rays)
        if (threadIdx.x == 0) { s_first = s_second = 0xFFFFFFFFu; s_bad = 0; }
        __syncthreads();
        const uint32_t limit = min(count - 1, kGridSearch);
        const float4 o0 = __ldg(rays), d0 = __ldg(rays + 1), o1 = __ldg(rays + 2), d1 = __ldg(rays + 3);
        float4 da[8], db[8];   // round 0's directions are requested together with the first two rays: one round trip
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            const uint32_t i = min(k * 1024 + threadIdx.x, limit - 1);
            da[k] = __ldg(rays + 2 * (size_t)i + 1); db[k] = __ldg(rays + 2 * (size_t)i + 3);
        }
        float sx = d1.x - d0.x, sy = d1.y - d0.y, sz = d1.z - d0.z;
        const bool by_origin = sx == 0.f && sy == 0.f && sz == 0.f;
        if (by_origin) { sx = o1.x - o0.x; sy = o1.y - o0.y; sz = o1.z - o0.z; }
        const bool usable = sx * sx + sy * sy + sz * sz > 0.f;
        auto row_end = [&](uint32_t i) -> bool {   // i + 1 < count
            const float4 a = __ldg(rays + 2 * (size_t)i + (by_origin ? 0 : 1)), b = __ldg(rays + 2 * (size_t)i + (by_origin ? 2 : 3));
            return (b.x - a.x) * sx + (b.y - a.y) * sy + (b.z - a.z) * sz < 0.f;
        };
        bool     e[8];
        uint32_t w0 = 0;
        for (; usable && w0 < limit; w0 += 8 * 1024)
        {   // 8 192 indices per round, eight loads in flight per thread; a round that holds the second row end is the last
#pragma unroll
            for (int k = 0; k < 8; ++k)
            {
                const uint32_t i = w0 + k * 1024 + threadIdx.x;
                if (w0 == 0 && !by_origin) e[k] = i < limit && (db[k].x - da[k].x) * sx + (db[k].y - da[k].y) * sy + (db[k].z - da[k].z) * sz < 0.f;
                else e[k] = i < limit && row_end(i);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (e[k]) atomicMin(&s_first, w0 + k * 1024 + threadIdx.x);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (e[k] && w0 + k * 1024 + threadIdx.x > s_first) atomicMin(&s_second, w0 + k * 1024 + threadIdx.x);
            __syncthreads();
            if (s_second != 0xFFFFFFFFu) break;
        }
        const uint32_t j1 = s_first, j2 = s_second;
        if (j2 != 0xFFFFFFFFu && j2 - j1 >= kGridMinWidth && j1 + 1 <= j2 - j1)
        {
            w    = j2 - j1;
            base = (int32_t)(j1 + 1) - (int32_t)w;
            // the last round (earlier ones held at most the first row end): a row end exactly at j1 + m W, nowhere else
#pragma unroll
            for (int k = 0; k < 8; ++k)
            {
                const uint32_t i = w0 + k * 1024 + threadIdx.x;
                if (e[k] != (i < limit && i >= j1 && (i - j1) % w == 0)) s_bad = 1;
            }
            // ... and a few rows spread over the rest of the batch end where they should, and not half a row earlier
            if (threadIdx.x < 64)
            {
                const uint64_t rows = ((uint64_t)count - (j1 + 1)) / w;
                const uint64_t r    = rows * (threadIdx.x + 1) / 65;
                const uint64_t i    = (uint64_t)j1 + r * w;
                if (r > 0 && i + 1 < count && (!row_end((uint32_t)i) || row_end((uint32_t)i - w / 2))) s_bad = 1;
            }
            __syncthreads();
            if (s_bad) w = 0;
        }
    }
    if (threadIdx.x == 0)
    {
        if (w)
        {   // at least eight rows, and the declined-chunk list must hold two chunks for every tile
            const uint64_t rows  = ((uint64_t)((int64_t)count - base) + w - 1) / w;
            const uint64_t tiles = (uint64_t)((w + 7) / 8) * ((rows + 7) / 8);
            if (rows < kGridMinRows || 2 * tiles > chunk_capacity || base > 0) w = 0;
        }
        words[0] = w;
        words[1] = (uint32_t)base;
        words[2] = mixed;
    }
}

// One packet, start to finish.  Returns false when the shared stack overflowed (nothing has been written then).
// The loop is written for issue slots (the kernel is issue bound, profiles/round2_summary.md): both slots are tested at every
// internal node (their arithmetic interleaves), a lane without a ray carries closest = -FLT_MAX so that no test of it can pass,
// one predicate per child ("any of my two rays wants it") feeds one VOTE.ANY, the near child is the one some slot-0 ray enters
// first, and the stack is addressed through a 32-bit shared-memory address.
// Leaves are never visited as nodes of their own: the builder marks in an internal node's `update` word which of its children are
// leaves (rr_internal.h node_update_word), and a leaf child is tested right where its box was -- its triangle is fetched only when
// some ray's box test passed, and only the slots that have such a ray run the Moeller-Trumbore test (a leaf's box is small: 60 %
// of the leaf visits of the C2 batch concern one of the two 32-ray slots only).  `cur` and the stack hold internal nodes only.
template <bool kFullHit, int kOct, bool kStatic>
__device__ __forceinline__ bool trace_packet(const TraceParams& P, uint32_t stack_lo, uint32_t gidx0, uint32_t gstep, const bool (&valid)[2],
                                             const float4 (&r0)[2], const float4 (&r1)[2], const RayState (&ray)[2])
{
    const float kNever = -3.402823466e+38f;
    float    closest[2]      = {valid[0] ? r1[0].w : kNever, valid[1] ? r1[1].w : kNever};
    uint32_t closest_addr[2] = {kInvalid, kInvalid};             // the leaf of the closest hit so far
    // shared-memory byte address of the next free entry; a warp's stack is the first 256 bytes of a 512-byte aligned block, so
    // "empty" and "full" are bit tests on the pointer (no bounds held in registers)
    uint32_t sp  = stack_lo;
    uint32_t cur = 0u;                                         // the root: an internal node (k_trace_packet checked)
    PairConsts pc;
    pc.inv[0] = pack2(ray[0].inv.x, ray[1].inv.x); pc.inv[1] = pack2(ray[0].inv.y, ray[1].inv.y); pc.inv[2] = pack2(ray[0].inv.z, ray[1].inv.z);
    pc.ox[0] = pack2(ray[0].oxinv.x, ray[1].oxinv.x); pc.ox[1] = pack2(ray[0].oxinv.y, ray[1].oxinv.y); pc.ox[2] = pack2(ray[0].oxinv.z, ray[1].oxinv.z);
    const float min_t[2] = {r0[0].w, r0[1].w};

    while (true)
    {
        float4 q0, q1, q2, q3;
        const float4* np = reinterpret_cast<const float4*>(P.bvh + cur);
        ldg_half_node(np, q0, q1);
        ldg_half_node(np + 2, q2, q3);
        float e0[2], e1[2], f0[2], f1[2];   // child 0: [e0, e1], child 1: [f0, f1], per slot
        slab_pair<kOct>(q0, q1, pc, closest, min_t, e0, e1);
        slab_pair<kOct>(q2, q3, pc, closest, min_t, f0, f1);
        const uint32_t tag = wbits(q3);
        if ((tag & (kNodeLeaf0 | kNodeLeaf1)) == 0)
        {   // two internal children
            const bool     any0  = __any_sync(0xffffffffu, e0[0] <= e1[0] || e0[1] <= e1[1]);
            const bool     any1  = __any_sync(0xffffffffu, f0[0] <= f1[0] || f0[1] <= f1[1]);
            // child 1 first: the builder's static order for this octant on treelet-optimised trees (rr_internal.h node_order_bits:
            // no FSETP, no VOTE), else when it is the nearer one for any slot-0 ray
            // (which of the two: decided once per launch from the root's word, k_trace_packet -- a test per node costs more than either)
            const bool first1 = kStatic ? (tag & (1u << (kNodeOrderShift + kOct))) != 0 : __any_sync(0xffffffffu, f0[0] < e0[0]);
            const bool     take1 = any1 && (!any0 || first1);
            if (any0 && any1)
            {   // defer the other child: one predicated store each instead of a select (the kernel is bound by the ALU pipe)
                if (sp & (kPacketStack * 4)) return false;
                __syncwarp();  // every lane's earlier pop of this slot has completed
                if (elect_one())
                {   // one lane writes the warp's stack; the barrier below orders the store before any lane's later pop
                    if (take1) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sp), "r"(wbits(q0)) : "memory");
                    else       asm volatile("st.shared.u32 [%0], %1;" ::"r"(sp), "r"(wbits(q1)) : "memory");
                }
                __syncwarp();
                sp += 4;
            }
            if (any0 || any1)
            {
                cur = take1 ? wbits(q1) : wbits(q0);
                continue;
            }
        }
        else
        {   // at least one leaf child: test it here; at most one internal child is left to descend into, nothing to defer.
            // Written so that one register (`other`) and two predicates survive a triangle test: the second child, leaf or not.
            const bool leaf0 = (tag & kNodeLeaf0) != 0, both = (tag & (kNodeLeaf0 | kNodeLeaf1)) == (kNodeLeaf0 | kNodeLeaf1);
            uint32_t leaf, other;
            bool     want0, want1, other0, other1;
            if (leaf0)
            {   // (a uniform branch instead of selects between predicates: those cost ~20 instructions)
                leaf = wbits(q0); other = wbits(q1);
                want0  = __any_sync(0xffffffffu, e0[0] <= e1[0]); want1  = __any_sync(0xffffffffu, e0[1] <= e1[1]);
                other0 = __any_sync(0xffffffffu, f0[0] <= f1[0]); other1 = __any_sync(0xffffffffu, f0[1] <= f1[1]);
            }
            else
            {
                leaf = wbits(q1); other = wbits(q0);
                want0  = __any_sync(0xffffffffu, f0[0] <= f1[0]); want1  = __any_sync(0xffffffffu, f0[1] <= f1[1]);
                other0 = __any_sync(0xffffffffu, e0[0] <= e1[0]); other1 = __any_sync(0xffffffffu, e0[1] <= e1[1]);
            }
            bool again = both;
            while (true)
            {
                if (want0 || want1)
                {
                    const float4*  lp = reinterpret_cast<const float4*>(P.bvh + leaf);
                    float4 l0, l1;
                    ldg_half_node(lp, l0, l1);
                    const float4   l2 = __ldg(lp + 2);
                    const uint32_t prim = wbits(l1);
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                    {
                        if (k == 0 ? want0 : want1)
                        {
                            float t;
                            if (tri_test(ray[k], min_t[k], l0, l1, l2, closest[k], t) &&
                                (t < closest[k] || (closest_addr[k] != kInvalid && prim < __ldg(reinterpret_cast<const uint32_t*>(P.bvh + closest_addr[k]) + 7))))
                            {   // tri_test accepted t <= closest: smaller t, or the same t and a lower primitive id (fetched from the
                                // leaf of the hit so far in that rare case)
                                closest[k]      = t;
                                closest_addr[k] = leaf;
                            }
                        }
                    }
                }
                if (!again) break;
                again = false;
                leaf = other; want0 = other0; want1 = other1;
            }
            if (!both && (other0 || other1))
            {   // the internal sibling (decided with the closest hits as they were when its box was tested)
                cur = other;
                continue;
            }
        }
        if ((sp & (2 * kPacketStack * 4 - 1)) == 0) break;
        sp -= 4;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(sp) : "memory");
    }
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (valid[k]) write_closest<kFullHit>(P, gidx0 + gstep * k, r0[k], r1[k], closest[k], closest_addr[k]);
    return true;
}

// (Round 2 also measured this kernel with the top 8 levels of the tree staged in shared memory per CTA -- BASELINE's "staging of top
// BVH levels": 6 492 against 6 970 Mrays/s on C2, profiles/round2_summary.md; the top of the tree is the part that always hits in L1
// anyway, and the tag test adds instructions to a loop that is bound by issue slots.  The variant is in the history, not here.)
template <bool kFullHit>
__global__ void __launch_bounds__(kPacketThreads, 8) k_trace_packet(TraceParams P)
{
    __shared__ __align__(2 * kPacketStack * 4) uint32_t s_stack[(kPacketThreads / 32) * 2 * kPacketStack];
    if (!resolve_scene<false>(P)) return;
    uint32_t count = P.ray_count;
    if (P.indirect) count = min(count, __ldg(P.indirect));
    const uint32_t lane  = threadIdx.x & 31;
    const uint32_t stack = (uint32_t)__cvta_generic_to_shared(s_stack + (threadIdx.x >> 5) * 2 * kPacketStack);
    // Packets need the builder's leaf flags and an internal root: a single-triangle geometry, or a node array that was not written
    // by this library's builder (no tag in the root's update word), goes to the per-ray kernel chunk by chunk.
    const uint32_t* root_words = reinterpret_cast<const uint32_t*>(P.bvh);
    const uint32_t root_tag = __ldg(root_words + 15);
    const bool packets_ok = __ldg(root_words + 3) != kInvalid && (root_tag & kNodeTagMask) == kNodeTag;
    const bool static_order = (root_tag & kNodeVoteOrder) == 0;   // a treelet-optimised tree (rr_internal.h)
    if (P.grid && __ldg(P.grid + 2))
    {   // k_detect_grid found the batch incoherent: every chunk goes to the per-ray kernel, which then runs as if there were no list
        if (blockIdx.x == 0 && threadIdx.x == 0) *P.chunk_count = kAllChunks;
        return;
    }
    const RayGrid  G = load_grid(P.grid, count);
    const uint64_t packets = G.w ? (uint64_t)G.tiles_x * G.tiles_y : ((uint64_t)count + 63) / 64;
    while (true)
    {
        uint32_t packet = 0;
        if (lane == 0) packet = atomicAdd(P.packet_ticket, 1u);
        packet = __shfl_sync(0xffffffffu, packet, 0);
        if (packet >= packets) break;
        uint32_t gidx[2];
        bool     valid[2];
        float4   r0[2], r1[2];
        RayState ray[2];
        int      oct[2];
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            if (G.w) valid[k] = grid_ray(G, packet, k, lane, count, gidx[k]);   // an 8 x 8 tile of the ray grid: 8 x 4 per slot
            else
            {   // 64 consecutive rays
                gidx[k]  = packet * 64 + k * 32 + lane;
                valid[k] = gidx[k] < count;
            }
            const uint32_t ridx = valid[k] ? gidx[k] : count - 1;
            r0[k] = __ldcs(P.rays + 2 * (size_t)ridx);
            r1[k] = __ldcs(P.rays + 2 * (size_t)ridx + 1);
            ray[k].set(v3(r0[k]), v3(r1[k]));
            oct[k] = ray[k].octant();
        }
        const uint32_t gstep = G.w ? 4 * G.w : 32u;                                      // gidx[1] - gidx[0]
        const uint32_t slots = G.w || (uint64_t)packet * 64 + 32 < count ? 3u : 1u;   // the last strip may hold one chunk only
        const int      oct0     = __shfl_sync(0xffffffffu, oct[0], 0);
        const bool     coherent = packets_ok && oct0 != 8 && __all_sync(0xffffffffu, oct[0] == oct0 && oct[1] == oct0);
        bool done = false;
        if (coherent)
        {
            if (static_order)
            {
                switch (oct0)
                {
                case 0: done = trace_packet<kFullHit, 0, true>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 1: done = trace_packet<kFullHit, 1, true>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 2: done = trace_packet<kFullHit, 2, true>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 3: done = trace_packet<kFullHit, 3, true>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 4: done = trace_packet<kFullHit, 4, true>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 5: done = trace_packet<kFullHit, 5, true>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 6: done = trace_packet<kFullHit, 6, true>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                default: done = trace_packet<kFullHit, 7, true>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                }
            }
            else
            {
                switch (oct0)
                {
                case 0: done = trace_packet<kFullHit, 0, false>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 1: done = trace_packet<kFullHit, 1, false>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 2: done = trace_packet<kFullHit, 2, false>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 3: done = trace_packet<kFullHit, 3, false>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 4: done = trace_packet<kFullHit, 4, false>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 5: done = trace_packet<kFullHit, 5, false>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                case 6: done = trace_packet<kFullHit, 6, false>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                default: done = trace_packet<kFullHit, 7, false>(P, stack, gidx[0], gstep, valid, r0, r1, ray); break;
                }
            }
        }
        if (!done && lane == 0)
        {   // hand both chunks to the per-ray kernel
            const uint32_t n = slots == 3u ? 2u : 1u;
            const uint32_t at = atomicAdd(P.chunk_count, n);
            P.chunk_list[at] = packet * 2;
            if (n == 2u) P.chunk_list[at + 1] = packet * 2 + 1;
        }
    }
}

// ---- on-device ray binning (RR_CUDA_OPTION_SORT_RAYS) ------------------------------------------------------------------------------
// Incoherent batches (diffuse bounces) make the 32 rays of a warp walk 32 different paths: every node fetch is a divergent load and
// the warp runs as long as its slowest ray.  With the option set, every rrCmdIntersect first computes a 30-bit key per ray --
// direction octant (3 bits) | 6-bit-per-axis Morton cell of the origin inside the root box (18) | 3 bits per axis of the normalised
// |direction| (9) -- sorts (key, ray index) with the builder's onesweep sort, and k_trace then takes its 32-ray chunks from the
// sorted sequence.  (RR_CUDA_SORT_KEY_MODE=0: 24-bit keys, 5-bit cells, 2-bit directions, three passes: 2 132 against 2 184 Mrays/s
// on the C3 diffuse batch; =2: direction before cell: 2 131.)  A ray is still traced by one lane with the reference's visit order and writes its hit at its own
// index, so results are bit-identical and in the client's order.  No reference counterpart (the reference traces in buffer order).
__device__ __forceinline__ uint32_t spread5(uint32_t v)
{   // 5 bits -> every third bit
    v = (v | (v << 8)) & 0x0000100Fu;
    v = (v | (v << 4)) & 0x000010C3u;
    v = (v | (v << 2)) & 0x00001249u;
    return v;
}
__device__ __forceinline__ uint32_t spread6(uint32_t v)
{   // 6 bits -> every third bit
    v = (v | (v << 8)) & 0x0000300Fu;
    v = (v | (v << 4)) & 0x000030C3u;
    v = (v | (v << 2)) & 0x00009249u;
    return v;
}
__global__ void __launch_bounds__(256) k_ray_keys(const void* scene, const float4* __restrict__ rays, uint32_t ray_count,
                                                   const uint32_t* __restrict__ indirect, uint32_t* __restrict__ keys, int mode)
{
    // root box: node 0 of the geometry, or of the TLAS when the buffer is a scene
    const uint4 m     = __ldg(reinterpret_cast<const uint4*>(scene));
    const char* base  = reinterpret_cast<const char*>(scene);
    if (m.x == kSceneMagic0 && m.y == kSceneMagic1 && m.z == kSceneMagic2 && m.w == kSceneMagic3)
        base += reinterpret_cast<const SceneHeader*>(scene)->nodes_off;
    const float4* root = reinterpret_cast<const float4*>(base);
    const float4  q0 = __ldg(root), q1 = __ldg(root + 1), q2 = __ldg(root + 2), q3 = __ldg(root + 3);
    float lo[3], hi[3];
    if (wbits(q0) != kInvalid)
    {
        lo[0] = fminf(q0.x, q2.x); lo[1] = fminf(q0.y, q2.y); lo[2] = fminf(q0.z, q2.z);
        hi[0] = fmaxf(q1.x, q3.x); hi[1] = fmaxf(q1.y, q3.y); hi[2] = fmaxf(q1.z, q3.z);
    }
    else
    {   // single leaf: its vertices (or its box, for a one-instance TLAS)
        lo[0] = fminf(q0.x, fminf(q1.x, q2.x)); lo[1] = fminf(q0.y, fminf(q1.y, q2.y)); lo[2] = fminf(q0.z, fminf(q1.z, q2.z));
        hi[0] = fmaxf(q0.x, fmaxf(q1.x, q2.x)); hi[1] = fmaxf(q0.y, fmaxf(q1.y, q2.y)); hi[2] = fmaxf(q0.z, fmaxf(q1.z, q2.z));
    }
    uint32_t count = ray_count;
    if (indirect) count = min(count, __ldg(indirect));
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ray_count; i += gridDim.x * blockDim.x)
    {
        uint32_t key = mode ? 0x3FFFFFFFu : 0x00FFFFFFu;  // rays past the device-side count sort last (ties keep index order: the sort is stable)
        if (i < count)
        {
            const float4 r0 = __ldg(rays + 2 * (size_t)i), r1 = __ldg(rays + 2 * (size_t)i + 1);
            const float  o[3] = {r0.x, r0.y, r0.z}, d[3] = {r1.x, r1.y, r1.z};
            const float  len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            uint32_t cell[3], dq = 0, oct = 0;
#pragma unroll
            for (int a = 0; a < 3; ++a)
            {
                const float ext = hi[a] - lo[a];
                const float t   = ext > 0.0f ? (o[a] - lo[a]) / ext * 32.0f : 0.0f;
                cell[a] = (uint32_t)fminf(fmaxf(t, 0.0f), 31.0f);                       // (NaN -> 0)
                const float q = len > 0.0f ? fabsf(d[a]) / len * 4.0f : 0.0f;
                dq |= (uint32_t)fminf(fmaxf(q, 0.0f), 3.0f) << (2 * a);
                oct |= (__float_as_uint(d[a]) >> 31) << a;
            }
            key = (oct << 21) | (((spread5(cell[0]) << 2) | (spread5(cell[1]) << 1) | spread5(cell[2])) << 6) | dq;
            if (mode)
            {   // experiment: 6-bit cells, 3 bits per axis of direction: 30-bit keys, four passes
                uint32_t c6[3], d3 = 0;
#pragma unroll
                for (int a = 0; a < 3; ++a)
                {
                    const float ext = hi[a] - lo[a];
                    const float t   = ext > 0.0f ? (o[a] - lo[a]) / ext * 64.0f : 0.0f;
                    c6[a] = (uint32_t)fminf(fmaxf(t, 0.0f), 63.0f);
                    const float q = len > 0.0f ? fabsf(d[a]) / len * 8.0f : 0.0f;
                    d3 |= (uint32_t)fminf(fmaxf(q, 0.0f), 7.0f) << (3 * a);
                }
                const uint32_t mc = (spread6(c6[0]) << 2) | (spread6(c6[1]) << 1) | spread6(c6[2]);
                key = mode == 1 ? (oct << 27) | (mc << 9) | d3 : (oct << 27) | (d3 << 18) | mc;
            }
        }
        keys[i] = key;
    }
}

// Second launch of every intersect: traces the (normally zero) rays whose deferred-node stack outgrew shared memory,
// with the generic loop and a global-memory stack.  Exits at once when the list is empty.
template <bool kAny, bool kFullHit, bool kTwoLevel>
__global__ void __launch_bounds__(kTraceThreads) k_trace_deep(TraceParams P)
{
    const uint32_t count = *P.overflow_count;
    if (count == 0) return;
    if (!resolve_scene<kTwoLevel>(P)) return;
    const uint32_t lane = threadIdx.x & 31;
    DeepStack st;
    st.base   = P.arena + (blockIdx.x * kTraceThreads + threadIdx.x);
    st.stride = gridDim.x * kTraceThreads;
    st.error  = P.error;
    while (true)
    {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(P.deep_ticket, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if ((uint64_t)chunk * 32 >= count) break;
        const uint32_t i     = chunk * 32 + lane;
        const bool     valid = i < count;
        const uint32_t gidx  = P.overflow_list[valid ? i : count - 1];
        const float4 r0 = __ldg(P.rays + 2 * (size_t)gidx), r1 = __ldg(P.rays + 2 * (size_t)gidx + 1);
        RayState ray;
        ray.set(v3(r0), v3(r1));
        trace_ray<kAny, kFullHit, kTwoLevel, 8>(P, st, gidx, valid, r0, r1, ray);
    }
}

// Resident CTAs per SM the persistent grid is sized for (one-level kernels fit 10 at 48 registers, two-level 8).
// RR_CUDA_TRACE_CTAS_PER_SM overrides it for tuning.
inline int ctas_per_sm(bool two_level)
{
    static int env = [] { const char* e = std::getenv("RR_CUDA_TRACE_CTAS_PER_SM"); return e ? std::atoi(e) : 0; }();
    const int v = env > 0 ? env : (two_level ? 8 : 10);
    return std::min(v, 16);
}
inline int trace_grid(const DeviceInfo& dev, uint32_t ray_count, int per_sm)
{
    const size_t need = ((size_t)ray_count + kTraceThreads - 1) / kTraceThreads;
    return (int)std::max<size_t>(1, std::min<size_t>(need, (size_t)dev.sm_count * per_sm));
}

constexpr int kDeepCtasPerSm = 2;

template <bool kAny, bool kFullHit, bool kTwoLevel, bool kList = false>
void launch(const DeviceInfo& dev, cudaStream_t s, const TraceParams& P)
{
    k_trace<kAny, kFullHit, kTwoLevel, kList><<<trace_grid(dev, P.ray_count, ctas_per_sm(kTwoLevel)), kTraceThreads, 0, s>>>(P);
    k_trace_deep<kAny, kFullHit, kTwoLevel><<<trace_grid(dev, P.ray_count, kDeepCtasPerSm), kTraceThreads, 0, s>>>(P);
}
}  // namespace

// Scratch: [256 B header: main ticket, overflow count, deep ticket, packet ticket, declined-chunk count, ray-grid row length and phase | overflow list: 4 B per
// ray | declined-chunk list: 12 B per 32 rays | deep-kernel stacks: kDeepStack words for each of its thread slots].  The reference
// asks for 256 B per ray (vlk/geometry_trace.cpp:169).
constexpr size_t kScratchHeader = 256;
static size_t overflow_list_bytes(uint32_t ray_count) { return align_up(sizeof(uint32_t) * (size_t)ray_count, 256); }
// (three entries per 32 rays: an 8 x 8 tiling of a ray grid has more halves than the batch has 32-ray chunks when tiles stick out)
static size_t chunk_list_entries(uint32_t ray_count) { return 3 * (((size_t)ray_count + 31) / 32) + 64; }
static size_t chunk_list_bytes(uint32_t ray_count) { return align_up(sizeof(uint32_t) * chunk_list_entries(ray_count), 256); }
static size_t deep_arena_bytes(const DeviceInfo& dev, uint32_t ray_count)
{
    return align_up((size_t)trace_grid(dev, ray_count, kDeepCtasPerSm) * kTraceThreads * kDeepStack * sizeof(uint32_t), 256);
}
// With RR_CUDA_OPTION_SORT_RAYS: + [keys 4 B | permutation 4 B per ray | scratch of the radix sort]
static size_t ray_sort_bytes(uint32_t ray_count) { return 2 * align_up(sizeof(uint32_t) * (size_t)ray_count, 256) + sort_layout(ray_count).total; }
size_t trace_scratch_size(const DeviceInfo& dev, uint32_t ray_count)
{
    return kScratchHeader + overflow_list_bytes(ray_count) + chunk_list_bytes(ray_count) + deep_arena_bytes(dev, ray_count) +
           (dev.sort_rays ? ray_sort_bytes(ray_count) : 0);
}

// Resident CTAs per SM of the packet kernel (RR_CUDA_PACKET_CTAS_PER_SM overrides it for tuning).
static int packet_ctas_per_sm()
{
    static int env = [] { const char* e = std::getenv("RR_CUDA_PACKET_CTAS_PER_SM"); return e ? std::atoi(e) : 0; }();
    return std::min(env > 0 ? env : 8, 16);
}

void trace(const DeviceInfo& dev, cudaStream_t s, const TraceArgs& a)
{
    if (a.ray_count == 0) return;
    if (a.scratch_bytes < trace_scratch_size(dev, a.ray_count)) throw std::runtime_error("trace scratch buffer too small");
    TraceParams P;
    P.bvh = static_cast<const Node*>(a.scene); P.instances = nullptr; P.rays = reinterpret_cast<const float4*>(a.rays); P.ray_count = a.ray_count;
    P.indirect = a.indirect_count; P.hits = a.hits; P.first_found = a.first_found_tie_rule ? 1 : 0;
    P.ticket = a.scratch; P.overflow_count = a.scratch + 1; P.deep_ticket = a.scratch + 2;
    P.packet_ticket = a.scratch + 3; P.chunk_count = a.scratch + 4;
    P.overflow_list = a.scratch + kScratchHeader / sizeof(uint32_t);
    uint32_t* chunk_list = P.overflow_list + overflow_list_bytes(a.ray_count) / sizeof(uint32_t);
    P.arena = chunk_list + chunk_list_bytes(a.ray_count) / sizeof(uint32_t);
    P.chunk_list = nullptr;
    P.grid = nullptr;
    P.perm = nullptr;
    P.error = dev.error_word;
    static const int force_generic = [] { const char* e = std::getenv("RR_CUDA_TRACE_GENERIC"); return e ? std::atoi(e) : 0; }();
    static const int no_packets    = [] { const char* e = std::getenv("RR_CUDA_TRACE_PACKETS"); return e && std::atoi(e) == 0 ? 1 : 0; }();
    P.force_generic = force_generic;
    // Experiment switch (north_star: "L2-persistence staging of top BVH levels"): RR_CUDA_L2_WINDOW_MB=<n> marks the first n MB
    // of the traced buffer as persisting in L2 for the kernels of this stream.  Measured on C2 / C3 / C4
    // (profiles/round2_summary.md): no gain -- the 33.6 MB Sponza BVH already stays resident in the 126 MB L2 because rays and
    // hits stream through with evict-first loads / stores -- so it is off by default.
    static const int l2_window_mb = [] { const char* e = std::getenv("RR_CUDA_L2_WINDOW_MB"); return e ? std::atoi(e) : 0; }();
    if (l2_window_mb > 0)
    {
        static std::once_flag limit_once[kMaxDevices];
        std::call_once(limit_once[dev.device % kMaxDevices], [&] { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)l2_window_mb << 20); });
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr  = const_cast<void*>(a.scene);
        attr.accessPolicyWindow.num_bytes = (size_t)l2_window_mb << 20;
        attr.accessPolicyWindow.hitRatio  = 1.0f;
        attr.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
        RR_CUDA_CHECK(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr));
    }
    RR_CUDA_CHECK(cudaMemsetAsync(a.scratch, 0, 8 * sizeof(uint32_t), s));
    const bool any = a.query == RR_INTERSECT_QUERY_ANY, full = a.output == RR_INTERSECT_QUERY_OUTPUT_FULL_HIT;
    if (dev.sort_rays)
    {   // bin the rays: keys -> 3-pass onesweep over (key, index) -> k_trace walks the permutation (no packets: binned neighbours
        // share an octant and a cell, not a path)
        char*            sb   = reinterpret_cast<char*>(P.arena) + deep_arena_bytes(dev, a.ray_count);
        uint32_t*        keys = reinterpret_cast<uint32_t*>(sb);
        uint32_t*        perm = reinterpret_cast<uint32_t*>(sb + align_up(sizeof(uint32_t) * (size_t)a.ray_count, 256));
        void*            ss   = sb + 2 * align_up(sizeof(uint32_t) * (size_t)a.ray_count, 256);
        const SortLayout SL   = sort_layout(a.ray_count);
        const int        grid = (int)std::min<size_t>(((size_t)a.ray_count + 255) / 256, (size_t)dev.sm_count * 8);
        static const int key_mode = [] { const char* e = std::getenv("RR_CUDA_SORT_KEY_MODE"); return e ? std::atoi(e) : 1; }();
        k_ray_keys<<<grid, 256, 0, s>>>(a.scene, P.rays, a.ray_count, a.indirect_count, keys, key_mode);
        *dev.launches += 1;
        sort_reset(dev, s, SL, ss);
        sort_histogram(dev, s, SL, ss, keys);
        // (the sorted keys are of no use afterwards: they land in the overflow list, which only k_trace writes, later)
        sort_pairs(dev, s, SL, ss, keys, nullptr, P.overflow_list, perm, key_mode ? 4 : 3);
        P.perm = perm;
    }
    // (the per-ray kernels take 8 x 4 half tiles too: +3 % on image-ordered any-hit / first-found batches; RR_CUDA_TRACE_TILES=0: off)
    static const int trace_tiles = [] { const char* e = std::getenv("RR_CUDA_TRACE_TILES"); return e ? std::atoi(e) : 1; }();
    const bool packets = !any && !a.first_found_tie_rule && !no_packets && !dev.sort_rays;
    if (dev.ray_grid_width != 1 && !dev.sort_rays && (packets || trace_tiles))
    {   // 8 x 8 tiles of the ray grid instead of 64 x 1 strips when the batch is an image in row order (k_detect_grid)
        k_detect_grid<<<1, 1024, 0, s>>>(P.rays, a.ray_count, a.indirect_count, a.scratch + 5, dev.ray_grid_width,
                                         (uint32_t)std::min<size_t>(chunk_list_entries(a.ray_count), 0xFFFFFFFFu));
        *dev.launches += 1;
        P.grid = a.scratch + 5;
    }
    // One-level kernels (they return at once when the buffer turns out to be a scene) ...
    {
        TraceParams Q = P;
        if (packets)
        {   // closest hit under the (t, prim) rule: coherent 64-ray packets first, whatever they decline goes to k_trace in list mode
            Q.chunk_list = chunk_list;
            // (a tiling has at most 3 / 2 as many packets as the batch has strips, or k_detect_grid refuses it)
            const size_t need = (3 * (((size_t)a.ray_count + 63) / 64) / 2 + 3) / 4;
            const int    grid = (int)std::max<size_t>(1, std::min<size_t>(need, (size_t)dev.sm_count * packet_ctas_per_sm()));
            if (full) k_trace_packet<true><<<grid, kPacketThreads, 0, s>>>(Q);
            else      k_trace_packet<false><<<grid, kPacketThreads, 0, s>>>(Q);
            *dev.launches += 1;
        }
        if (any) { if (full) launch<true, true, false>(dev, s, Q); else launch<true, false, false>(dev, s, Q); }
        else if (Q.chunk_list) { if (full) launch<false, true, false, true>(dev, s, Q); else launch<false, false, false, true>(dev, s, Q); }
        else     { if (full) launch<false, true, false>(dev, s, Q); else launch<false, false, false>(dev, s, Q); }
    }
    // ... and two-level kernels (they return at once when it is a geometry).
    if (any) { if (full) launch<true, true, true>(dev, s, P); else launch<true, false, true>(dev, s, P); }
    else     { if (full) launch<false, true, true>(dev, s, P); else launch<false, false, true>(dev, s, P); }
    *dev.launches += 2;
    *dev.launches += 2;
    RR_CUDA_CHECK(cudaGetLastError());
}

}  // namespace rr
