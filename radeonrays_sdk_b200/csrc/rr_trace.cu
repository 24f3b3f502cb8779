// rr_trace.cu -- closest-hit / any-hit BVH2 traversal, one level (geometry) and two level (scene).
//
// Restates vlk/kernels/isect.comp:88-246 and isect_2l.comp:105-323 (+ common.h:103-209): one ray per
// thread, while-while traversal, near child first (`c1first = hit1 && t0_0 > t0_1`), the other child
// deferred on a LIFO stack, slab test with explicit fma, Moller-Trumbore with the shader's evaluation
// order (library is built with --fmad=false so nothing else is contracted).  The visit order is the
// reference's, which is what makes ANY-hit ids and closest-hit ties reproducible.
//
// B200 mapping:
//  * persistent CTAs (a multiple of the SM count); each warp strides over 32-ray chunks, so per-thread state
//    -- and therefore the stack arena -- is bounded by resident threads, not by ray_count (the reference asks
//    for 256 B of global stack per ray: 4 GiB for a 16 Mi batch, vlk/geometry_trace.cpp:169);
//  * traversal stack: kSmemStack entries per thread in shared memory laid out [level][thread] (bank = lane,
//    conflict free), deeper entries spill to the client's scratch buffer with the same layout (coalesced);
//  * nodes are fetched as 4 x 16-byte read-only loads (ld.global.nc.v4) of the 64-byte aligned node.
// Roofline: compulsory HBM traffic is 32 B/ray in + 16 B (or 4 B) out; the BVH is L2 resident, so the kernel
// is bound by L1/L2 latency and issue rate, not HBM.
#include <algorithm>
#include <cstdlib>

#include "rr_internal.h"

namespace rr
{
namespace
{
constexpr int kTraceThreads = 128;
constexpr int kSmemStack    = 32;   // entries per thread kept in shared memory
constexpr int kSpillStack   = 96;   // further entries per thread in the scratch arena
constexpr int kCtasPerSm    = 12;  // upper bound on resident CTAs per SM used for sizing the grid / spill arena

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

struct RayState
{
    Vec3  o, d, inv, oxinv;
    __device__ __forceinline__ void set(Vec3 oo, Vec3 dd)
    {
        o = oo; d = dd;
        inv   = v3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
        oxinv = v3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    }
};

// fast_intersect_aabb, common.h:150-164
__device__ __forceinline__ void slab(float4 bmin, float4 bmax, const RayState& r, float t_max, float t_min, float& t0, float& t1)
{
    const float fx = __fmaf_rn(bmax.x, r.inv.x, r.oxinv.x), fy = __fmaf_rn(bmax.y, r.inv.y, r.oxinv.y), fz = __fmaf_rn(bmax.z, r.inv.z, r.oxinv.z);
    const float nx = __fmaf_rn(bmin.x, r.inv.x, r.oxinv.x), ny = __fmaf_rn(bmin.y, r.inv.y, r.oxinv.y), nz = __fmaf_rn(bmin.z, r.inv.z, r.oxinv.z);
    const float ax = fmaxf(fx, nx), ay = fmaxf(fy, ny), az = fmaxf(fz, nz);
    const float ix = fminf(fx, nx), iy = fminf(fy, ny), iz = fminf(fz, nz);
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

struct Stack
{
    uint32_t* smem;    // &s_stack[tid]
    uint32_t* spill;   // &arena[slot]
    uint32_t  spill_stride;
    int       sp;
    __device__ __forceinline__ void push(uint32_t v)
    {
        if (sp < kSmemStack) smem[sp * kTraceThreads] = v;
        else if (sp < kSmemStack + kSpillStack) spill[(size_t)(sp - kSmemStack) * spill_stride] = v;
        else return;  // deeper than any tree this builder can produce; drop rather than corrupt
        ++sp;
    }
    __device__ __forceinline__ uint32_t pop()
    {
        --sp;
        return sp < kSmemStack ? smem[sp * kTraceThreads] : spill[(size_t)(sp - kSmemStack) * spill_stride];
    }
};

struct TraceParams
{
    const Node*           bvh;
    const InstanceRecord* instances;
    const float4*         rays;
    uint32_t              ray_count;
    const uint32_t*       indirect;
    void*                 hits;
    uint32_t*             arena;
    uint32_t*             ticket;   // chunk counter, first word of the scratch buffer
    int                   first_found;
};

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
template <bool kAny, bool kFullHit, bool kTwoLevel>
__device__ __forceinline__ void trace_ray(const TraceParams& P, Stack& st, uint32_t gidx)
{
    const float4 r0 = __ldg(P.rays + 2 * (size_t)gidx), r1 = __ldg(P.rays + 2 * (size_t)gidx + 1);
    const float  min_t = r0.w;
    RayState ray;
    ray.set(v3(r0), v3(r1));
    float    closest      = r1.w;
    uint32_t closest_addr = kInvalid, closest_prim = kInvalid, closest_inst = kInvalid;
    uint32_t cur_inst     = kInvalid;
    const Node* cur_bvh   = P.bvh;
    st.sp = 0;
    st.push(kInvalid);
    uint32_t addr = 0;
    while (addr != kInvalid)
    {
        const float4* np = reinterpret_cast<const float4*>(cur_bvh + addr);
        // all four quads are requested at once: the fourth is dead for leaves, but issuing it after the
        // leaf/internal branch would put a second dependent load on every internal-node visit
        const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
        if (wbits(q0) != kInvalid)
        {
            float a0, a1, b0, b1;
            slab(q0, q1, ray, closest, min_t, a0, a1);
            slab(q2, q3, ray, closest, min_t, b0, b1);
            const bool t0 = a0 <= a1, t1 = b0 <= b1;
            if (t0 || t1)
            {
                const bool c1first = t1 && (a0 > b0);
                const uint32_t c0 = wbits(q0), c1 = wbits(q1);
                uint32_t deferred;
                if (c1first || !t0) { addr = c1; deferred = c0; }
                else { addr = c0; deferred = c1; }
                if (t0 && t1) st.push(deferred);
                continue;
            }
        }
        else if (kTwoLevel && cur_inst == kInvalid)
        {   // top-level leaf: enter the instance (isect_2l.comp:231-245)
            cur_inst = wbits(q1);
            const InstanceRecord* rec = P.instances + cur_inst;
            Vec3 oo, od;
            transform_ray(rec, ray.o, ray.d, oo, od);
            ray.set(oo, od);
            cur_bvh = rec->blas;
            st.push(kSentinel);
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
                            reinterpret_cast<float4*>(P.hits)[gidx] =
                                make_float4(uv.x, uv.y, __uint_as_float(kTwoLevel ? cur_inst : 0u), __uint_as_float(prim));
                        }
                        else
                            reinterpret_cast<uint32_t*>(P.hits)[gidx] = kTwoLevel ? cur_inst : prim;
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
            reinterpret_cast<float4*>(P.hits)[gidx] =
                make_float4(uv.x, uv.y, __uint_as_float(kTwoLevel ? closest_inst : 0u), __uint_as_float(closest_prim));
        }
        else
            reinterpret_cast<uint32_t*>(P.hits)[gidx] = kTwoLevel ? closest_inst : closest_prim;
    }
    else
    {   // miss: only the id word is written (isect.comp:238-245)
        if (kFullHit) reinterpret_cast<uint32_t*>(P.hits)[4 * (size_t)gidx + 2] = kInvalid;
        else reinterpret_cast<uint32_t*>(P.hits)[gidx] = kInvalid;
    }
}
#undef RR_POP_NEXT

// Persistent warps pull one 32-ray chunk at a time from a global ticket (P.ticket, zeroed by a memset node before
// the launch) and trace it to completion.  Against a static warp-stride assignment this removes the end-of-kernel
// imbalance: +29 % on coherent primary rays, +39 % on shadow rays (profiles/round1_trace_modes.md).
template <bool kAny, bool kFullHit, bool kTwoLevel>
__global__ void __launch_bounds__(kTraceThreads, kTwoLevel ? 8 : 10) k_trace(TraceParams P)
{
    __shared__ uint32_t s_stack[kSmemStack * kTraceThreads];
    uint32_t count = P.ray_count;
    if (P.indirect) count = min(count, __ldg(P.indirect));  // isect.comp:98-103
    const uint32_t lane = threadIdx.x & 31;
    Stack st;
    st.smem         = s_stack + threadIdx.x;
    st.spill        = P.arena + (blockIdx.x * kTraceThreads + threadIdx.x);
    st.spill_stride = gridDim.x * kTraceThreads;
    while (true)
    {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(P.ticket, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if ((uint64_t)chunk * 32 >= count) break;
        const uint32_t gidx = chunk * 32 + lane;
        if (gidx < count) trace_ray<kAny, kFullHit, kTwoLevel>(P, st, gidx);
    }
}

// Resident CTAs per SM the persistent grid is sized for (one-level kernels fit 10 at 48 registers, two-level 8).
// RR_CUDA_TRACE_CTAS_PER_SM overrides it for tuning.
inline int ctas_per_sm(bool two_level)
{
    static int env = [] { const char* e = std::getenv("RR_CUDA_TRACE_CTAS_PER_SM"); return e ? std::atoi(e) : 0; }();
    const int v = env > 0 ? env : (two_level ? 8 : 10);
    return std::min(v, kCtasPerSm);
}
inline int trace_grid(const DeviceInfo& dev, uint32_t ray_count, int per_sm)
{
    const size_t need = ((size_t)ray_count + kTraceThreads - 1) / kTraceThreads;
    return (int)std::max<size_t>(1, std::min<size_t>(need, (size_t)dev.sm_count * per_sm));
}

template <bool kAny, bool kFullHit, bool kTwoLevel>
void launch(const DeviceInfo& dev, cudaStream_t s, const TraceParams& P)
{
    k_trace<kAny, kFullHit, kTwoLevel><<<trace_grid(dev, P.ray_count, ctas_per_sm(kTwoLevel)), kTraceThreads, 0, s>>>(P);
}
}  // namespace

// Scratch: [256 B header holding the chunk ticket | spill arena: kSpillStack words for every resident thread slot]
// (never more slots than rays).
constexpr size_t kScratchHeader = 256;
size_t trace_scratch_size(const DeviceInfo& dev, uint32_t ray_count)
{
    return kScratchHeader + (size_t)trace_grid(dev, ray_count, kCtasPerSm) * kTraceThreads * kSpillStack * sizeof(uint32_t);
}

void trace(const DeviceInfo& dev, cudaStream_t s, const TraceArgs& a)
{
    if (a.ray_count == 0) return;
    if (a.scratch_bytes < trace_scratch_size(dev, a.ray_count)) throw std::runtime_error("trace scratch buffer too small");
    TraceParams P;
    P.bvh = a.bvh; P.instances = a.instances; P.rays = reinterpret_cast<const float4*>(a.rays); P.ray_count = a.ray_count;
    P.indirect = a.indirect_count; P.hits = a.hits; P.ticket = a.scratch; P.arena = a.scratch + kScratchHeader / sizeof(uint32_t); P.first_found = a.first_found_tie_rule ? 1 : 0;
    RR_CUDA_CHECK(cudaMemsetAsync(a.scratch, 0, sizeof(uint32_t), s));
    const bool any = a.query == RR_INTERSECT_QUERY_ANY, full = a.output == RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, two = a.instances != nullptr;
    if (!two)
    {
        if (any) { if (full) launch<true, true, false>(dev, s, P); else launch<true, false, false>(dev, s, P); }
        else     { if (full) launch<false, true, false>(dev, s, P); else launch<false, false, false>(dev, s, P); }
    }
    else
    {
        if (any) { if (full) launch<true, true, true>(dev, s, P); else launch<true, false, true>(dev, s, P); }
        else     { if (full) launch<false, true, true>(dev, s, P); else launch<false, false, true>(dev, s, P); }
    }
    ++*dev.launches;
    RR_CUDA_CHECK(cudaGetLastError());
}

}  // namespace rr
