// rr_treelet.cu -- treelet restructuring (Karras & Aila 2013) of a built BLAS, 3 rounds with
// min_prims = 64, 128, 256 exactly like the reference (vlk/restructure_hlbvh.cpp:151-277,
// init_primitive_count.comp, find_treelet_roots.comp, restructure_bvh.comp).
//
// B200 mapping: one WARP per treelet instead of the reference's 64-thread work-group with thread-0 serial
// sections and ~10 barriers per treelet: the 7 treelet leaves live in lanes 0..6, the 128-subset dynamic
// programme is spread over the 32 lanes level by level (popcount 2..7) with __syncwarp between levels,
// and the rebuilt 6 internal nodes are written as whole 64-byte nodes by lanes 0..5.
//
// Result parity (same tree as the reference's kernel, checked by tests/ against the CPU oracle):
//  * treelet growth uses the reference's selection rule verbatim (`largest == 0 || area > largest`,
//    restructure_bvh.comp:184-197);
//  * candidate partitions of a subset s are enumerated in the reference's order p_k = deposit(k+1, s minus
//    its lowest bit) (Algorithm 3 of the paper; restructure_bvh.comp:349, 394-407) and the FIRST minimum is
//    kept (strict `<` / `>` in the shader); |s|=7 looks at k = 0..61 only, as the shader's tree reduction
//    does (:361-372); |s|=6 replays the shader's 8-thread chunked walk literally, including its quirk that
//    chunk starts are computed without masking (:281-300); |s| in [2,5] stores p or s^p by popcount parity
//    (:412), |s| in {6,7} store p;
//  * areas are 2*dot(e, e.zxy), costs 1.2f*area + cost[p] + cost[s^p], all binary32 without contraction.
#include <algorithm>

#include "rr_internal.h"

namespace rr
{
namespace
{
constexpr int   kTreeletWarps = 4;
constexpr float kCInt         = 1.2f;

__device__ __forceinline__ float3 xyz(float4 q) { return make_float3(q.x, q.y, q.z); }
__device__ __forceinline__ uint32_t wbits(float4 q) { return __float_as_uint(q.w); }
__device__ __forceinline__ float4 pack(float3 v, uint32_t w) { return make_float4(v.x, v.y, v.z, __uint_as_float(w)); }
__device__ __forceinline__ float3 min3(float3 a, float3 b) { return make_float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__device__ __forceinline__ float3 max3(float3 a, float3 b) { return make_float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }

// GetAabbSurfaceArea, restructure_bvh.comp:90-94: 2 * dot(e, e.zxy)
__device__ __forceinline__ float box_area(float3 lo, float3 hi)
{
    const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
    return 2.0f * ((ex * ez + ey * ex) + ez * ey);
}
// GetNodeAabb, restructure_bvh.comp:71-88: 3 points for a leaf, 4 for an internal node, grown from +-FLT_MAX.
__device__ __forceinline__ void treelet_node_box(const Node* nodes, uint32_t addr, uint32_t leaf0, float3& lo, float3& hi)
{
    const float4* p  = reinterpret_cast<const float4*>(nodes + addr);
    const float4  q0 = __ldcg(p), q1 = __ldcg(p + 1), q2 = __ldcg(p + 2);
    lo = make_float3(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f);
    hi = make_float3(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
    lo = min3(lo, xyz(q0)); hi = max3(hi, xyz(q0));
    lo = min3(lo, xyz(q1)); hi = max3(hi, xyz(q1));
    lo = min3(lo, xyz(q2)); hi = max3(hi, xyz(q2));
    if (addr < leaf0)
    {
        const float4 q3 = __ldcg(p + 3);
        lo = min3(lo, xyz(q3)); hi = max3(hi, xyz(q3));
    }
}

// k-th candidate partition of subset s in the reference's enumeration order: bits of (k+1) deposited into
// delta = s without its lowest set bit (the lowest leaf always stays on the other side).
__device__ __forceinline__ uint32_t partition_k(uint32_t s, uint32_t k)
{
    uint32_t delta = (s - 1u) & s, v = k + 1u, out = 0;
    while (delta)
    {
        const uint32_t low = delta & (0u - delta);
        if (v & 1u) out |= low;
        v >>= 1;
        delta ^= low;
    }
    return out;
}

// find_treelet_roots.comp:63-98
__global__ void __launch_bounds__(256)
    k_find_treelet_roots(const Node* __restrict__ nodes, uint32_t n, uint32_t min_prims, uint32_t* __restrict__ counts,
                         uint32_t* __restrict__ root_count, uint32_t* __restrict__ roots)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t prim  = 1;
    uint32_t index = reinterpret_cast<const uint32_t*>(nodes + (n - 1 + j))[11];
    while (index != kInvalid)
    {
        const uint32_t old = atomicExch(&counts[index], prim);
        prim += old;
        if (old == 0) break;  // first arrival
        if (prim >= min_prims)
        {
            roots[atomicAdd(root_count, 1u)] = index;
            break;
        }
        index = reinterpret_cast<const uint32_t*>(nodes + index)[11];
    }
}

struct WarpScratch
{
    float    area[128];
    float    cost[128];
    float3   lo[7], hi[7];
    uint8_t  part[128];
    // rebuilt internal nodes: [slot] = node index, subset, left subset, parent
    uint32_t in_node[6], in_parent[6];
    uint8_t  in_mask[6], in_left[6];
    uint32_t leaf_parent[7];
};

__global__ void __launch_bounds__(kTreeletWarps * 32)
    k_restructure(Node* __restrict__ nodes, uint32_t n, uint32_t* __restrict__ counts, const uint32_t* __restrict__ root_count,
                  const uint32_t* __restrict__ roots)
{
    __shared__ WarpScratch s_all[kTreeletWarps];
    // candidate partitions of every subset with 2..5 leaves, in the reference's enumeration order: computed once per CTA
    // (thread t fills the row of subset t) instead of bit-depositing k+1 for each of the ~700 candidates of every treelet
    __shared__ uint8_t s_part_table[128][16];
    {
        const uint32_t s = threadIdx.x, bits = __popc(s);
        if (bits >= 2 && bits <= 5)
            for (uint32_t k = 0; k < (1u << (bits - 1)) - 1u; ++k) s_part_table[s][k] = (uint8_t)partition_k(s, k);
    }
    __syncthreads();
    const int      lane  = threadIdx.x & 31;
    const uint32_t w     = blockIdx.x * kTreeletWarps + (threadIdx.x >> 5);
    if (w >= *root_count) return;
    WarpScratch&   S     = s_all[threadIdx.x >> 5];
    const uint32_t leaf0 = n - 1;
    uint32_t       node  = roots[w];

    while (true)
    {
        // ---- form the treelet: lanes 0..6 own the treelet leaves (restructure_bvh.comp:170-217) ----
        uint32_t my_leaf = kInvalid;  // node index held by this lane (lanes 0..6)
        float    my_area = 0.f;
        float3   my_lo = make_float3(0, 0, 0), my_hi = my_lo;
        uint32_t internal_nodes[6];   // replicated in every lane
        uint32_t internal_update[6];
        uint32_t root_parent;
        {
            const float4* rp = reinterpret_cast<const float4*>(nodes + node);
            const float4  q0 = __ldcg(rp), q1 = __ldcg(rp + 1), q2 = __ldcg(rp + 2), q3 = __ldcg(rp + 3);
            internal_nodes[0]  = node;
            internal_update[0] = wbits(q3);
            root_parent        = wbits(q2);
            // A node stores the boxes of its two children, and they are bit-equal to what GetNodeAabb (restructure_bvh.comp:71-88)
            // computes from the child itself -- the union of the child's own two stored boxes, or of its three vertices; min / max
            // are exact -- so the box of a new treelet leaf comes from the node just fetched instead of a second, dependent load
            // of the child (two L2 round trips per growth step were ~40 % of a treelet's latency, and the climb is serial).
            if (lane == 0) { my_leaf = wbits(q0); my_lo = xyz(q0); my_hi = xyz(q1); }
            if (lane == 1) { my_leaf = wbits(q1); my_lo = xyz(q2); my_hi = xyz(q3); }
            if (lane < 2) my_area = box_area(my_lo, my_hi);
        }
        for (int size = 2; size < 7; ++size)
        {
            float    largest = 0.0f;
            uint32_t pick = 0, slot = 0;
            for (int i = 0; i < size; ++i)
            {
                const uint32_t li = __shfl_sync(0xffffffffu, my_leaf, i);
                const float    ai = __shfl_sync(0xffffffffu, my_area, i);
                if (li < leaf0 && (largest == 0.0f || ai > largest)) { largest = ai; pick = li; slot = i; }
            }
            const float4* pp = reinterpret_cast<const float4*>(nodes + pick);
            const float4  q0 = __ldcg(pp), q1 = __ldcg(pp + 1), q2 = __ldcg(pp + 2), q3 = __ldcg(pp + 3);
            internal_nodes[size - 1]  = pick;
            internal_update[size - 1] = wbits(q3);
            if (lane == (int)slot) { my_leaf = wbits(q0); my_lo = xyz(q0); my_hi = xyz(q1); }
            if (lane == size) { my_leaf = wbits(q1); my_lo = xyz(q2); my_hi = xyz(q3); }
            if (lane == (int)slot || lane == size) my_area = box_area(my_lo, my_hi);
        }
        if (lane < 7) { S.lo[lane] = my_lo; S.hi[lane] = my_hi; }
        __syncwarp();

        // ---- subset areas (restructure_bvh.comp:223-240) and singleton costs (:249-252) ----
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            const uint32_t m = lane + 32 * q;
            float3 lo = make_float3(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f);
            float3 hi = make_float3(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
#pragma unroll
            for (int i = 0; i < 7; ++i)
                if (m & (1u << i)) { lo = min3(lo, S.lo[i]); hi = max3(hi, S.hi[i]); }
            const float a = box_area(lo, hi);
            S.area[m] = a;
            S.cost[m] = (__popc(m) == 1) ? (kCInt * a) / 1.0f : 0.0f;
            S.part[m] = 0;
        }
        __syncwarp();

        // ---- dynamic programme over subsets by popcount ----
        for (int bits = 2; bits <= 5; ++bits)
        {
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                const uint32_t s = lane + 32 * q;
                if (__popc(s) != bits) continue;
                const uint32_t K = (1u << (bits - 1)) - 1u;
                float    lowest = 3.402823466e+38f;
                uint32_t best = 0;
                for (uint32_t k = 0; k < K; ++k)
                {
                    const uint32_t p = s_part_table[s][k];
                    const float    c = S.cost[p] + S.cost[s ^ p];
                    if (best == 0 || c < lowest) { lowest = c; best = p; }
                }
                S.cost[s] = kCInt * S.area[s] + lowest;
                S.part[s] = (uint8_t)((bits & 1) ? best : (s ^ best));
            }
            __syncwarp();
        }
        {   // |s| = 6 (restructure_bvh.comp:264-338): the shader gives each of 7 subsets to 8 threads; thread t starts
            // at (p0 - delta*4t) & s and walks at most 4 partitions.  For subsets with a gap in their bit pattern that
            // start is NOT the 4t-th partition of the sequence, so some partitions are visited twice and others never:
            // we replay the shader's walk literally (4 lanes per subset, two shader threads per lane) and merge with
            // its tree reduction's rule (strict '>' : the lower thread index wins ties).
            const int      mi = lane >> 2, chunk = lane & 3;
            const uint32_t s  = mi < 7 ? (0x7fu ^ (1u << (6 - mi))) : 0x7fu;  // 0x3f,0x5f,0x6f,0x77,0x7b,0x7d,0x7e for mi = 0..6
            const uint32_t delta = (s - 1u) & s, p0 = (0u - delta) & s;
            float    lowest = 3.402823466e+38f;
            uint32_t best = 0;
            if (mi < 7)
            {
#pragma unroll
                for (uint32_t tt = 0; tt < 2; ++tt)
                {
                    const uint32_t t = 2u * chunk + tt;
                    uint32_t p = (p0 - delta * t * 4u) & s;
                    float    lo_t = 3.402823466e+38f;
                    uint32_t best_t = 0;
                    int      counter = 0;
                    do
                    {
                        const float c = S.cost[p] + S.cost[s ^ p];
                        if (best_t == 0 || c < lo_t) { lo_t = c; best_t = p; }
                        p = (p - delta) & s;
                        ++counter;
                    } while (p != 0 && counter < 4);
                    if (tt == 0 || lowest > lo_t) { lowest = lo_t; best = best_t; }
                }
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1)
            {
                const float    oc = __shfl_xor_sync(0xffffffffu, lowest, o);
                const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, o);
                const bool     lower = (chunk & o) == 0;
                if (lower ? (lowest > oc) : !(oc > lowest)) { lowest = oc; best = ob; }
            }
            if (mi < 7 && chunk == 0)
            {
                S.cost[s] = kCInt * S.area[s] + lowest;
                S.part[s] = (uint8_t)best;
            }
        }
        __syncwarp();
        {   // |s| = 7: candidates k = 0..61 (the reference's tree reduction never reads k = 62)
            const uint32_t s = 0x7fu;
            float    lowest = 3.402823466e+38f;
            uint32_t best_k = 0xFFFFu, best = 0;
            for (uint32_t k = 2 * lane; k < 62u && k < 2u * lane + 2; ++k)
            {
                const uint32_t p = (k + 1u) << 1;  // = partition_k(0x7f, k): k+1 deposited into bits 1..6
                const float    c = S.cost[p] + S.cost[s ^ p];
                if (best == 0 || c < lowest) { lowest = c; best = p; best_k = k; }
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const float    oc = __shfl_xor_sync(0xffffffffu, lowest, o);
                const uint32_t ok = __shfl_xor_sync(0xffffffffu, best_k, o);
                const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, o);
                if (oc < lowest || (oc == lowest && ok < best_k)) { lowest = oc; best_k = ok; best = ob; }
            }
            if (lane == 0)
            {
                S.cost[s] = kCInt * S.area[s] + lowest;
                S.part[s] = (uint8_t)best;
            }
        }
        __syncwarp();

        // ---- rebuild topology (restructure_bvh.comp:417-475): lane 0 replays the reference's stack walk ----
        if (lane == 0)
        {
            uint32_t st_mask[7], st_slot[7];
            uint32_t allocated = 1, sp = 1;
            st_mask[0] = 0x7f; st_slot[0] = 0;
            S.in_node[0] = internal_nodes[0]; S.in_parent[0] = root_parent;
            while (sp > 0)
            {
                --sp;
                const uint32_t pm = st_mask[sp], ps = st_slot[sp];
                const uint32_t lm = S.part[pm], rm = pm ^ lm;
                S.in_mask[ps] = (uint8_t)pm;
                S.in_left[ps] = (uint8_t)lm;
                const uint32_t pnode = S.in_node[ps];
                if (__popc(lm) > 1)
                {
                    S.in_node[allocated] = internal_nodes[allocated]; S.in_parent[allocated] = pnode;
                    st_mask[sp] = lm; st_slot[sp] = allocated; ++sp; ++allocated;
                }
                else S.leaf_parent[31 - __clz(lm)] = pnode;
                if (__popc(rm) > 1)
                {
                    S.in_node[allocated] = internal_nodes[allocated]; S.in_parent[allocated] = pnode;
                    st_mask[sp] = rm; st_slot[sp] = allocated; ++sp; ++allocated;
                }
                else S.leaf_parent[31 - __clz(rm)] = pnode;
            }
        }
        __syncwarp();
        // Which node represents subset m: a treelet leaf if |m| = 1, else the internal slot holding mask m.
        auto node_of = [&](uint32_t m) -> uint32_t {
            if (__popc(m) == 1) return __shfl_sync(0xffffffffu, my_leaf, 31 - __clz(m));
            uint32_t r = kInvalid;
            for (int i = 0; i < 6; ++i)
                if (S.in_mask[i] == m) r = S.in_node[i];
            return r;
        };
        {   // all lanes take part in the shuffles inside node_of; lanes 0..5 write one internal node each
            const int      slot = lane < 6 ? lane : 0;
            const uint32_t pm = S.in_mask[slot], lm = S.in_left[slot], rm = pm ^ lm;
            uint32_t c0 = kInvalid, c1 = kInvalid;
            // node_of shuffles need uniform participation: evaluate for every slot on every lane
            for (int sl = 0; sl < 6; ++sl)
            {
                const uint32_t a = node_of(S.in_left[sl]);
                const uint32_t b = node_of((uint32_t)S.in_mask[sl] ^ (uint32_t)S.in_left[sl]);
                if (sl == slot) { c0 = a; c1 = b; }
            }
            if (lane < 6)
            {
                float3 llo = make_float3(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f), lhi = make_float3(-llo.x, -llo.x, -llo.x);
                float3 rlo = llo, rhi = lhi;
#pragma unroll
                for (int i = 0; i < 7; ++i)
                {
                    if (lm & (1u << i)) { llo = min3(llo, S.lo[i]); lhi = max3(lhi, S.hi[i]); }
                    if (rm & (1u << i)) { rlo = min3(rlo, S.lo[i]); rhi = max3(rhi, S.hi[i]); }
                }
                float4* np = reinterpret_cast<float4*>(nodes + S.in_node[slot]);
                np[0] = pack(llo, c0);
                np[1] = pack(lhi, c1);
                np[2] = pack(rlo, S.in_parent[slot]);
                np[3] = pack(rhi, node_update_word(c0, c1, n - 1u, internal_update[slot], node_order_bits(llo, lhi, rlo, rhi)));  // flags follow the new children
            }
            if (lane < 7) reinterpret_cast<uint32_t*>(nodes + my_leaf)[11] = S.leaf_parent[lane];
        }

        // ---- climb (restructure_bvh.comp:497-541) ----
        uint32_t go = 0;
        __syncwarp();
        if (lane == 0 && root_parent != kInvalid)
        {
            __threadfence();
            const uint32_t old = atomicAdd(&counts[root_parent], 1u);
            if (old != 0) { __threadfence(); go = 1; }
        }
        go = __shfl_sync(0xffffffffu, go, 0);
        if (!go) return;
        node = root_parent;
    }
}
}  // namespace

// scratch: [root_count (256 B) | counts 4(2n-1) | roots 4(n/64+1)]   (reference: restructure_hlbvh.cpp:306-310)
size_t treelet_scratch_size(uint32_t n)
{
    return 256 + align_up(sizeof(uint32_t) * (2 * (size_t)n), 256) + align_up(sizeof(uint32_t) * ((size_t)n / 64 + 2), 256);
}

void restructure_blas(const DeviceInfo& dev, cudaStream_t s, Node* nodes, uint32_t n, void* scratch)
{
    if (n < 64) return;  // no subtree can reach min_prims: every round is a no-op (as in the reference)
    char*     sc         = (char*)scratch;
    uint32_t* root_count = reinterpret_cast<uint32_t*>(sc);
    uint32_t* counts     = reinterpret_cast<uint32_t*>(sc + 256);
    size_t    counts_sz  = align_up(sizeof(uint32_t) * (2 * (size_t)n), 256);
    uint32_t* roots      = reinterpret_cast<uint32_t*>(sc + 256 + counts_sz);
    const uint32_t rounds[3] = {64, 128, 256};
    for (uint32_t min_prims : rounds)
    {
        RR_CUDA_CHECK(cudaMemsetAsync(sc, 0, 256 + counts_sz, s));
        k_find_treelet_roots<<<(n + 255) / 256, 256, 0, s>>>(nodes, n, min_prims, counts, root_count, roots);
        const uint32_t max_roots = n / min_prims + 1;
        k_restructure<<<(max_roots + kTreeletWarps - 1) / kTreeletWarps, kTreeletWarps * 32, 0, s>>>(nodes, n, counts, root_count, roots);
        *dev.launches += 2;
    }
    // The nodes are re-linked: the tree is no longer the Karras tree of the deltas kept in the geometry buffer's tail, so an
    // update must use the generic refit (header word of the tail, see update_blas).
    RR_CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<char*>(nodes) + blas_layout(n, false).tail_off, 0, sizeof(uint32_t), s));
    RR_CUDA_CHECK(cudaGetLastError());
}

}  // namespace rr
