// rr_sort.cu -- stable LSD radix sort of (u32 key, u32 value) pairs, onesweep style.
//
// Replaces the reference's 8 x (histogram -> 1/3/5-dispatch scan -> scatter) pipeline
// (vlk/radix_sort.cpp:157-338, radix_sort_group_histograms.comp, radix_sort_scatter_keys.comp,
// scan_exclusive_add*.comp) with: one digit-histogram accumulation (fused into the Morton kernel on
// the build path) + 4 passes of 8 bits, each a single kernel that ranks a tile with warp match,
// resolves its global offsets by decoupled look-back and scatters through shared memory.
// Semantics are the reference's (vlk/radix_sort.cpp:202-215): stable, keys compared as unsigned bit
// patterns, so equal keys keep ascending input order.
//
// HBM traffic per pair: histogram read 4 B (0 when fused) + 4 passes x (8 B read + 8 B write), minus the
// 4-B value read of pass 0 when values are the identity.  Roofline: HBM bandwidth.
#include <algorithm>
#include <mutex>

#include "rr_internal.h"

namespace rr
{
namespace
{
constexpr int      kRadixBits   = 8;
constexpr int      kRadix       = 1 << kRadixBits;
constexpr int      kPasses      = 4;
constexpr int      kSortThreads = 256;                      // one thread per digit in the look-back
constexpr int      kSortWarps   = kSortThreads / 32;
constexpr int      kSortIpt     = 16;
constexpr int      kSortTile    = kSortThreads * kSortIpt;  // 4096 pairs per tile
constexpr uint32_t kFlagAgg     = 1u << 30;                 // tile aggregate available
constexpr uint32_t kFlagPrefix  = 2u << 30;                 // inclusive prefix available
constexpr uint32_t kValueMask   = (1u << 30) - 1;
constexpr int      kLookAhead   = 4;                        // predecessor tiles polled per look-back step

constexpr size_t kSortSmemBytes = sizeof(uint32_t) * (kSortWarps * kRadix + 2 * kSortTile + 2 * kRadix + 16);

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// Look-back words are single-word messages (flag + count in one u32): relaxed gpu-scope accesses suffice.
__device__ __forceinline__ uint32_t ld_status(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Lanes of the warp holding the same 8-bit digit.  Eight ballots instead of MATCH.ANY: on sm_100 the match
// instruction serialises over the distinct values in the warp and was 45 % of this kernel's stall samples
// (profiles/round1_build.md); ballots are fixed cost.
__device__ __forceinline__ uint32_t match_digit(uint32_t d)
{
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < kRadixBits; ++b)
    {
        const bool     bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// Exclusive scan of one value per thread over the 256 threads of the CTA.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* s_part /*>=8*/, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t  inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_part[warp] = inc;
    __syncthreads();
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w)
        if (w < warp) add += s_part[w];
    __syncthreads();
    return inc - v + add;
}

__global__ void __launch_bounds__(256) k_sort_histogram(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_hist[kPasses * kRadix];
    for (int i = threadIdx.x; i < kPasses * kRadix; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const uint32_t k = keys[i];
#pragma unroll
        for (int p = 0; p < kPasses; ++p) atomicAdd(&s_hist[p * kRadix + ((k >> (p * kRadixBits)) & (kRadix - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPasses * kRadix; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// One onesweep pass.  grid = number of tiles; tiles are claimed through an atomic ticket so that a
// tile's predecessors are always resident or finished (forward progress of the look-back).
template <bool kIdentityValues>
__global__ void __launch_bounds__(kSortThreads, 5)
    k_onesweep_pass(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                    uint32_t* __restrict__ vals_out, uint32_t n, int shift, const uint32_t* __restrict__ pass_hist,
                    uint32_t* __restrict__ ticket, uint32_t* __restrict__ status)
{
    extern __shared__ uint32_t smem[];
    uint32_t* s_warp_hist   = smem;                               // [kSortWarps][kRadix]
    uint32_t* s_keys        = s_warp_hist + kSortWarps * kRadix;  // [kSortTile]
    uint32_t* s_vals        = s_keys + kSortTile;                 // [kSortTile]
    uint32_t* s_digit_base  = s_vals + kSortTile;                 // [kRadix]
    uint32_t* s_global_base = s_digit_base + kRadix;              // [kRadix]
    uint32_t* s_misc        = s_global_base + kRadix;             // [16]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_misc[8] = atomicAdd(ticket, 1u);
#pragma unroll
    for (int i = 0; i < kSortWarps; ++i) s_warp_hist[i * kRadix + tid] = 0;
    __syncthreads();
    const uint32_t tile  = s_misc[8];
    const uint32_t base  = tile * (uint32_t)kSortTile;
    const uint32_t valid = min((uint32_t)kSortTile, n - base);

    // ---- load keys, warp-striped: item j of lane l in warp w is element w*512 + j*32 + l of the tile
    uint32_t       key[kSortIpt];
    const uint32_t woff = warp * (32 * kSortIpt) + lane;
    if (valid == (uint32_t)kSortTile)
    {
#pragma unroll
        for (int j = 0; j < kSortIpt; ++j) key[j] = keys_in[base + woff + j * 32];
    }
    else
    {
#pragma unroll
        for (int j = 0; j < kSortIpt; ++j)
        {
            const uint32_t t = woff + j * 32;
            key[j]           = (t < valid) ? keys_in[base + t] : 0xFFFFFFFFu;
        }
    }

    // ---- rank inside the warp (stable: by item, then by lane)
    uint32_t  rank[kSortIpt];
    uint32_t* my_hist = s_warp_hist + warp * kRadix;
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j)
    {
        const uint32_t d      = (key[j] >> shift) & (kRadix - 1);
        const uint32_t peers  = match_digit(d);
        const int      leader = __ffs(peers) - 1;
        uint32_t       prev   = 0;
        if (lane == leader)
        {
            prev       = my_hist[d];
            my_hist[d] = prev + __popc(peers);
        }
        prev    = __shfl_sync(0xffffffffu, prev, leader);
        rank[j] = prev + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // ---- per-digit (one thread each): exclusive scan over warps, tile totals, decoupled look-back
    uint32_t padded_count = 0;
    {
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w)
        {
            const uint32_t c              = s_warp_hist[w * kRadix + tid];
            s_warp_hist[w * kRadix + tid] = acc;
            acc += c;
        }
        padded_count = acc;
    }
    const uint32_t tile_count = (tid == kRadix - 1) ? padded_count - ((uint32_t)kSortTile - valid) : padded_count;  // padding keys are 0xFFFFFFFF
    uint32_t*      my_status  = status + (size_t)tile * kRadix + tid;
    st_status(my_status, tile_count | (tile == 0 ? kFlagPrefix : kFlagAgg));   // publish before anything else
    const uint32_t digit_base = block_exclusive_scan_256(padded_count, s_misc, tid);
    const uint32_t hist_excl  = block_exclusive_scan_256(pass_hist[tid], s_misc, tid);
    uint32_t excl = 0;
    if (tile != 0)
    {
        int t = (int)tile - 1;
        while (true)
        {
            uint32_t s[kLookAhead];
#pragma unroll
            for (int k = 0; k < kLookAhead; ++k) s[k] = (t - k >= 0) ? ld_status(status + (size_t)(t - k) * kRadix + tid) : kFlagPrefix;
            int  k    = 0;
            bool done = false;
#pragma unroll
            for (; k < kLookAhead; ++k)
            {
                if ((s[k] >> 30) == 0) break;
                excl += s[k] & kValueMask;
                if (s[k] & kFlagPrefix) { done = true; break; }
            }
            if (done) break;
            t -= k;
        }
        st_status(my_status, ((excl + tile_count) & kValueMask) | kFlagPrefix);
    }
    s_digit_base[tid]  = digit_base;
    s_global_base[tid] = hist_excl + excl - digit_base;
    __syncthreads();

    // ---- scatter keys and values into shared memory in digit order, then stream them out coalesced
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j)
    {
        const uint32_t d   = (key[j] >> shift) & (kRadix - 1);
        const uint32_t pos = s_digit_base[d] + my_hist[d] + rank[j];
        const uint32_t t   = woff + j * 32;
        s_keys[pos]        = key[j];
        uint32_t v         = 0;
        if (kIdentityValues) v = base + t;
        else if (t < valid) v = vals_in[base + t];
        s_vals[pos] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSortIpt; ++j)
    {
        const uint32_t p = tid + j * kSortThreads;
        if (p < valid)
        {
            const uint32_t k = s_keys[p];
            const uint32_t g = s_global_base[(k >> shift) & (kRadix - 1)] + p;
            keys_out[g]      = k;
            vals_out[g]      = s_vals[p];
        }
    }
}
}  // namespace

SortLayout sort_layout(uint32_t n)
{
    SortLayout L;
    L.n     = n;
    L.tiles = (n + kSortTile - 1) / kSortTile;
    if (L.tiles == 0) L.tiles = 1;
    size_t off    = 0;
    L.hist_off    = off; off += align_up(sizeof(uint32_t) * kPasses * kRadix, 256);
    L.counter_off = off; off += 256;
    L.status_off  = off; off += align_up(sizeof(uint32_t) * (size_t)kPasses * L.tiles * kRadix, 256);
    L.tmp_keys_off = off; off += align_up(sizeof(uint32_t) * (size_t)n, 256);
    L.tmp_vals_off = off; off += align_up(sizeof(uint32_t) * (size_t)n, 256);
    L.total        = off;
    return L;
}

uint32_t* sort_hist_ptr(const SortLayout& L, void* scratch) { return reinterpret_cast<uint32_t*>((char*)scratch + L.hist_off); }

void sort_reset(const DeviceInfo&, cudaStream_t s, const SortLayout& L, void* scratch)
{
    // hist + tickets + status are contiguous: one memset node.
    RR_CUDA_CHECK(cudaMemsetAsync((char*)scratch + L.hist_off, 0, L.tmp_keys_off - L.hist_off, s));
}

void sort_histogram(const DeviceInfo& dev, cudaStream_t s, const SortLayout& L, void* scratch, const uint32_t* keys)
{
    if (L.n == 0) return;
    int grid = (int)std::min<size_t>(((size_t)L.n + 255) / 256, (size_t)dev.sm_count * 8);
    k_sort_histogram<<<grid, 256, 0, s>>>(keys, L.n, sort_hist_ptr(L, scratch));
    ++*dev.launches;
    RR_CUDA_CHECK(cudaGetLastError());
}

void sort_pairs(const DeviceInfo& dev, cudaStream_t s, const SortLayout& L, void* scratch, uint32_t* keys_in, const uint32_t* vals_in,
                uint32_t* keys_out, uint32_t* vals_out, int passes)
{
    passes = std::max(1, std::min(passes, kPasses));
    if (L.n == 0) return;
    static std::once_flag attr_once[kMaxDevices];  // once per device, thread-safe (contexts may live on different threads)
    std::call_once(attr_once[dev.device % kMaxDevices], [] {
        RR_CUDA_CHECK(cudaFuncSetAttribute(k_onesweep_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemBytes));
        RR_CUDA_CHECK(cudaFuncSetAttribute(k_onesweep_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemBytes));
    });
    char*     sc       = (char*)scratch;
    uint32_t* hist     = reinterpret_cast<uint32_t*>(sc + L.hist_off);
    uint32_t* tickets  = reinterpret_cast<uint32_t*>(sc + L.counter_off);
    uint32_t* status   = reinterpret_cast<uint32_t*>(sc + L.status_off);
    uint32_t* tmp_keys = reinterpret_cast<uint32_t*>(sc + L.tmp_keys_off);
    uint32_t* tmp_vals = reinterpret_cast<uint32_t*>(sc + L.tmp_vals_off);
    const uint32_t* kin = keys_in;
    const uint32_t* vin = vals_in;
    for (int p = 0; p < passes; ++p)
    {   // ping-pong so that the last pass lands in keys_out / vals_out
        uint32_t* kout = ((passes - 1 - p) & 1) ? tmp_keys : keys_out;
        uint32_t* vout = ((passes - 1 - p) & 1) ? tmp_vals : vals_out;
        uint32_t* st   = status + (size_t)p * L.tiles * kRadix;
        if (p == 0 && vin == nullptr)
            k_onesweep_pass<true><<<L.tiles, kSortThreads, kSortSmemBytes, s>>>(kin, nullptr, kout, vout, L.n, p * kRadixBits,
                                                                             hist + p * kRadix, tickets + p, st);
        else
            k_onesweep_pass<false><<<L.tiles, kSortThreads, kSortSmemBytes, s>>>(kin, vin, kout, vout, L.n, p * kRadixBits,
                                                                              hist + p * kRadix, tickets + p, st);
        ++*dev.launches;
        kin = kout;
        vin = vout;
    }
    RR_CUDA_CHECK(cudaGetLastError());
}

}  // namespace rr
