// Microbenchmark (experiment, not product): how fast can one SM deliver a WARP-UNIFORM 64-byte record to all 32 lanes?
// Decides the data layout of the packet traversal: if the L1 return path charges a broadcast load
// like a per-lane load (128 B/clk/SM counted over lanes x bytes), a 64-B BVH2 node costs 16 clk per warp-visit whatever
// the address pattern.  Variants:
//   0 uniform LDG.128 x4 (one 64-B node, same address in every lane)
//   1 uniform LDS.128 x4 (node staged in shared memory)
//   2 lane l < 16 loads word l of the node (one 4-B LDG), 16 SHFL.IDX broadcast the words
//   3 LDC from a __constant__ array with a warp-uniform dynamic index (16 words)
//   4 per-lane divergent LDG.256 x2 at a 64-B stride (k_trace's pattern today)
//   5 uniform LDG.128 x2 (a 32-B compact node)
//   6 uniform LDG.64 x1 + LDG.128 x1 (24 B)
// Prints cycles per warp-visit per SM (all resident warps together) = elapsed SM cycles / visits issued on that SM.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kNodes = 1024;  // 64 KB: L1 resident
__constant__ float4 c_nodes[kNodes * 4 / 4 * 1];  // 16 KB of constant: 256 nodes x 4 quads

__device__ __forceinline__ float sum4(float4 v) { return (v.x + v.y) + (v.z + v.w); }

template <int V>
__global__ void __launch_bounds__(128) bench(const float4* __restrict__ nodes, int iters, float* out, long long* cycles)
{
    __shared__ float4 s_nodes[256 * 4];
    for (int i = threadIdx.x; i < 256 * 4; i += blockDim.x) s_nodes[i] = nodes[i];
    __syncthreads();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    float acc = 0.f;
    uint32_t idx = warp * 7u;
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i)
    {
        idx = idx * 1664525u + 1013904223u;
        const uint32_t n = (idx >> 8) & (V == 1 || V == 3 ? 255u : (uint32_t)(kNodes - 1));
        if (V == 0)
        {
            const float4 a = __ldg(nodes + n * 4), b = __ldg(nodes + n * 4 + 1), c = __ldg(nodes + n * 4 + 2), d = __ldg(nodes + n * 4 + 3);
            acc += sum4(a) + sum4(b) + sum4(c) + sum4(d);
        }
        else if (V == 1)
        {
            const float4 a = s_nodes[n * 4], b = s_nodes[n * 4 + 1], c = s_nodes[n * 4 + 2], d = s_nodes[n * 4 + 3];
            acc += sum4(a) + sum4(b) + sum4(c) + sum4(d);
        }
        else if (V == 2)
        {
            const float w = __ldg(reinterpret_cast<const float*>(nodes + n * 4) + (lane & 15));
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) s += __shfl_sync(0xffffffffu, w, k);
            acc += s;
        }
        else if (V == 3)
        {
            const float4 a = c_nodes[n * 4], b = c_nodes[n * 4 + 1], c = c_nodes[n * 4 + 2], d = c_nodes[n * 4 + 3];
            acc += sum4(a) + sum4(b) + sum4(c) + sum4(d);
        }
        else if (V == 4)
        {
            const uint32_t m = (n + lane * 17u) & (kNodes - 1);
            const float4 a = __ldg(nodes + m * 4), b = __ldg(nodes + m * 4 + 1), c = __ldg(nodes + m * 4 + 2), d = __ldg(nodes + m * 4 + 3);
            acc += sum4(a) + sum4(b) + sum4(c) + sum4(d);
        }
        else if (V == 5)
        {
            const float4 a = __ldg(nodes + n * 4), b = __ldg(nodes + n * 4 + 1);
            acc += sum4(a) + sum4(b);
        }
        else if (V == 6)
        {
            const float4 a = __ldg(nodes + n * 4);
            const float2 b = __ldg(reinterpret_cast<const float2*>(nodes + n * 4 + 1));
            acc += sum4(a) + b.y + b.x;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, const float4* d_nodes, float* d_out, long long* d_cycles, int ctas_per_sm, int sms)
{
    const int iters = 20000, grid = sms * ctas_per_sm;
    bench<V><<<grid, 128>>>(d_nodes, 100, d_out, d_cycles);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<V><<<grid, 128>>>(d_nodes, iters, d_out, d_cycles);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = (long long*)malloc(sizeof(long long) * grid);
    cudaMemcpy(h, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    const double visits_per_sm = (double)iters * ctas_per_sm * 4;
    std::printf("%-44s ctas/sm %2d: %.2f clk per warp-visit per SM (kernel %.3f ms, %s)\n", name, ctas_per_sm, mean / visits_per_sm, ms,
                cudaGetErrorString(cudaGetLastError()));
    free(h);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    float4* h_nodes = (float4*)malloc(sizeof(float4) * kNodes * 4);
    for (int i = 0; i < kNodes * 4; ++i) h_nodes[i] = make_float4(i * 0.5f, i * 0.25f, i * 0.125f, (float)i);
    float4* d_nodes; float* d_out; long long* d_cycles;
    cudaMalloc(&d_nodes, sizeof(float4) * kNodes * 4);
    cudaMalloc(&d_out, sizeof(float) * sms * 16 * 128);
    cudaMalloc(&d_cycles, sizeof(long long) * sms * 16);
    cudaMemcpy(d_nodes, h_nodes, sizeof(float4) * kNodes * 4, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_nodes, h_nodes, sizeof(c_nodes));
    for (int c : {4, 10})
    {
        run<0>("0 uniform LDG.128 x4 (64 B)", d_nodes, d_out, d_cycles, c, sms);
        run<1>("1 uniform LDS.128 x4 (64 B)", d_nodes, d_out, d_cycles, c, sms);
        run<2>("2 16 lanes LDG.32 + 16 SHFL", d_nodes, d_out, d_cycles, c, sms);
        run<3>("3 LDC uniform dynamic index (64 B)", d_nodes, d_out, d_cycles, c, sms);
        run<4>("4 per-lane LDG.128 x4, 64-B nodes (today)", d_nodes, d_out, d_cycles, c, sms);
        run<5>("5 uniform LDG.128 x2 (32 B)", d_nodes, d_out, d_cycles, c, sms);
        run<6>("6 uniform LDG.128 + LDG.64 (24 B)", d_nodes, d_out, d_cycles, c, sms);
    }
    return 0;
}
