// Thin driver around the REFERENCE's own CPU code (bvh_analyzer/*.h), compiled from the sources
// where they lie under /root/reference by oracle/Makefile into oracle/_ref/ (git-ignored).
// TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing from the reference is copied into this repository.
//
// It loads a VkBvhNode dump (bvh_analyzer/transform.h:31-41) + an RRRay dump, re-indexes with the
// reference's Transform2, runs the reference's IsValid / CalculateSAH (bvh.h:130-223) and times ONLY
// the `#pragma omp parallel for` loop over BvhIntersect<2> that CheckQuality runs (bvh.h:87-93,226-319)
// -- the stock CheckQuality also does IsValid, SAH, a per-ray omp critical and two JPEG writes.
//
// usage: bvh_analyzer_trace bvh.bin n_internal n_triangles rays.bin n_rays repeats [hits_out.bin [brute_out.bin]]
// prints one JSON line.
//   hits_out.bin : bvh::Hit[n_rays] exactly as BvhIntersect<2> returned them (bvh.h:226-319)
//   brute_out.bin: per ray {float t; uint32 prim; uint32 count; uint32 pad} from the reference's own
//                  Triangle::Intersect (triangle.h:34-70) applied to EVERY triangle of the dump: the (t, prim)-minimum and
//                  the number of triangles the reference's test accepts -- value-level fixtures produced by reference code,
//                  which tests/ compare with the oracle's and the GPU's closest hits.
// RR_REF_STOCK_SCHEDULE=1 runs the trace loop with the stock `#pragma omp parallel for` (bvh.h:87) instead of
// schedule(dynamic, 1024).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>
#include <omp.h>

#define STB_IMAGE_WRITE_IMPLEMENTATION
#define private public  // reach Bvh<2>::IsValid / CalculateSAH, which the reference keeps private
#include "bvh.h"
#include "transform.h"
#undef private

int main(int argc, char** argv)
{
    if (argc < 7) { std::fprintf(stderr, "usage: %s bvh.bin n_internal n_tris rays.bin n_rays repeats [hits.bin]\n", argv[0]); return 2; }
    const size_t n_internal = std::strtoull(argv[2], nullptr, 10), n_tris = std::strtoull(argv[3], nullptr, 10);
    const size_t n_rays = std::strtoull(argv[5], nullptr, 10);
    const int repeats = std::atoi(argv[6]);
    std::vector<bvh::VkBvhNode> vk(n_internal + n_tris);
    std::vector<bvh::Ray> rays(n_rays);
    {
        std::ifstream f(argv[1], std::ifstream::binary);
        if (!f.read((char*)vk.data(), vk.size() * sizeof(bvh::VkBvhNode))) { std::fprintf(stderr, "short bvh file\n"); return 1; }
        std::ifstream g(argv[4], std::ifstream::binary);
        if (!g.read((char*)rays.data(), rays.size() * sizeof(bvh::Ray))) { std::fprintf(stderr, "short ray file\n"); return 1; }
    }
    bvh::Bvh<2u> bvh2(n_internal, n_tris);
    bvh2.TransformBvh(bvh::Transform2<bvh::VkBvhNode>, vk.data(), nullptr);
    const bool valid = bvh2.IsValid();
    const float sah = valid ? bvh2.CalculateSAH() : 0.f;
    std::vector<bvh::Hit> hits(n_rays);
    const char* se = std::getenv("RR_REF_STOCK_SCHEDULE");
    const bool  stock = se && std::atoi(se) != 0;
    double best = 1e30, total = 0;
    double node_tests = 0, tri_tests = 0;
    for (int r = 0; r < repeats && valid; ++r)
    {
        double nt = 0, tt = 0;
        auto t0 = std::chrono::steady_clock::now();
        if (stock)
        {
#pragma omp parallel for reduction(+ : nt, tt)
            for (long long i = 0; i < (long long)n_rays; ++i)
            {
                bvh::BvhIntersect<2u> isect(bvh2, bvh::QueryType::kClosestHit);
                hits[i] = isect(rays[i]);
                nt += isect.stats().num_internal_node_tests;
                tt += isect.stats().num_triangle_tests;
            }
        }
        else
        {
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : nt, tt)
            for (long long i = 0; i < (long long)n_rays; ++i)
            {
                bvh::BvhIntersect<2u> isect(bvh2, bvh::QueryType::kClosestHit);
                hits[i] = isect(rays[i]);
                nt += isect.stats().num_internal_node_tests;
                tt += isect.stats().num_triangle_tests;
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        double s = std::chrono::duration<double>(t1 - t0).count();
        best = s < best ? s : best;
        total += s;
        node_tests = nt; tri_tests = tt;
    }
    size_t hit_count = 0;
    for (auto& h : hits) hit_count += (h.inst_id != bvh::kInvalidID);
    if (argc > 7) { std::ofstream o(argv[7], std::ofstream::binary); o.write((char*)hits.data(), hits.size() * sizeof(bvh::Hit)); }
    if (argc > 8 && valid)
    {
        struct Brute { float t; uint32_t prim; uint32_t count; uint32_t pad; };
        std::vector<Brute> out(n_rays);
        auto const& tris = bvh2.Primitives();
#pragma omp parallel for schedule(dynamic, 16)
        for (long long i = 0; i < (long long)n_rays; ++i)
        {
            Brute b{0.f, bvh::kInvalidID, 0u, 0u};
            for (auto const& tri : tris)
            {
                bvh::float2 uv;
                float       t;
                if (tri.Intersect(rays[i], uv, t))
                {
                    ++b.count;
                    if (b.prim == bvh::kInvalidID || t < b.t || (t == b.t && tri.prim_id < b.prim)) { b.t = t; b.prim = tri.prim_id; }
                }
            }
            out[i] = b;
        }
        std::ofstream o(argv[8], std::ofstream::binary);
        o.write((char*)out.data(), out.size() * sizeof(Brute));
    }
    std::printf("{\"is_valid\": %s, \"sah\": %.6f, \"rays\": %zu, \"repeats\": %d, \"best_s\": %.6f, \"mean_s\": %.6f, "
                "\"mrays_per_s\": %.4f, \"threads\": %d, \"hit_count\": %zu, \"avg_node_tests\": %.3f, \"avg_tri_tests\": %.3f}\n",
                valid ? "true" : "false", sah, n_rays, repeats, best, repeats ? total / repeats : 0.0,
                valid && repeats ? n_rays / (total / repeats) / 1e6 : 0.0, omp_get_max_threads(), hit_count,
                n_rays ? node_tests / n_rays : 0.0, n_rays ? tri_tests / n_rays : 0.0, stock ? "true" : "false");
    return valid ? 0 : 3;
}
