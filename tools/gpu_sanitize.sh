#!/bin/bash
# compute-sanitizer over the small build / scene / trace parity tests (memcheck, then racecheck on shared memory).
mkdir -p gpurun_out
SEL='small_random or sizes_around or clustered or octree or cornell or single_triangle or duplicate'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_build.py tests/test_gpu_scene.py -x -q -k "$SEL or scene or tlas or instance" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
tail -6 gpurun_out/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_build.py -x -q -k "sizes_around or clustered or octree" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitize_racecheck.log
tail -6 gpurun_out/sanitize_racecheck.log
