#!/bin/bash
# Round-2 evidence run on one B200: launch list of bench.py (every kernel incl. the trace kernels), full ncu capture of the
# dominant kernel, and compute-sanitizer (memcheck / synccheck / racecheck) over the trace, scene, refit and treelet tests.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-c3 --no-c4 > gpurun_out/r2_bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench.txt 2>&1; tail -30 gpurun_out/r2_launches_bench.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_packet -s 2 -c 1 -o gpurun_out/r2_prof_packet -f \
  python tools/profile_trace.py --reps 3 > gpurun_out/r2_ncu_packet.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_trace$' -s 1 -c 1 -o gpurun_out/r2_prof_diffuse -f \
  python tools/profile_trace.py --rays diffuse --reps 2 > gpurun_out/r2_ncu_diffuse.log 2>&1
SEL_TRACE='single_triangle or indirect or deep_stack or tie_rules or misaligned or ray_grid'
SEL_BUILD='update or refit_hand_over or quality or restructure or uint16 or small_random or sizes_around'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_trace.py tests/test_gpu_scene.py tests/test_gpu_build.py \
  -x -q -k "$SEL_TRACE or scene or instance or external or $SEL_BUILD" > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_sanitize_memcheck.log
tail -5 gpurun_out/r2_sanitize_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 99 python -m pytest tests/test_gpu_trace.py tests/test_gpu_scene.py tests/test_gpu_build.py \
  -x -q -k "single_triangle or indirect or deep_stack or ray_grid or instanced_grid or update_refit or update_after_quality or sizes_around" > gpurun_out/r2_sanitize_synccheck.log 2>&1; echo "synccheck exit $?" >> gpurun_out/r2_sanitize_synccheck.log
tail -4 gpurun_out/r2_sanitize_synccheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_trace.py tests/test_gpu_build.py \
  -x -q -k "single_triangle or indirect or deep_stack or ray_grid or sizes_around or clustered or update_after_quality" > gpurun_out/r2_sanitize_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2_sanitize_racecheck.log
tail -4 gpurun_out/r2_sanitize_racecheck.log
