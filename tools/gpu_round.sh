#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), per-config tools, ncu launch lists and full captures.
mkdir -p gpurun_out
rm -f gpurun_out/trace_modes.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
for mode in "primary closest quality" "primary closest fast" "diffuse closest quality" "shadow any quality"; do
  set -- $mode
  echo "== $mode" >> gpurun_out/trace_modes.log
  timeout 300 python tools/profile_trace.py --rays $1 --query $2 --bvh $3 --reps 4 >> gpurun_out/trace_modes.log 2>&1
done
echo "== primary closest quality (generic loop)" >> gpurun_out/trace_modes.log
RR_CUDA_TRACE_GENERIC=1 timeout 300 python tools/profile_trace.py --reps 4 >> gpurun_out/trace_modes.log 2>&1
cat gpurun_out/trace_modes.log
timeout 600 python tools/bench_scene.py --check 200000 > gpurun_out/bench_scene.log 2>&1; tail -1 gpurun_out/bench_scene.log
timeout 600 python tools/bench_build.py --sizes 1000x500,4000x1000,5000x5000 --reps 5 > gpurun_out/bench_build.log 2>&1; cat gpurun_out/bench_build.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_build50m.csv python tools/bench_build.py --sizes 5000x5000 --reps 1 > gpurun_out/build50m_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_trace$ -s 1 -c 1 -o gpurun_out/prof_trace -f python tools/profile_trace.py --reps 2 > gpurun_out/ncu_trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_onesweep|k_emit|k_morton|k_scene_aabb|k_refit' -c 24 -o gpurun_out/prof_build -f python tools/bench_build.py --sizes 4000x1000 --reps 1 > gpurun_out/ncu_build.log 2>&1
