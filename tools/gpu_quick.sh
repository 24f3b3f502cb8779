#!/bin/bash
# Short GPU visit while iterating on the build kernels: build/refit parity tests, large-mesh timings, launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_build.py tests/test_gpu_scene.py -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -12 gpurun_out/pytest_quick.log
timeout 600 python tools/bench_build.py --sizes 512x512,4000x1000,5000x5000 --reps 5 > gpurun_out/bench_build.log 2>&1; cat gpurun_out/bench_build.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_build50m.csv python tools/bench_build.py --sizes 5000x5000 --reps 1 > gpurun_out/build50m_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_build50m.csv | grep "rr::"
