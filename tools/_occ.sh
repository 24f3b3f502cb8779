timeout 600 python -m pytest tests/test_gpu_build.py tests/test_gpu_scene.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -3
python tools/bench_build.py --sizes 5000x5000 --reps 5 2>&1 | cut -c1-120
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/l.csv python tools/bench_build.py --sizes 5000x5000 --reps 1 > /dev/null 2>&1; python tools/summarize_launches.py gpurun_out/l.csv | grep "rr::"
