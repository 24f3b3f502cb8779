timeout 600 python -m pytest tests/test_gpu_build.py tests/test_gpu_scene.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -3
for o in 2 3; do echo "occ $o"; RR_EMIT_OCC=$o python tools/bench_build.py --sizes 5000x5000 --reps 5 2>&1 | cut -c1-120; done
