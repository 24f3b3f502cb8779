python -m pytest tests/test_gpu_trace.py -q -k "grid or primary_640 or cornell_1024 or indirect" 2>&1 | tail -2
for i in 1 2; do python bench.py --steps 30 --warmup 5 --no-c3 --no-c4 --no-c5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['trace_variants']['closest_full_hit_other_bvh_mrays'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_detect_grid|k_trace" -c 12 --csv --log-file gpurun_out/launches_grid2.csv python tools/profile_trace.py --reps 2 > /dev/null 2>&1
