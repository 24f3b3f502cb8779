#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_emit_leaves|k_emit_window' -c 2 -o gpurun_out/prof_emit -f python tools/bench_build.py --sizes 4000x1000 --reps 1 > gpurun_out/ncu_emit.log 2>&1
tail -3 gpurun_out/ncu_emit.log
