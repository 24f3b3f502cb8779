#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_restructure' -s 3 -c 1 -o gpurun_out/prof_treelet -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_treelet.log 2>&1
tail -2 gpurun_out/ncu_treelet.log
