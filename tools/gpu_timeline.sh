#!/bin/bash
mkdir -p gpurun_out
python tools/emit_timeline.py 5000x5000 > gpurun_out/emit_timeline.log 2>&1
cat gpurun_out/emit_timeline.log
bash tools/gpu_quick.sh 2>&1 | grep -E "passed|failed|exit|5000x5000|k_emit" | cut -c1-220
