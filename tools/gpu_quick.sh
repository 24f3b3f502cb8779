#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size.py -x -q --durations=5 > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -25 gpurun_out/pytest_quick.log
