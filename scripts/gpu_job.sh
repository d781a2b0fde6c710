#!/bin/bash
# The job the retry loop (scripts/gpu.sh) runs on the GPU box; edited as the round goes on.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/r02_pytest_b.txt; tail -40 gpurun_out/r02_pytest_b.txt
timeout 200 python scripts/stage_times.py 60 > gpurun_out/r02_stage_times_b.txt 2>&1; cat gpurun_out/r02_stage_times_b.txt
