#!/bin/bash
# round-2 job d: full GPU test tier, then compute-sanitizer (all four tools, PDL on and off)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/r02_pytest_d.txt 2>&1
tail -n 25 gpurun_out/r02_pytest_d.txt
python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --deselect tests/test_gpu_configs.py::test_finite_time_delta_reference_default > gpurun_out/r02_pytest_d_all.txt 2>&1
tail -n 12 gpurun_out/r02_pytest_d_all.txt
sed -i 's/timeout 900/timeout 300/' scripts/gpu_job_sanitizer.sh
bash scripts/gpu_job_sanitizer.sh
