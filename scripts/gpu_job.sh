#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -k "ferns or icl_nuim or finite_time or depth_cutoff or capacity_overflow or per_iteration or local_loop or gl_golden or reference_shaders" 2>&1 > gpurun_out/r02_pytest_c.txt; grep -E "^E  |^FAILED|passed|failed|^tests/.*Error|^___" gpurun_out/r02_pytest_c.txt | cut -c1-400 | head -120
