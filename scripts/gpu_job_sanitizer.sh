#!/bin/bash
# compute-sanitizer over every kernel family (scripts/sanitize_run.py), programmatic dependent launch on and off; then the GPU test
# tier twice more (flakiness check of the sensitivity-calibrated tests). Summaries -> gpurun_out/r02_sanitizer_final_*.txt
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck initcheck synccheck racecheck; do
  for pdl in 0 1; do
    out=gpurun_out/r02_sanitizer_final_${tool}_nopdl${pdl}.txt
    EF_NO_PDL=$pdl timeout 500 $CS --tool $tool --print-limit 50 python scripts/sanitize_run.py > $out 2>&1
    echo "$tool EF_NO_PDL=$pdl: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok' $out | tr '\n' ' ')"
  done
done
for i in 1 2; do
  timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r02_pytest_repeat_$i.txt 2>&1
  tail -n 2 gpurun_out/r02_pytest_repeat_$i.txt
done
