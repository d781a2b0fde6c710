#!/bin/bash
# compute-sanitizer over every kernel family, with and without programmatic dependent launch; summaries -> gpurun_out/r02_sanitizer_*.txt
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck initcheck synccheck racecheck; do
  for pdl in 0 1; do
    out=gpurun_out/r02_sanitizer_${tool}_nopdl${pdl}.txt
    EF_NO_PDL=$pdl timeout 900 $CS --tool $tool --print-limit 20 python scripts/sanitize_run.py > $out 2>&1
    echo "== $tool EF_NO_PDL=$pdl: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok' $out | tr '\n' ' ')"
  done
done
