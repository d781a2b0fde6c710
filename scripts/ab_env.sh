#!/bin/bash
# A/B of the environment switches on the GPU box: headline frames/s (quick bench) for each setting; appends to gpurun_out/r02_ab.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "== $*" >> gpurun_out/r02_ab.txt
  env "$@" timeout 300 python bench.py --quick --no-cpu-baseline --steps 120 --warmup 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        print('value %.1f e2e %.1f ms/frame %.4f roof_us %.2f full_us %.2f launches/frame %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['duration_us'], d['roofline']['full_iteration']['duration_us'], d['launches_per_frame']))
" >> gpurun_out/r02_ab.txt
}
run EF_DUMMY=1
run EF_IT1_PREFETCH=0
run EF_IT2_MAXBLOCKS=64
run EF_IT2_MAXBLOCKS=32
run EF_IT2_MAXBLOCKS=16
run EF_IT2_MAXBLOCKS=32 EF_IT1_PREFETCH=0
run EF_NO_PDL=1
cat gpurun_out/r02_ab.txt
