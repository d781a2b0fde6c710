#!/bin/bash
# round-2 job f: fused single-launch iteration (k_iter) x cluster settings: tests, stage times, quick bench A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r02_pytest_f.txt 2>&1
tail -n 15 gpurun_out/r02_pytest_f.txt
rm -f gpurun_out/r02_stage_times_f.txt gpurun_out/r02_ab_f.txt
for cfg in "EF_ITER_FUSED=1 EF_GN_CLUSTER=0" "EF_ITER_FUSED=1 EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=1" "EF_ITER_FUSED=1 EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=2" "EF_ITER_FUSED=0 EF_GN_CLUSTER=0" "EF_ITER_FUSED=1 EF_GN_CLUSTER=0 EF_IT1_PREFETCH=0"; do
  echo "== stage times: $cfg" | tee -a gpurun_out/r02_stage_times_f.txt
  env $cfg timeout 300 python scripts/stage_times.py 60 2>&1 | tail -13 | tee -a gpurun_out/r02_stage_times_f.txt
done
run() {
  echo "== $*" >> gpurun_out/r02_ab_f.txt
  env "$@" timeout 300 python bench.py --quick --no-cpu-baseline --steps 120 --warmup 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        print('value %.1f e2e %.1f ms/frame %.4f launches/frame %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['launches_per_frame']))
" >> gpurun_out/r02_ab_f.txt
}
run EF_ITER_FUSED=1 EF_GN_CLUSTER=0
run EF_ITER_FUSED=1 EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=1
run EF_ITER_FUSED=1 EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=2
run EF_ITER_FUSED=0 EF_GN_CLUSTER=0
run EF_ITER_FUSED=0 EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=1
cat gpurun_out/r02_ab_f.txt
