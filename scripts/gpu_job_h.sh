#!/bin/bash
# round-2 job h: after the GN tail / model-side / clean-flags changes: tests, stage times (A/B of the fused model side), full bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r02_pytest_h.txt 2>&1
tail -n 8 gpurun_out/r02_pytest_h.txt
rm -f gpurun_out/r02_stage_times_h.txt
for cfg in "EF_DUMMY=1" "EF_FUSED_MODEL=0" "EF_GN_CLUSTER=0"; do
  echo "== stage times: $cfg" | tee -a gpurun_out/r02_stage_times_h.txt
  env $cfg timeout 300 python scripts/stage_times.py 60 2>&1 | tail -13 | tee -a gpurun_out/r02_stage_times_h.txt
done
timeout 900 python bench.py > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_h.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], 'launches', d.get('launches_per_frame'))
print('nola', d.get('no_lookahead'))
print('roofline', d['roofline'])
print('1280', d.get('value_1280x960'))
for k, v in d.get('large_map', {}).items():
    print(k, 'value', v.get('value'), 'ms', v.get('ms_per_step'))
    for kk, vv in v.get('map_stage_rooflines', {}).items():
        print('   ', kk, 'us', round(vv['duration_us'], 1), 'GB/s', round(vv['achieved']), 'frac', round(vv.get('frac', 0), 3))
print('cpu_baseline', d.get('cpu_baseline'))
print('tracking_only', d.get('tracking_only'))
PY
