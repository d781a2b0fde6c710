#!/bin/bash
# The evidence run of a round on one B200: GPU test tier, stage times with the A/B switches, the N=1 bench line of both arms, the
# per-frame launch list. Outputs land in gpurun_out/ (copied into profiles/ by hand once read).
#   scripts/gpu.sh 3000 'bash scripts/gpu_job_final.sh'
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r02_pytest_final.txt 2>&1
tail -n 6 gpurun_out/r02_pytest_final.txt
rm -f gpurun_out/r02_stage_times_final.txt
for cfg in "EF_DUMMY=1" "EF_SO3_CLUSTER=0" "EF_FUSED_MODEL=0" "EF_VISIBLE_LIST=0" "EF_GN_CLUSTER=0" "EF_IT1_PREFETCH=0" "EF_NO_PDL=1"; do
  echo "== stage times: $cfg" | tee -a gpurun_out/r02_stage_times_final.txt
  env $cfg timeout 300 python scripts/stage_times.py 60 2>&1 | tail -13 | tee -a gpurun_out/r02_stage_times_final.txt
done
timeout 600 python bench.py --impl reference --steps 40 --warmup 5 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_final.json') if l.startswith('{')][-1])
r, r2 = d['roofline'], d.get('roofline_1280x960', {})
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['ms_per_step'], 4), 'launches', d.get('launches_per_frame'))
print('  nola', d.get('no_lookahead', {}).get('value'), d.get('no_lookahead', {}).get('e2e'))
print('  roof640 cold/warm/full', round(r['duration_us'], 2), round(r.get('duration_warm_us', 0), 2), round(r.get('full_iteration', {}).get('duration_us', 0), 2), 'frac', r.get('frac'), 'traffic', r.get('traffic'))
print('  roof1280 cold/warm/full', round(r2.get('duration_us', 0), 2), round(r2.get('duration_warm_us', 0), 2), round(r2.get('full_iteration', {}).get('duration_us', 0), 2), 'frac', r2.get('frac'))
print('  1280', d.get('value_1280x960', {}).get('value'), d.get('value_1280x960', {}).get('e2e'))
for k, v in d.get('large_map', {}).items():
    print('  ', k, 'value', round(v.get('value', 0), 1), 'e2e', round(v.get('e2e', 0), 1), 'ms', round(v.get('ms_per_step', 0), 4))
print('  cpu_baseline', d.get('cpu_baseline', {}).get('value'), 'tracking_only', d.get('tracking_only', {}).get('ours_ms'), d.get('tracking_only', {}).get('reference_ms'))
print('  clocks', d.get('clocks'))
r = json.loads([l for l in open('gpurun_out/r02_bench_reference_final.json') if l.startswith('{')][-1])
print('reference arm', r['value'], r['cpu_baseline'])
PY
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --cache-control none -s 400 -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --quick --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
python scripts/launch_frame_share.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_value_frame.txt 2>&1; head -36 gpurun_out/r02_launches_value_frame.txt
