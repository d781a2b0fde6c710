#!/bin/bash
# visible-list index pass: full GPU tests, large-map bench A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "visible or pipeline or predict_indices or gl_golden or fuse or sequence" > gpurun_out/r02_pytest_vis.txt 2>&1
tail -n 4 gpurun_out/r02_pytest_vis.txt
for cfg in "EF_DUMMY=1" "EF_VISIBLE_LIST=0"; do
  echo "== $cfg"
  env $cfg timeout 300 python scripts/stage_times.py 60 2>&1 | grep "index\|total"
  env $cfg timeout 900 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/r02_bench_vis_$cfg.json 2>/dev/null
  python - "gpurun_out/r02_bench_vis_$cfg.json" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), '1280', round(d['value_1280x960']['value'], 1))
for k, v in d.get('large_map', {}).items():
    print('  ', k, 'value', round(v['value'], 1), 'ms', round(v['ms_per_step'], 4))
PY
done
