#!/bin/bash
# large-map passes after the memory-level-parallelism changes: map tests, stage times, bench (large maps), ncu at 5 M
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "map or raycast or index or fuse or clean or pipeline or gl_golden or loop or capacity or first_frame" > gpurun_out/r02_pytest_map.txt 2>&1
tail -n 5 gpurun_out/r02_pytest_map.txt
timeout 300 python scripts/stage_times.py 60 2>&1 | tail -13
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_map.json 2> gpurun_out/r02_bench_map.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_map.json') if l.startswith('{')][-1])
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), '1280', d.get('value_1280x960', {}).get('value'))
for k, v in d.get('large_map', {}).items():
    print('  ', k, 'value', round(v.get('value', 0), 1), 'e2e', round(v.get('e2e', 0), 1), 'ms', round(v.get('ms_per_step', 0), 4))
    for kk, vv in v.get('map_stage_rooflines', {}).items():
        print('      ', kk, 'us', round(vv['duration_us'], 1), 'GB/s', round(vv['achieved']), 'frac', round(vv.get('frac', 0), 3))
PY
NCU="ncu --clock-control none"
for k in k_clean_flags k_splat_scatter k_index_scatter; do
  timeout 600 $NCU --set full --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_${k}_5M python scripts/prof_largemap.py 5000000 > /dev/null 2>&1
  ncu -i gpurun_out/r02_${k}_5M.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py $k > gpurun_out/r02_${k}_5M.summary.txt 2>&1; head -9 gpurun_out/r02_${k}_5M.summary.txt
done
