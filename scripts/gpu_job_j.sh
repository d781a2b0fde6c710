#!/bin/bash
# round-2 job j: validation of the fused candidate kernel / peeled k_iter1, final numbers, dense-pass captures for roofline.traffic
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r02_pytest_j.txt 2>&1
tail -n 6 gpurun_out/r02_pytest_j.txt
rm -f gpurun_out/r02_stage_times_j.txt
for cfg in "EF_DUMMY=1" "EF_FUSED_MODEL=0" "EF_GN_CLUSTER=0" "EF_IT1_PREFETCH=0" "EF_NO_PDL=1"; do
  echo "== stage times: $cfg" | tee -a gpurun_out/r02_stage_times_j.txt
  env $cfg timeout 300 python scripts/stage_times.py 60 2>&1 | tail -13 | tee -a gpurun_out/r02_stage_times_j.txt
done
timeout 900 python bench.py > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err
EF_IT1_PREFETCH=0 timeout 600 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/r02_bench_j_noprefetch.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_j.json', 'gpurun_out/r02_bench_j_noprefetch.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    r, r2 = d['roofline'], d.get('roofline_1280x960', {})
    print(f, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['ms_per_step'], 4), 'launches', d.get('launches_per_frame'))
    print('  nola', d.get('no_lookahead', {}).get('value'), d.get('no_lookahead', {}).get('e2e'))
    print('  roof640 cold/warm/full', round(r['duration_us'], 2), round(r.get('duration_warm_us', 0), 2), round(r.get('full_iteration', {}).get('duration_us', 0), 2), 'frac', r.get('frac'))
    print('  roof1280 cold/warm/full', round(r2.get('duration_us', 0), 2), round(r2.get('duration_warm_us', 0), 2), round(r2.get('full_iteration', {}).get('duration_us', 0), 2), 'frac', r2.get('frac'))
    print('  1280', d.get('value_1280x960', {}).get('value'), d.get('value_1280x960', {}).get('e2e'))
    for k, v in d.get('large_map', {}).items():
        print('  ', k, 'value', round(v.get('value', 0), 1), 'e2e', round(v.get('e2e', 0), 1), 'ms', round(v.get('ms_per_step', 0), 4))
        for kk, vv in v.get('map_stage_rooflines', {}).items():
            print('      ', kk, 'us', round(vv['duration_us'], 1), 'GB/s', round(vv['achieved']), 'frac', round(vv.get('frac', 0), 3))
    print('  cpu_baseline', d.get('cpu_baseline', {}).get('value'), 'tracking_only', d.get('tracking_only', {}).get('ours_ms'), d.get('tracking_only', {}).get('reference_ms'))
PY
timeout 600 python bench.py --impl reference --steps 40 --warmup 5 > gpurun_out/r02_bench_reference_j.json 2> gpurun_out/r02_bench_reference_j.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_reference_j.json') if l.startswith('{')][-1]); print('reference arm', d['value'], d['cpu_baseline'])"
NCU="ncu --clock-control none"
for sc in 1 2; do
  timeout 400 $NCU --set full --import-source on -k regex:k_iter1 -s 31 -c 1 -f -o gpurun_out/r02_k_iter1_dense_$sc python scripts/prof_icp.py $sc > /dev/null 2>&1
  ncu -i gpurun_out/r02_k_iter1_dense_$sc.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py k_iter1 > gpurun_out/r02_k_iter1_dense_$sc.summary.txt 2>&1; head -12 gpurun_out/r02_k_iter1_dense_$sc.summary.txt
done
timeout 300 $NCU --set full --import-source on -k regex:k_sobel_cand_compact -s 3 -c 1 -f -o gpurun_out/r02_k_sobel_cand_compact_640 python scripts/prof_frames.py 8 > /dev/null 2>&1
ncu -i gpurun_out/r02_k_sobel_cand_compact_640.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py k_sobel_cand_compact > gpurun_out/r02_k_sobel_cand_compact_640.summary.txt 2>&1; head -8 gpurun_out/r02_k_sobel_cand_compact_640.summary.txt
