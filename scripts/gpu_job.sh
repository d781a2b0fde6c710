#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 2>&1 > gpurun_out/r02_pytest_c.txt; grep -E "^E  |^FAILED|passed|failed|^tests/.*Error|^___" gpurun_out/r02_pytest_c.txt | cut -c1-300 | head -150
timeout 200 python scripts/stage_times.py 60 > gpurun_out/r02_stage_times_c.txt 2>&1; cat gpurun_out/r02_stage_times_c.txt
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_c.json').read().strip().split('\n')[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'nola',d.get('no_lookahead'))
    print('roof',d['roofline']['duration_us'],d['roofline']['frac'],'full',d['roofline']['full_iteration']['duration_us'])
    print('1280',d.get('value_1280x960'))
    for n,v in d.get('large_map',{}).items():
        print(n,'value',v['value'],'e2e',v['e2e'],'ms',v['ms_per_step'])
        for s,x in v['map_stage_rooflines'].items(): print('   ',s,'us',round(x['duration_us'],1),'GB/s',round(x['achieved'],0),'frac',round(x['frac'],3))
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/r02_bench_c.err').read()[-2000:])
PY
