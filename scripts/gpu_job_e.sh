#!/bin/bash
# round-2 job e: split clean + cluster Gauss-Newton: smoke, tests, stage times and A/B of the cluster switches
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/r02_pytest_e.txt 2>&1
tail -n 15 gpurun_out/r02_pytest_e.txt
for cfg in "EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=2" "EF_GN_CLUSTER=0" "EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=3" "EF_GN_CLUSTER=8 EF_GN_CLUSTER_LEVELS=2" "EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=1"; do
  echo "== stage times: $cfg"
  env $cfg timeout 300 python scripts/stage_times.py 60 2>&1 | tail -13 | tee -a gpurun_out/r02_stage_times_e.txt
done
rm -f gpurun_out/r02_ab_e.txt
run() {
  echo "== $*" >> gpurun_out/r02_ab_e.txt
  env "$@" timeout 300 python bench.py --quick --no-cpu-baseline --steps 120 --warmup 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        print('value %.1f e2e %.1f ms/frame %.4f nola %.1f launches/frame %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['no_lookahead']['value'], d['launches_per_frame']))
" >> gpurun_out/r02_ab_e.txt
}
run EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=2
run EF_GN_CLUSTER=0
run EF_GN_CLUSTER=16 EF_GN_CLUSTER_LEVELS=3
run EF_GN_CLUSTER=8 EF_GN_CLUSTER_LEVELS=2
cat gpurun_out/r02_ab_e.txt
timeout 900 python bench.py > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_e.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'nola', d.get('no_lookahead'))
for k, v in d.get('large_map', {}).items():
    print(k, 'value', v.get('value'), 'ms', v.get('ms_per_step'))
    for kk, vv in v.get('map_stage_rooflines', {}).items():
        print('   ', kk, 'us', round(vv['duration_us'], 1), 'GB/s', round(vv['achieved']), 'frac', round(vv.get('frac', 0), 3))
PY
echo "== initcheck after zeroing the reduction scratch"
timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool initcheck --print-limit 200 python scripts/sanitize_run.py > gpurun_out/r02_sanitizer_initcheck_e.txt 2>&1
grep -E "ERROR SUMMARY|sanitize_run ok" gpurun_out/r02_sanitizer_initcheck_e.txt
grep "    at " gpurun_out/r02_sanitizer_initcheck_e.txt | sort | uniq -c | sort -rn | head -20
for tool in memcheck racecheck synccheck; do
  timeout 400 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 50 python scripts/sanitize_run.py > gpurun_out/r02_sanitizer_${tool}_e.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok' gpurun_out/r02_sanitizer_${tool}_e.txt | tr '\n' ' ')"
done
