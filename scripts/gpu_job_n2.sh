#!/bin/bash
# two ranks on one box: both arms of bench.py exactly as the driver launches them
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02_bench_reference_n2.json 2> gpurun_out/r02_bench_reference_n2.err
tail -c 600 gpurun_out/r02_bench_reference_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_n2.json') if l.startswith('{')][-1])
print('N=2 value', d['value'], 'e2e', d['e2e']['value'], 'n_gpus', d['n_gpus'], 'ms', d['ms_per_step'], 'scaling', d['scaling'], 'clocks', d.get('clocks'))
PY
tail -3 gpurun_out/r02_bench_n2.err
