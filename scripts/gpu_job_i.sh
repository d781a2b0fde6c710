#!/bin/bash
# round-2 job i: tests, the reference arm, ncu re-captures of the kernels changed since job g, launch list, phase profile
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r02_pytest_i.txt 2>&1
tail -n 6 gpurun_out/r02_pytest_i.txt
timeout 600 python bench.py --impl reference --steps 40 --warmup 5 > gpurun_out/r02_bench_reference_i.json 2> gpurun_out/r02_bench_reference_i.err
tail -c 1500 gpurun_out/r02_bench_reference_i.json
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --cache-control none -s 400 -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --quick --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
python scripts/launch_frame_share.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_value_frame.txt 2>&1; head -36 gpurun_out/r02_launches_value_frame.txt
for ks in k_iter2:25 k_iter1:25 k_clean_flags:3 k_model_level0:3 k_model_level_down:5; do
  k=${ks%%:*}; skip=${ks##*:}
  timeout 400 $NCU --set full --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02_${k}_640 python scripts/prof_frames.py 8 > /dev/null 2>&1
  ncu -i gpurun_out/r02_${k}_640.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py $k > gpurun_out/r02_${k}_640.summary.txt 2>&1; cat gpurun_out/r02_${k}_640.summary.txt
done
for k in k_clean_flags; do
  timeout 600 $NCU --set full --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_${k}_5M python scripts/prof_largemap.py 5000000 > /dev/null 2>&1
  ncu -i gpurun_out/r02_${k}_5M.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py $k > gpurun_out/r02_${k}_5M.summary.txt 2>&1; cat gpurun_out/r02_${k}_5M.summary.txt
done
# dominant kernel, isolated dense pass at level 0 (what bench.py's roofline times), for roofline.traffic
timeout 400 $NCU --set full --import-source on -k regex:k_iter1 -s 39 -c 1 -f -o gpurun_out/r02_k_iter1_dense_640 python scripts/prof_icp.py > /dev/null 2>&1
ncu -i gpurun_out/r02_k_iter1_dense_640.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py k_iter1 > gpurun_out/r02_k_iter1_dense_640.summary.txt 2>&1; cat gpurun_out/r02_k_iter1_dense_640.summary.txt
EF_LIB=build/libefusion_prof.so EF_GN_CLUSTER=0 timeout 200 python scripts/phase_profile.py > gpurun_out/r02_phase_profile.txt 2>&1; cat gpurun_out/r02_phase_profile.txt
