#!/bin/bash
# ncu evidence for round 2: launch list of a frame (shares), --set full captures of the kernels VERDICT r01 asked for, the same
# for the full-map passes on a 5 M-surfel map, and the latency floor microbenchmark. Everything lands in gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r02_pytest_g.txt 2>&1
tail -n 12 gpurun_out/r02_pytest_g.txt
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --cache-control none -s 400 -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --quick --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
python scripts/launch_frame_share.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_value_frame.txt 2>&1; head -40 gpurun_out/r02_launches_value_frame.txt
for ks in k_iter2:25 k_iter1:25 k_gn_cluster:3 k_preprocess_depth:3 k_fuse_associate:3 k_index_resolve:6 k_clean_flags:3 k_clean_move:3 k_splat_scatter:3 k_index_scatter:6; do
  k=${ks%%:*}; skip=${ks##*:}
  timeout 400 $NCU --set full --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02_${k}_640 python scripts/prof_frames.py 8 > /dev/null 2>&1
  ncu -i gpurun_out/r02_${k}_640.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py $k > gpurun_out/r02_${k}_640.summary.txt 2>&1; cat gpurun_out/r02_${k}_640.summary.txt
done
for k in k_clean_flags k_clean_move k_splat_scatter k_index_scatter; do
  timeout 600 $NCU --set full --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_${k}_5M python scripts/prof_largemap.py 5000000 > /dev/null 2>&1
  ncu -i gpurun_out/r02_${k}_5M.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py $k > gpurun_out/r02_${k}_5M.summary.txt 2>&1; cat gpurun_out/r02_${k}_5M.summary.txt
done
./build/latency_floor > gpurun_out/r02_latency_floor.txt 2>&1; cat gpurun_out/r02_latency_floor.txt
EF_LIB=build/libefusion_prof.so timeout 200 python scripts/phase_profile.py > gpurun_out/r02_phase_profile.txt 2>&1; cat gpurun_out/r02_phase_profile.txt
