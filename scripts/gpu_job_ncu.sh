#!/bin/bash
# ncu --set full captures of the hot kernels inside a running 640x480 sequence and on a 5 M-surfel map, the latency-floor
# microbenchmark and the phase profile of one Gauss-Newton iteration. Summaries (scripts/ncu_extract.py) -> gpurun_out/.
#   usage: bash scripts/gpu_job_ncu.sh [kernel:skip ...]     (default: the kernels added last in round 2)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --clock-control none"
LIST="${@:-k_so3_cluster:3 k_index_scatter:6 k_index_scatter:7 k_iter2:25 k_gn_cluster:3}"
i=0
for ks in $LIST; do
  k=${ks%%:*}; skip=${ks##*:}; i=$((i+1))
  out=gpurun_out/r02_${k}_640_s${skip}
  timeout 400 $NCU --set full --import-source on -k regex:$k -s $skip -c 1 -f -o $out python scripts/prof_frames.py 8 > /dev/null 2>&1
  ncu -i $out.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py $k > $out.summary.txt 2>&1; head -14 $out.summary.txt
done
for k in k_index_scatter; do
  for skip in 2 3; do
    out=gpurun_out/r02_${k}_5M_s${skip}
    timeout 600 $NCU --set full --import-source on -k regex:$k -s $skip -c 1 -f -o $out python scripts/prof_largemap.py 5000000 > /dev/null 2>&1
    ncu -i $out.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_extract.py $k > $out.summary.txt 2>&1; head -9 $out.summary.txt
  done
done
./build/latency_floor > gpurun_out/r02_latency_floor.txt 2>&1
EF_LIB=build/libefusion_prof.so EF_GN_CLUSTER=0 timeout 200 python scripts/phase_profile.py > gpurun_out/r02_phase_profile.txt 2>&1; cat gpurun_out/r02_phase_profile.txt
