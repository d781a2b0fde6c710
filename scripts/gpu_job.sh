#!/bin/bash
# The job the retry loop (scripts/gpu.sh) runs on the GPU box; edited as the round goes on.
mkdir -p gpurun_out
{ echo "== ldconfig"; ldconfig -p | grep -Ei "egl|glx|opengl|gles|libGL|glvnd|vulkan|osmesa" ; echo "== nvidia libs"; ls /usr/lib/x86_64-linux-gnu | grep -i -E "nvidia|egl|gl" ; echo "== glvnd egl vendors"; ls -la /usr/share/glvnd/egl_vendor.d /etc/glvnd/egl_vendor.d 2>&1; echo "== find"; find / -xdev \( -name "libEGL*" -o -name "libGLX*" -o -name "libOpenGL*" -o -name "libGLESv2*" -o -name "libnvidia-egl*" -o -name "libnvidia-gl*" -o -name "libGL.so*" -o -name "libOSMesa*" \) 2>/dev/null | head -50; echo "== dri"; ls -la /dev/dri 2>&1; echo "== nvidia-smi"; nvidia-smi; nproc; free -g | head -2; echo "== NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"; ls /dev | grep -i nvidia; echo "== egl_probe"; ./oracle/_ref/egl_probe; echo "rc=$?"; } > gpurun_out/r02_egl_probe.txt 2>&1
tail -12 gpurun_out/r02_egl_probe.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -30 | tee gpurun_out/r02_pytest_a.txt
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 3000 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
timeout 900 bash scripts/ab_env.sh 2>&1 | tail -20
