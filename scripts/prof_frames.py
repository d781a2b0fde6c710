"""Runs N frames of the 640x480 sequence through ef_process_frame (for ncu captures of the per-frame kernels)."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elasticfusion_b200 import synth, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
K = synth.K_DEFAULT
frames = list(synth.sequence(n, K, seed=42, noise=True))
BIG = 2147483647 // 2
ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=1000000, time_delta=BIG))
for i, (rgb, d, _) in enumerate(frames):
    ctx.process_frame(rgb, d, i)
print("ok", ctx.map_count())
