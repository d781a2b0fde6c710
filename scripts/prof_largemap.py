"""A resident map of N surfels (default 5 M) at 640x480 and a few frames + stage calls on it, for ncu captures of the full-map
passes (k_index_scatter, k_clean_compact, k_splat_scatter)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elasticfusion_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
K = synth.K_DEFAULT
frames = list(synth.sequence(4, K, seed=42, noise=True))
BIG = 2147483647 // 2
ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=n + 600_000, time_delta=BIG))
ctx.process_frame(frames[0][0], frames[0][1], 0)
ctx.map_upload(synth.room_surfels(n, np.linalg.inv(synth.trajectory(1, seed=42)[0]), focal=K.fx))
ctx.predict()
for i in (1, 2, 3):
    ctx.process_frame(frames[i][0], frames[i][1], i)
print("ok", ctx.map_count())
