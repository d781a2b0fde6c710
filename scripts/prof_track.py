"""Sets up a tracked 640x480 state and runs one full getIncrementalTransformation (for ncu captures)."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elasticfusion_b200 import synth, capi
K = synth.K_DEFAULT
frames = list(synth.sequence(4, K, seed=42, noise=True))
BIG = 2147483647 // 2
ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=1000000, time_delta=BIG))
for i, (rgb, d, _) in enumerate(frames):
    ctx.process_frame(rgb, d, i)
T = ctx.get_pose()
R = T[:3, :3].astype(np.float32); t = T[:3, 3].astype(np.float32)
for _ in range(3):
    ctx.icp_step_async(0, R, t, np.linalg.inv(R).astype(np.float32), t)
ctx.sync()
print("ok", ctx.map_count())
