"""Tracks a few frames at the given resolution, then issues ICP-reduction launches at level 0 (for ncu --set full captures)."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elasticfusion_b200 import synth, capi
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
K = synth.K_DEFAULT.scaled(scale)
frames = list(synth.sequence(3, K, seed=42, noise=True))
ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=2000000, time_delta=2147483647 // 2))
for i, (rgb, d, _) in enumerate(frames):
    ctx.process_frame(rgb, d, i)
T = ctx.get_pose()
R = T[:3, :3].astype(np.float32); t = T[:3, 3].astype(np.float32)
ctx.icp_step_async(0, R, t, np.linalg.inv(R).astype(np.float32), t)
ctx.sync()
for _ in range(4):
    ctx.icp_step_async(0)
ctx.sync()
print("ok", K.width, K.height, ctx.map_count())
