"""Workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every kernel family of the library at a small
size, no torch — first-frame map, 5 tracked + fused frames with look-ahead, the local loop closure front half, clean with a
deformation graph, the stage API of the reductions. Run under `compute-sanitizer --tool <t> python scripts/sanitize_run.py`,
with and without EF_NO_PDL=1."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elasticfusion_b200 import capi, synth

K = synth.Intrinsics(320, 240, 264.0, 264.0, 160.0, 120.0)
frames = list(synth.sequence(8, K, seed=3, noise=True))
ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=150000, time_delta=3, close_loops=1, count_thresh=1000))
ctx.prefetch_frame(frames[0][0], frames[0][1])
for i in range(6):
    ctx.process_frame_device(None, None, i)
    ctx.prefetch_frame(frames[i + 1][0], frames[i + 1][1])
    ctx.finish_frame()
ctx.process_frame(None, None, 6)
info, *_ = ctx.local_loop_result()
m = ctx.map_download()
nodes = np.zeros((8, 16), np.float32)
for k in range(8):
    nodes[k, 0:3] = m[k * (len(m) // 8), 0:3]
    nodes[k, 3:12] = np.eye(3).reshape(-1)
    nodes[k, 12:15] = 0.001 * k
    nodes[k, 15] = 1 + k
ctx.process_frame_begin(frames[7][0], frames[7][1], 7)
ctx.process_frame_end(nodes=nodes)
T = ctx.get_pose()
R, t = T[:3, :3].astype(np.float32), T[:3, 3].astype(np.float32)
for lv in (0, 1, 2):
    ctx.icp_step(lv, R, t, np.linalg.inv(R).astype(np.float32), t)
    ctx.rgb_residual(lv, np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    ctx.rgb_step(lv, 10.0)
ctx.so3_step(np.eye(3, dtype=np.float32), np.eye(3, dtype=np.float32), np.eye(3, dtype=np.float32))
ctx.odom_track(T, rgb_only=True, so3=False)
ctx.map_raycast(None, 20.0, 10.0, ctx.get_tick(), ctx.get_tick(), 3, 1)
print("sanitize_run ok: surfels", ctx.map_count(), "loop ran", info["ran"], "launches", ctx.launch_count())
ctx.close()
