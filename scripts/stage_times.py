"""Per-stage device time of ef_process_frame_device (EF_STAGE_TIMING=1): median over the steady-state frames."""
import os, sys, ctypes as C, numpy as np
os.environ["EF_STAGE_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elasticfusion_b200 import synth, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
K = synth.K_DEFAULT
frames = list(synth.sequence(n, K, seed=42, noise=True))
BIG = 2147483647 // 2
ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=5000000, time_delta=BIG))
lib = capi.lib()
lib.ef_debug_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
names = ["", "upload+bilateral", "pyramids+so3", "model pyramids", "sobel+cand", "gauss-newton", "finish", "index#1", "fuse", "index#2", "clean", "predict"]
rows = []
for i, (rgb, d, _) in enumerate(frames):
    ctx.process_frame(rgb, d, i)
    out = (C.c_float * 16)()
    k = lib.ef_debug_stage_ms(ctx.h_ctx, out)
    if i >= 10:
        rows.append([out[j] for j in range(1, k)])
m = np.median(np.array(rows), axis=0) * 1e3
for nm, v in zip(names[1:], m):
    print(f"{nm:14s} {v:8.1f} us")
print(f"{'total':14s} {m.sum():8.1f} us")
