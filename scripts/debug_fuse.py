import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, '/root/repo/tests')
from elasticfusion_b200 import synth, capi
from oracle import ef_oracle as eo
from util import run_oracle
K = synth.K_DEFAULT
frames = list(synth.sequence(5, K, seed=42, noise=True))
f = run_oracle(frames, K, 4)
rgb, depth, _ = frames[4]
filt = eo.bilateral(depth, 3.0); dm, dmf = eo.metric(depth, 3.0), eo.metric(filt, 3.0)
m = f.map(); T = f.pose; tick = f.tick; BIG = 2147483647//2
ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=500000, time_delta=BIG))
ctx.upload("RGB", rgb); ctx.upload("DEPTH_METRIC", dm); ctx.upload("DEPTH_METRIC_FILTERED", dmf)
idx = eo.predict_indices(m, T, tick, 20.0, BIG, K)
fused, new = eo.fuse(m, T, tick, rgb, dm, dmf, *idx, 20.0, 0.73, K)
ctx.map_upload(m); ctx.map_predict_indices(T, tick, 20.0, BIG); ctx.map_fuse(T, tick, 20.0, 0.73)
got = ctx.map_download()
close = np.isclose(got, fused, rtol=2e-6, atol=1e-7)
bad = np.where(~close.all(axis=1))[0]
print("bad rows", len(bad), "changed rows oracle", (fused != m).any(axis=1).sum(), "changed rows gpu", (got != m).any(axis=1).sum())
print("bad columns histogram", (~close[bad]).sum(axis=0))
for r in bad[:6]:
    print(r, "orig ", m[r]); print(r, "oracle", fused[r]); print(r, "gpu   ", got[r])
upd_o = (fused != m).any(axis=1); upd_g = (got != m).any(axis=1)
print("updated only oracle", (upd_o & ~upd_g).sum(), "only gpu", (~upd_o & upd_g).sum())
