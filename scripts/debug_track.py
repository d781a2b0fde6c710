import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, '/root/repo/tests')
from elasticfusion_b200 import synth, capi
from oracle import ef_oracle as eo
from util import run_oracle, rgba_of, rel_err
K = synth.K_DEFAULT
frames = list(synth.sequence(5, K, seed=42, noise=True))
for cfg in (dict(), dict(rgb_only=True, so3=False)):
    f = run_oracle(frames, K, 3)
    rgb, depth, _ = frames[3]
    filt = eo.bilateral(depth, 3.0)
    od = f.odometry(); T_prev = f.pose
    vtx, nrm, img = f.buffer("fill_vertex"), f.buffer("fill_normal"), f.buffer("fill_image")
    od.init_icp_model(vtx, nrm, T_prev); od.init_rgb_model(img); od.init_icp_depth(filt, 20.0); od.init_rgb(rgba_of(rgb))
    ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=400000))
    ctx.upload("FILL_VERTEX", vtx); ctx.upload("FILL_NORMAL", nrm); ctx.upload("FILL_IMAGE", img); ctx.upload("DEPTH_FILTERED", filt); ctx.upload("RGBA", rgba_of(rgb))
    for lv in range(3): ctx.upload("LAST_NEXT_IMAGE", od.buffer("lastNextImage", lv), level=lv)
    ctx.odom_init_icp_model(ctx.buffer_ptr("FILL_VERTEX")[0], ctx.buffer_ptr("FILL_NORMAL")[0], T_prev)
    ctx.odom_init_rgb_model(ctx.buffer_ptr("FILL_IMAGE")[0]); ctx.odom_init_icp_depth(ctx.buffer_ptr("DEPTH_FILTERED")[0], 20.0); ctx.odom_init_rgb(ctx.buffer_ptr("RGBA")[0])
    To, tro = od.track(T_prev, **cfg); Tp, trp = ctx.odom_track(T_prev, **cfg)
    print("cfg", cfg, "records", len(tro), len(trp), "pose diff", np.abs(To-Tp).max())
    for a, b in zip(trp, tro):
        if a["kind"] == 1:
            print(" so3", a["iter"], b["iter"], a["so3_residual"], b["so3_residual"], rel_err(a["A_so3"], b["A_so3"]))
        else:
            print(" se3 L%d i%d | L%d i%d cnt %d/%d icpn %g/%g sig %g/%g relA %.2e relArgb %.2e relAicp %.2e db %.2e dres %.2e" % (a["level"], a["iter"], b["level"], b["iter"], a["rgb_count"], b["rgb_count"], a["icp_residual"][1], b["icp_residual"][1], a["sigma_val"], b["sigma_val"], rel_err(a["lastA"], b["lastA"]), rel_err(a["A_rgb"], b["A_rgb"]) if np.abs(b["A_rgb"]).max()>0 else 0, rel_err(a["A_icp"], b["A_icp"]) if np.abs(b["A_icp"]).max()>0 else 0, np.abs(a["lastb"]-b["lastb"]).max()/max(np.abs(b["lastb"]).max(),1e-30), np.abs(a["result"]-b["result"]).max()))
    ctx.close()
