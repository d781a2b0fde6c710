"""Pins the CPU oracle (tracking half) against golden vectors produced by the REFERENCE's own CUDA kernels.

tests/golden/ref_tracking_160x120.npz was generated on a B200 by tests/golden/make_golden.py from oracle/_ref/libef_ref.so
(reference Core/Cuda/reduce.cu + cudafuncs.cu, unmodified). Tolerances cover only what legitimately differs between the
reference build and the oracle: FMA contraction and approximate rsqrtf on the GPU, and summation order."""
import os

import numpy as np
import pytest

from util import rel_err

G_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tracking_160x120.npz")


@pytest.fixture(scope="module")
def gold():
    from elasticfusion_b200 import synth
    from oracle import ef_oracle as eo

    g = dict(np.load(G_PATH))
    w, h, fx, fy, cx, cy = g["K"]
    K = synth.Intrinsics(int(w), int(h), fx, fy, cx, cy)
    od = eo.Odometry(K.width, K.height, K.cx, K.cy, K.fx, K.fy)
    od.init_first_rgb(g["prev_rgba"])
    od.init_icp_model(g["vtx"], g["nrm"], g["T"])
    od.init_rgb_model(g["img"])
    od.init_icp_depth(g["filt"], 20.0)
    od.init_rgb(g["rgba"])
    return g, K, od


@pytest.mark.parametrize("lv", (0, 1, 2))
def test_pyramids_match_reference(gold, lv):
    g, K, od = gold
    assert np.array_equal(od.buffer("depth_tmp", lv), g[f"depth_tmp{lv}"])
    for name in ("vmap_curr", "nmap_curr", "vmap_g_prev", "nmap_g_prev", "lastDepth", "nextDepth"):
        a = od.buffer(name, lv)
        nan_ref = np.unpackbits(g[f"{name}{lv}_nan"])[:a.size].reshape(a.shape).astype(bool)
        if name.startswith(("vmap", "nmap")):  # only the x plane flags validity
            r = a.shape[0] // 3
            assert (np.isnan(a[:r]) != nan_ref[:r]).mean() < 1e-4, name
            ok = ~np.isnan(a[:r]) & ~nan_ref[:r]
            both = np.concatenate([ok, ok, ok], 0)
            s = np.array([a[both].astype(np.float64).sum(), np.abs(a[both].astype(np.float64)).sum()])
        else:
            assert np.array_equal(np.isnan(a), nan_ref), name
            s = np.array([np.nansum(a.astype(np.float64)), np.nansum(np.abs(a.astype(np.float64)))])
        assert abs(s[1] - g[f"{name}{lv}_sum"][1]) <= 2e-4 * g[f"{name}{lv}_sum"][1], (name, s, g[f"{name}{lv}_sum"])
    for name in ("lastImage", "nextImage", "lastNextImage"):
        a, b = od.buffer(name, lv).astype(int), g[f"{name}{lv}"].astype(int)
        assert np.abs(a - b).max() <= 1 and (a != b).mean() < 2e-3, name


@pytest.mark.parametrize("lv", (0, 1, 2))
def test_reductions_match_reference(gold, lv):
    from oracle import ef_oracle as eo

    g, K, od = gold
    d = np.float32(1 << lv)
    fx, fy, cx, cy = np.float32(K.fx) / d, np.float32(K.fy) / d, np.float32(K.cx) / d, np.float32(K.cy) / d
    ang = float(np.sin(np.float32(20.0) * np.float32(3.14159254) / np.float32(180.0)))
    A, b, r = eo.icp_step(g["Rc"], g["tc"], od.buffer("vmap_curr", lv), od.buffer("nmap_curr", lv), g["Rpi"], g["tp"], fx, fy, cx, cy,
                          od.buffer("vmap_g_prev", lv), od.buffer("nmap_g_prev", lv), 0.10, ang)
    assert abs(r[1] - g[f"icp_r{lv}"][1]) <= max(2, 1e-3 * g[f"icp_r{lv}"][1])
    assert rel_err(A, g[f"icp_A{lv}"]) < 2e-3 and rel_err(b, g[f"icp_b{lv}"]) < 2e-3
    dIdx, dIdy = eo.sobel(od.buffer("nextImage", lv))
    assert (dIdx != g[f"dIdx{lv}"]).mean() < 0.02 and np.abs(dIdx.astype(int) - g[f"dIdx{lv}"]).max() <= 2
    min_scale = np.float32(((5, 3, 1)[lv] ** 2) / 0.125 ** 2)
    corres, sig, cnt = eo.rgb_residual(min_scale, dIdx, dIdy, od.buffer("lastDepth", lv), od.buffer("nextDepth", lv), od.buffer("lastImage", lv),
                                       od.buffer("nextImage", lv), 0.07, g[f"kt{lv}"], g[f"krk{lv}"])
    assert abs(cnt - g[f"res{lv}"][1]) <= max(3, 0.02 * g[f"res{lv}"][1]), (cnt, g[f"res{lv}"])
    cloud = eo.project_points(od.buffer("lastDepth", lv), fx, fy, cx, cy)
    Ar, br = eo.rgb_step(corres, float(np.sqrt(np.float32(g[f"res{lv}"][1]))), cloud, fx, fy, dIdx, dIdy, 0.125)
    assert rel_err(Ar, g[f"rgb_A{lv}"]) < 0.05  # correspondence sets differ by the intensity rounding flips above


def test_so3_step_matches_reference(gold):
    from oracle import ef_oracle as eo

    g, K, od = gold
    A, b, r = eo.so3_step(od.buffer("lastNextImage", 2), od.buffer("nextImage", 2), g["so3_H"], g["so3_kinv"], g["so3_krlr"])
    assert r[1] == g["so3_r"][1]
    assert rel_err(A, g["so3_A"]) < 2e-2 and abs(r[0] - g["so3_r"][0]) < 2e-2 * g["so3_r"][0]


def test_full_track_matches_reference(gold):
    """getIncrementalTransformation: oracle host loop + kernels vs the reference kernels driven by the harness (Eigen host math)."""
    from oracle import ef_oracle as eo

    g, K, od = gold
    T, tr = od.track(g["T"])
    ref_tr = g["track_trace"].view(eo.TRACE_DTYPE)
    assert len(tr) == len(ref_tr)
    for a, b in zip(tr, ref_tr):
        assert (a["kind"], a["level"], a["iter"]) == (b["kind"], b["level"], b["iter"])
    first = [i for i, t in enumerate(ref_tr) if t["kind"] == 0][0]
    assert rel_err(tr[first]["A_icp"], ref_tr[first]["A_icp"]) < 5e-3
    assert np.abs(T[:3, 3] - g["track_T"][:3, 3]).max() < 1e-4 and np.abs(T[:3, :3] - g["track_T"][:3, :3]).max() < 1e-4
