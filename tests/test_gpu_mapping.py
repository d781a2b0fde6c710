"""GPU parity tests of the mapping half + depth preprocess + the whole-frame pipeline, through the C ABI, vs the oracle."""
import numpy as np
import pytest

from util import assert_same, rel_err, run_oracle

pytestmark = pytest.mark.gpu

MAXD = 20.0
BIG = 2147483647 // 2


def make_ctx(K, **kw):
    from elasticfusion_b200 import capi

    kw.setdefault("capacity", 500000)
    kw.setdefault("time_delta", BIG)
    return capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, **kw))


def test_preprocess_depth(frames, K):
    """13x13 bilateral + both metric conversions in one launch. The filter output is an integer rounding of a ratio of
    exp-weighted sums; libm expf (oracle) and CUDA expf differ by <= 2 ulp, so allow 1 mm flips on <= 1e-4 of the pixels."""
    from oracle import ef_oracle as eo

    ctx = make_ctx(K)
    try:
        depth = frames[0][1]
        ctx.upload("DEPTH_RAW", depth)
        ctx.preprocess_depth(ctx.buffer_ptr("DEPTH_RAW")[0], 3.0, ctx.buffer_ptr("DEPTH_FILTERED")[0],
                             ctx.buffer_ptr("DEPTH_METRIC")[0], ctx.buffer_ptr("DEPTH_METRIC_FILTERED")[0])
        filt = ctx.download("DEPTH_FILTERED")
        ref = eo.bilateral(depth, 3.0)
        diff = np.abs(filt.astype(np.int32) - ref.astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() <= 1e-4, (diff.max(), (diff > 0).mean())
        assert_same(ctx.download("DEPTH_METRIC"), eo.metric(depth, 3.0), "metric raw")
        assert_same(ctx.download("DEPTH_METRIC_FILTERED"), eo.metric(filt, 3.0), "metric filtered")
        assert (filt > 0).mean() > 0.5
    finally:
        ctx.close()


@pytest.fixture(scope="module")
def mstate(frames, K):
    """Oracle after 4 frames (a map with updated + fresh surfels) and frame 4's preprocessed inputs."""
    from oracle import ef_oracle as eo

    f = run_oracle(frames, K, 4)
    rgb, depth, _ = frames[4]
    filt = eo.bilateral(depth, 3.0)
    st = dict(f=f, rgb=rgb, depth=depth, filt=filt, dm=eo.metric(depth, 3.0), dmf=eo.metric(filt, 3.0), map=f.map(), T=f.pose,
              tick=f.tick, K=K)
    ctx = make_ctx(K)
    ctx.upload("RGB", rgb)
    ctx.upload("DEPTH_RAW", depth)
    ctx.upload("DEPTH_FILTERED", filt)
    ctx.upload("DEPTH_METRIC", st["dm"])
    ctx.upload("DEPTH_METRIC_FILTERED", st["dmf"])
    ctx.map_upload(st["map"])
    st["ctx"] = ctx
    yield st
    ctx.close()


def test_first_frame_initialise(frames, K):
    from oracle import ef_oracle as eo

    rgb, depth, _ = frames[0]
    filt = eo.bilateral(depth, 3.0)
    dm, dmf = eo.metric(depth, 3.0), eo.metric(filt, 3.0)
    raw = eo.feedback_buffer(rgb, dm, K, 1, MAXD)
    fil = eo.feedback_buffer(rgb, dmf, K, 1, MAXD)
    ref = eo.map_initialise(raw, fil)
    ctx = make_ctx(K)
    try:
        ctx.upload("RGB", rgb)
        ctx.upload("DEPTH_METRIC", dm)
        ctx.upload("DEPTH_METRIC_FILTERED", dmf)
        ctx.map_initialise()
        got = ctx.map_download()
        assert len(got) == len(ref) and len(ref) > 100000
        cols = [0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11]
        assert_same(got[:, cols], ref[:, cols], "initial surfels")
        # confidence = exp(-r^2/0.72): expf implementations differ by ulps
        assert rel_err(got[:, 3], ref[:, 3]) < 1e-6
    finally:
        ctx.close()


def test_predict_indices(mstate):
    from oracle import ef_oracle as eo

    ctx, K = mstate["ctx"], mstate["K"]
    ref = eo.predict_indices(mstate["map"], mstate["T"], mstate["tick"], MAXD, BIG, K)
    ctx.map_predict_indices(mstate["T"], mstate["tick"], MAXD, BIG)
    for name, r in zip(("INDEX", "VERT_CONF", "COLOR_TIME", "NORM_RAD"), ref):
        assert_same(ctx.download(name), r, name)
    assert (ref[0] > 0).mean() > 0.3


def test_fuse_then_clean(mstate):
    """data association + merge + new-surfel emission, then the index map on the updated model and the clean pass."""
    from oracle import ef_oracle as eo

    ctx, K, T, tick = mstate["ctx"], mstate["K"], mstate["T"], mstate["tick"]
    w = 0.73
    idx = eo.predict_indices(mstate["map"], T, tick, MAXD, BIG, K)
    fused, new = eo.fuse(mstate["map"], T, tick, mstate["rgb"], mstate["dm"], mstate["dmf"], *idx, MAXD, w, K)
    ctx.map_upload(mstate["map"])
    ctx.map_predict_indices(T, tick, MAXD, BIG)
    ctx.map_fuse(T, tick, MAXD, w)
    got, got_new = ctx.map_download(), ctx.map_download_new()
    changed = (fused != mstate["map"]).any(axis=1).sum()
    assert changed > 1000 and len(new) > 100
    # acosf (normal-angle gate) and expf (confidence) are the only non-IEEE-exact operations: a handful of surfels
    # may associate differently; everything else must agree bit for bit.
    assert len(got_new) == len(new) or abs(len(got_new) - len(new)) <= 2
    if len(got_new) == len(new):
        cols = [0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11]
        assert_same(got_new[:, cols], new[:, cols], "new unstable surfels")
        assert rel_err(got_new[:, 3], new[:, 3]) < 1e-6
    same_rows = ((got == fused) | (np.isnan(got) & np.isnan(fused))).all(axis=1)
    close_rows = np.isclose(got, fused, rtol=2e-6, atol=1e-7, equal_nan=True).all(axis=1)
    assert (~close_rows).sum() <= 2, f"{(~close_rows).sum()} surfels differ after fuse"
    assert same_rows.mean() > 0.9

    # clean on the oracle's fused state so the comparison stays exact
    ctx.map_upload(fused)
    idx2 = eo.predict_indices(fused, T, tick, MAXD, BIG, K)
    ref = eo.clean(fused, new, T, tick, *idx2, 10.0, BIG, MAXD, K)
    ctx.map_predict_indices(T, tick, MAXD, BIG)
    if len(got_new) == len(new):
        ctx.map_clean(T, tick, 10.0, BIG, MAXD)
        out = ctx.map_download()
        assert len(out) == len(ref)
        cols = [0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11]
        assert_same(out[:, cols], ref[:, cols], "map after clean")
        assert rel_err(out[:, 3], ref[:, 3]) < 1e-6


def test_raycast_and_fill_in(frames, K):
    """combinedPredict on a map with stable surfels (confidence forced above the threshold), all four attachments,
    then the three fill-in passes and the density test."""
    from oracle import ef_oracle as eo

    f = run_oracle(frames, K, 3)
    m = f.map()
    m[:, 3] += 10.0  # make the surfels stable so the raycast renders them
    T, tick = f.pose, f.tick
    ref = eo.combined_predict(m, T, MAXD, 10.0, tick, tick, BIG, K)
    assert (ref[1][..., 2] > 0).mean() > 0.5
    ctx = make_ctx(K)
    try:
        ctx.map_upload(m)
        ctx.map_raycast(T, MAXD, 10.0, tick, tick, BIG, 0)
        for name, r in zip(("IMAGE", "VERTEX", "NORMAL", "TIME"), ref):
            assert_same(ctx.download(name), r, name)
        d_ref = eo.combined_predict(m, T, MAXD, 10.0, tick, tick, BIG, K, depth_only=True)
        ctx.map_raycast(T, MAXD, 10.0, tick, tick, BIG, 2)
        assert_same(ctx.download("SYNTH_DEPTH"), d_ref, "synthesizeDepth")
        rgb, depth, _ = frames[3]
        filt = eo.bilateral(depth, 3.0)
        ctx.upload("RGB", rgb)
        ctx.upload("DEPTH_FILTERED", filt)
        ctx.map_fill_in(False, False)
        assert_same(ctx.download("FILL_VERTEX"), eo.fill_vertex(ref[1], filt, 0, K), "fill vertex")
        assert_same(ctx.download("FILL_NORMAL"), eo.fill_normal(ref[2], filt, 0, K), "fill normal")
        assert_same(ctx.download("FILL_IMAGE"), eo.fill_image(ref[0], rgb, 0), "fill image")
        assert ctx.dense_enough() == eo.dense_enough(ref[0])
    finally:
        ctx.close()


def test_pipeline_matches_oracle(frames, K):
    """processFrame over the 8-frame sequence: per-frame pose within 2e-5 m of the oracle run and surfel counts within
    0.1 % (the two runs differ only by float summation order in the reductions and ulp-level expf/acosf)."""
    f = run_oracle(frames, K, 0)
    ctx = make_ctx(K, skip_mid_predict=0)
    try:
        for i, (rgb, depth, _) in enumerate(frames):
            f.process_frame(rgb, depth, i * 33333)
            ctx.process_frame(rgb, depth, i * 33333)
            Tp, To = ctx.get_pose(), f.pose
            assert np.abs(Tp[:3, 3] - To[:3, 3]).max() < 2e-5, (i, Tp[:3, 3], To[:3, 3])
            assert np.abs(Tp[:3, :3] - To[:3, :3]).max() < 2e-5
            assert abs(ctx.map_count() - f.count) <= max(2, 1e-3 * f.count), (i, ctx.map_count(), f.count)
        assert ctx.get_tick() == f.tick
    finally:
        ctx.close()


def test_sequence_ate_vs_oracle(K):
    """BASELINE.json parity bar for the whole path: estimated SE(3) trajectory within 1e-3 m ATE of the oracle run on the
    same 60-frame noisy synthetic sequence (model-based tracking kicks in once surfels become stable), surfel count within
    0.5 %, and both runs agree with ground truth to the same accuracy."""
    from elasticfusion_b200 import synth

    n = 60
    frames = list(synth.sequence(n, K, seed=11, noise=True))
    f = run_oracle(frames, K, 0, capacity=600000)
    ctx = make_ctx(K, capacity=600000)
    est_p, est_o = [], []
    try:
        for i, (rgb, depth, _) in enumerate(frames):
            f.process_frame(rgb, depth, i * 33333)
            ctx.process_frame(rgb, depth, i * 33333)
            est_p.append(ctx.get_pose())
            est_o.append(f.pose)
        est_p, est_o = np.array(est_p), np.array(est_o)
        gt = np.array([fr[2] for fr in frames])
        ate_po = synth.ate_rmse(est_p, est_o)
        assert ate_po < 1e-3, ate_po
        assert abs(synth.ate_rmse(est_p, gt) - synth.ate_rmse(est_o, gt)) < 1e-3
        assert abs(ctx.map_count() - f.count) <= 5e-3 * f.count, (ctx.map_count(), f.count)
        assert np.abs(est_p[-1][:3, :3] - est_o[-1][:3, :3]).max() < 1e-3
    finally:
        ctx.close()


def test_lookahead_matches_plain_calls(frames, K):
    """ef_prefetch_frame (host and device variants): staging the next frame on the side stream while the current one is in
    flight must not change anything -- poses, surfel counts and the final map are bit-identical to the plain per-frame calls.
    Also the state machine: consuming without a pending frame, or passing a frame while one is pending, is EF_ESTATE."""
    import torch
    from elasticfusion_b200 import capi

    def run(mode):
        ctx = make_ctx(K)
        poses, counts = [], []
        try:
            if mode == "plain":
                for i, (rgb, depth, _) in enumerate(frames):
                    ctx.process_frame(rgb, depth, i)
                    poses.append(ctx.get_pose())
                    counts.append(ctx.map_count())
            elif mode == "host":
                with pytest.raises(capi.EfError):
                    ctx.process_frame(None, None, 0)  # nothing staged yet
                ctx.prefetch_frame(frames[0][0], frames[0][1])
                with pytest.raises(capi.EfError):
                    ctx.prefetch_frame(frames[0][0], frames[0][1])  # one pending frame at most
                with pytest.raises(capi.EfError):
                    ctx.process_frame(frames[0][0], frames[0][1], 0)  # the staged frame has to be consumed first
                for i in range(len(frames)):
                    ctx.process_frame_device(None, None, i)
                    if i + 1 < len(frames):
                        ctx.prefetch_frame(frames[i + 1][0], frames[i + 1][1])
                    ctx.finish_frame()
                    poses.append(ctx.get_pose())
                    counts.append(ctx.map_count())
            else:
                dev = [(torch.from_numpy(np.ascontiguousarray(r)).cuda(), torch.from_numpy(np.ascontiguousarray(d).view(np.int16)).cuda())
                       for r, d, _ in frames]
                torch.cuda.synchronize()
                ctx.prefetch_frame_device(dev[0][0].data_ptr(), dev[0][1].data_ptr())
                for i in range(len(frames)):
                    ctx.process_frame_device(None, None, i)
                    if i + 1 < len(frames):
                        ctx.prefetch_frame_device(dev[i + 1][0].data_ptr(), dev[i + 1][1].data_ptr())
                ctx.finish_frame()
                poses.append(ctx.get_pose())
                counts.append(ctx.map_count())
            return np.array(poses), counts, ctx.map_download()
        finally:
            ctx.close()

    p0, c0, m0 = run("plain")
    p1, c1, m1 = run("host")
    p2, c2, m2 = run("device")
    assert np.array_equal(p0, p1) and c0 == c1
    assert np.array_equal(m0, m1, equal_nan=True)
    assert np.array_equal(p0[-1], p2[-1]) and c0[-1] == c2[-1]
    assert np.array_equal(m0, m2, equal_nan=True)
