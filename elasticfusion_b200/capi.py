"""ctypes binding of libefusion.so (include/efusion_b200.h).

This is the thinnest possible host layer: it loads the in-tree shared library that holds the sm_100a kernels and
calls its C ABI. There is no CPU fallback: if the library is missing or the call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libefusion.so")
_LIB = None


class EfError(RuntimeError):
    pass


class EfConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("time_delta", C.c_int32), ("count_thresh", C.c_int32), ("err_thresh", C.c_float),
                ("cov_thresh", C.c_float), ("close_loops", C.c_int32), ("iclnuim", C.c_int32), ("reloc", C.c_int32),
                ("photo_thresh", C.c_float), ("confidence", C.c_float), ("depth_cutoff", C.c_float),
                ("icp_weight", C.c_float), ("fast_odom", C.c_int32), ("fern_thresh", C.c_float), ("so3", C.c_int32),
                ("frame_to_frame_rgb", C.c_int32), ("capacity", C.c_int32), ("device", C.c_int32),
                ("skip_mid_predict", C.c_int32)]


class EfLoopResult(C.Structure):
    _fields_ = [("ran", C.c_int32), ("accepted", C.c_int32), ("n_constraints", C.c_int32), ("lastICPError", C.c_float),
                ("lastICPCount", C.c_float), ("cov_diag", C.c_double * 6), ("T_wc_est", C.c_double * 16)]


TRACE_DTYPE = np.dtype([
    ("kind", "<i4"), ("level", "<i4"), ("iter", "<i4"), ("rgb_count", "<i4"), ("rgb_sigma", "<i4"),
    ("sigma_val", "<f4"),
    ("A_icp", "<f4", (36,)), ("b_icp", "<f4", (6,)), ("icp_residual", "<f4", (2,)),
    ("A_rgb", "<f4", (36,)), ("b_rgb", "<f4", (6,)),
    ("A_so3", "<f4", (9,)), ("b_so3", "<f4", (3,)), ("so3_residual", "<f4", (2,)),
    ("lastA", "<f8", (36,)), ("lastb", "<f8", (6,)), ("result", "<f8", (6,)),
], align=True)

STATS_DTYPE = np.dtype([("lastICPError", "<f4"), ("lastICPCount", "<f4"), ("lastRGBError", "<f4"),
                        ("lastRGBCount", "<f4"), ("lastSO3Error", "<f4"), ("lastSO3Count", "<f4"),
                        ("lastA", "<f8", (36,)), ("lastb", "<f8", (6,))], align=True)

DATATERM_DTYPE = np.dtype([("zero_x", "<i2"), ("zero_y", "<i2"), ("one_x", "<i2"), ("one_y", "<i2"),
                           ("diff", "<f4"), ("valid", "<i4")])

# buffer ids (include/efusion_b200.h)
BUF = dict(RGB=0, DEPTH_RAW=1, DEPTH_FILTERED=2, DEPTH_METRIC=3, DEPTH_METRIC_FILTERED=4, RGBA=5, INDEX=10, VERT_CONF=11,
           COLOR_TIME=12, NORM_RAD=13, IMAGE=14, VERTEX=15, NORMAL=16, TIME=17, OLD_IMAGE=18, OLD_VERTEX=19,
           OLD_NORMAL=20, OLD_TIME=21, SYNTH_DEPTH=22, FILL_IMAGE=30, FILL_VERTEX=31, FILL_NORMAL=32, VMAP_CURR=40,
           NMAP_CURR=41, VMAP_G_PREV=42, NMAP_G_PREV=43, LAST_DEPTH=44, NEXT_DEPTH=45, LAST_IMAGE=46, NEXT_IMAGE=47,
           LAST_NEXT_IMAGE=48, DIDX=49, DIDY=50, DEPTH_TMP=51, CORRES=52, VMAPS_TMP=53)

_BUF_FMT = {  # id -> (dtype, channels/planes kind)
    0: (np.uint8, "c3"), 1: (np.uint16, "c1"), 2: (np.uint16, "c1"), 3: (np.float32, "c1"), 4: (np.float32, "c1"),
    5: (np.uint8, "c4"), 10: (np.uint32, "c1"), 11: (np.float32, "c4"), 12: (np.float32, "c4"), 13: (np.float32, "c4"),
    14: (np.uint8, "c4"), 15: (np.float32, "c4"), 16: (np.float32, "c4"), 17: (np.uint16, "c1"), 18: (np.uint8, "c4"),
    19: (np.float32, "c4"), 20: (np.float32, "c4"), 21: (np.uint16, "c1"), 22: (np.float32, "c1"),
    30: (np.uint8, "c4"), 31: (np.float32, "c4"), 32: (np.float32, "c4"),
    40: (np.float32, "p3"), 41: (np.float32, "p3"), 42: (np.float32, "p3"), 43: (np.float32, "p3"),
    44: (np.float32, "c1"), 45: (np.float32, "c1"), 46: (np.uint8, "c1"), 47: (np.uint8, "c1"), 48: (np.uint8, "c1"),
    49: (np.int16, "c1"), 50: (np.int16, "c1"), 51: (np.uint16, "c1"), 52: (DATATERM_DTYPE, "c1"),
    53: (np.float32, "c4"),
}


def lib():
    """Loads libefusion.so; raises if the CUDA extension has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("EF_LIB", LIB_PATH)  # EF_LIB: an instrumented build of the same library (scripts/phase_profile.py)
        if not os.path.exists(path):
            raise EfError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` or ./build.sh")
        _LIB = C.CDLL(path)
        _LIB.ef_error_string.restype = C.c_char_p
        _LIB.ef_stream.restype = C.c_void_p
    return _LIB


def _chk(rc):
    if rc != 0:
        raise EfError(f"libefusion call failed ({rc}): {lib().ef_error_string(rc).decode()}")


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _f(x):
    return C.c_float(float(x))


def _T(T):
    return None if T is None else np.ascontiguousarray(T, np.float64)


def default_config(width, height, fx, fy, cx, cy, **overrides) -> EfConfig:
    cfg = EfConfig()
    lib().ef_default_config(C.byref(cfg), width, height, _f(fx), _f(fy), _f(cx), _f(cy))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    return cfg


class Context:
    """One EfContext: one device, one stream. Mirrors the stage API of include/efusion_b200.h."""

    def __init__(self, cfg: EfConfig, stream: int | None = None):
        self.cfg = cfg
        self.w, self.h = cfg.width, cfg.height
        self.h_ctx = C.c_void_p()
        _chk(lib().ef_create(C.byref(cfg), C.c_void_p(stream) if stream else None, C.byref(self.h_ctx)))

    def close(self):
        if getattr(self, "h_ctx", None):
            lib().ef_destroy(self.h_ctx)
            self.h_ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- named buffers
    def _shape(self, bid, level):
        dt, kind = _BUF_FMT[bid % 100]
        r, c = (self.h >> level, self.w >> level) if bid % 100 >= 40 and bid % 100 != 53 else (self.h, self.w)
        if kind == "c1":
            return dt, (r, c)
        if kind == "p3":
            return dt, (3 * r, c)
        return dt, (r, c, int(kind[1]))

    def buffer_ptr(self, name, level=0, which=0):
        bid = BUF[name] + (100 * which if BUF[name] >= 40 else 0)
        ptr = C.c_void_p()
        nbytes = C.c_size_t()
        _chk(lib().ef_buffer(self.h_ctx, bid, level, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def download(self, name, level=0, which=0):
        bid = BUF[name] + (100 * which if BUF[name] >= 40 else 0)
        dt, shape = self._shape(bid, level)
        out = np.zeros(shape, dt)
        _chk(lib().ef_download(self.h_ctx, bid, level, _p(out), C.c_size_t(out.nbytes)))
        return out

    def upload(self, name, arr, level=0, which=0):
        bid = BUF[name] + (100 * which if BUF[name] >= 40 else 0)
        dt, shape = self._shape(bid, level)
        a = np.ascontiguousarray(arr, dt)
        assert a.shape == tuple(shape), (a.shape, shape)
        _chk(lib().ef_upload(self.h_ctx, bid, level, _p(a), C.c_size_t(a.nbytes)))

    def sync(self):
        _chk(lib().ef_sync(self.h_ctx))

    @property
    def stream(self):
        return lib().ef_stream(self.h_ctx)

    def launch_count(self):
        n = C.c_int64()
        _chk(lib().ef_launch_count(self.h_ctx, C.byref(n)))
        return n.value

    # ---- whole frame
    def process_frame(self, rgb, depth, timestamp=0, weight_multiplier=1.0, T_wc=None):
        """rgb = depth = None consumes the frame staged by prefetch_frame()."""
        if rgb is None and depth is None:
            _chk(lib().ef_process_frame(self.h_ctx, None, None, C.c_int64(timestamp), _f(weight_multiplier), _p(_T(T_wc))))
            return
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        _chk(lib().ef_process_frame(self.h_ctx, _p(rgb), _p(depth), C.c_int64(timestamp), _f(weight_multiplier), _p(_T(T_wc))))

    def prefetch_frame(self, rgb, depth):
        """Look-ahead: stage + preprocess the NEXT frame (host arrays) on the side stream."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        _chk(lib().ef_prefetch_frame(self.h_ctx, _p(rgb), _p(depth)))

    def prefetch_frame_device(self, rgb_ptr, depth_ptr):
        _chk(lib().ef_prefetch_frame_device(self.h_ctx, C.c_void_p(rgb_ptr), C.c_void_p(depth_ptr)))

    def finish_frame(self):
        _chk(lib().ef_finish_frame(self.h_ctx))

    def process_frame_device(self, rgb_ptr, depth_ptr, timestamp=0, weight_multiplier=1.0, T_wc=None):
        _chk(lib().ef_process_frame_device(self.h_ctx, C.c_void_p(rgb_ptr), C.c_void_p(depth_ptr), C.c_int64(timestamp),
                                           _f(weight_multiplier), _p(_T(T_wc))))

    def process_frame_begin(self, rgb, depth, timestamp=0, weight_multiplier=1.0, T_wc=None):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        _chk(lib().ef_process_frame_begin(self.h_ctx, _p(rgb), _p(depth), C.c_int64(timestamp), _f(weight_multiplier), _p(_T(T_wc))))

    def process_frame_end(self, T_override=None, nodes=None, fern_accepted=False):
        nd = None if nodes is None else np.ascontiguousarray(nodes, np.float32).reshape(-1, 16)
        _chk(lib().ef_process_frame_end(self.h_ctx, _p(_T(T_override)), _p(nd), 0 if nd is None else len(nd), int(fern_accepted)))

    def local_loop_result(self):
        """(info dict, src (n,3), dst (n,3), times (n,)) of the last frame's local loop closure front half."""
        res = EfLoopResult()
        cap = (self.w // 20) * (self.h // 20)
        src = np.zeros((cap, 3), np.float64)
        dst = np.zeros((cap, 3), np.float64)
        tm = np.zeros(cap, np.int32)
        n = C.c_int32()
        _chk(lib().ef_local_loop_result(self.h_ctx, C.byref(res), _p(src), _p(dst), _p(tm), cap, C.byref(n)))
        info = dict(ran=res.ran, accepted=res.accepted, n_constraints=res.n_constraints, lastICPError=res.lastICPError,
                    lastICPCount=res.lastICPCount, cov_diag=np.array(res.cov_diag[:]), T_wc_est=np.array(res.T_wc_est[:]).reshape(4, 4))
        return info, src[:n.value].copy(), dst[:n.value].copy(), tm[:n.value].copy()

    def predict(self):
        _chk(lib().ef_predict(self.h_ctx))

    def get_pose(self):
        T = np.zeros((4, 4), np.float64)
        _chk(lib().ef_get_pose(self.h_ctx, _p(T)))
        return T

    def set_pose(self, T):
        _chk(lib().ef_set_pose(self.h_ctx, _p(_T(T))))

    def get_tick(self):
        t = C.c_int32()
        _chk(lib().ef_get_tick(self.h_ctx, C.byref(t)))
        return t.value

    def set_tick(self, t):
        _chk(lib().ef_set_tick(self.h_ctx, int(t)))

    def set(self, **kw):
        fns = dict(rgb_only=("ef_set_rgb_only", int), icp_weight=("ef_set_icp_weight", _f), pyramid=("ef_set_pyramid", int),
                   fast_odom=("ef_set_fast_odom", int), so3=("ef_set_so3", int),
                   frame_to_frame_rgb=("ef_set_frame_to_frame_rgb", int),
                   confidence_threshold=("ef_set_confidence_threshold", _f), depth_cutoff=("ef_set_depth_cutoff", _f))
        for k, v in kw.items():
            name, conv = fns[k]
            _chk(getattr(lib(), name)(self.h_ctx, conv(v)))

    # ---- tracker stages
    def odom_init_icp_depth(self, depth_ptr, cutoff, which=0):
        _chk(lib().ef_odom_init_icp_depth(self.h_ctx, which, C.c_void_p(depth_ptr), _f(cutoff)))

    def odom_init_icp_pred(self, vtx_ptr, nrm_ptr, which=0):
        _chk(lib().ef_odom_init_icp_pred(self.h_ctx, which, C.c_void_p(vtx_ptr), C.c_void_p(nrm_ptr)))

    def odom_init_icp_model(self, vtx_ptr, nrm_ptr, T_wc, which=0):
        _chk(lib().ef_odom_init_icp_model(self.h_ctx, which, C.c_void_p(vtx_ptr), C.c_void_p(nrm_ptr), _p(_T(T_wc))))

    def odom_init_rgb(self, rgba_ptr, which=0):
        _chk(lib().ef_odom_init_rgb(self.h_ctx, which, C.c_void_p(rgba_ptr)))

    def odom_init_rgb_model(self, rgba_ptr, which=0):
        _chk(lib().ef_odom_init_rgb_model(self.h_ctx, which, C.c_void_p(rgba_ptr)))

    def odom_init_first_rgb(self, rgba_ptr, which=0):
        _chk(lib().ef_odom_init_first_rgb(self.h_ctx, which, C.c_void_p(rgba_ptr)))

    def odom_track(self, T_wc, rgb_only=False, icp_weight=10.0, pyramid=True, fast_odom=False, so3=True, which=0, max_trace=48):
        T = _T(T_wc).copy()
        trace = np.zeros(max_trace, TRACE_DTYPE)
        n = C.c_int32()
        _chk(lib().ef_odom_track(self.h_ctx, which, _p(T), int(rgb_only), _f(icp_weight), int(pyramid), int(fast_odom), int(so3),
                                 _p(trace), max_trace, C.byref(n)))
        return T, trace[:n.value]

    def odom_stats(self, which=0):
        st = np.zeros(1, STATS_DTYPE)
        _chk(lib().ef_odom_stats(self.h_ctx, which, _p(st)))
        return st[0]

    def odom_covariance(self, which=0):
        cov = np.zeros((6, 6), np.float64)
        _chk(lib().ef_odom_covariance(self.h_ctx, which, _p(cov)))
        return cov

    def icp_step(self, level, Rcurr, tcurr, Rprev_inv, tprev, which=0):
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        A, b, res = np.zeros((6, 6), np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32)
        Rc, tc, Rp, tp = f32(Rcurr), f32(tcurr), f32(Rprev_inv), f32(tprev)
        _chk(lib().ef_icp_step(self.h_ctx, which, level, _p(Rc), _p(tc), _p(Rp), _p(tp), _p(A), _p(b), _p(res)))
        return A, b, res

    def icp_step_async(self, level, Rcurr=None, tcurr=None, Rprev_inv=None, tprev=None, which=0):
        if Rcurr is None:
            _chk(lib().ef_icp_step_async(self.h_ctx, which, level, None, None, None, None))
            return
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        Rc, tc, Rp, tp = f32(Rcurr), f32(tcurr), f32(Rprev_inv), f32(tprev)
        _chk(lib().ef_icp_step_async(self.h_ctx, which, level, _p(Rc), _p(tc), _p(Rp), _p(tp)))

    def icp_dense_pass_async(self, level, which=0):
        _chk(lib().ef_icp_dense_pass_async(self.h_ctx, which, level))

    def rgb_residual(self, level, krkinv, kt, which=0):
        kk = np.ascontiguousarray(krkinv, np.float32)
        k3 = np.ascontiguousarray(kt, np.float32)
        sigma, count = C.c_int32(), C.c_int32()
        _chk(lib().ef_rgb_residual(self.h_ctx, which, level, _p(kk), _p(k3), C.byref(sigma), C.byref(count)))
        return sigma.value, count.value

    def rgb_step(self, level, sigma, which=0):
        A, b = np.zeros((6, 6), np.float32), np.zeros(6, np.float32)
        _chk(lib().ef_rgb_step(self.h_ctx, which, level, _f(sigma), _p(A), _p(b)))
        return A, b

    def so3_step(self, image_basis, kinv, krlr, which=0):
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        A, b, res = np.zeros((3, 3), np.float32), np.zeros(3, np.float32), np.zeros(2, np.float32)
        ib, ki, kr = f32(image_basis), f32(kinv), f32(krlr)
        _chk(lib().ef_so3_step(self.h_ctx, which, _p(ib), _p(ki), _p(kr), _p(A), _p(b), _p(res)))
        return A, b, res

    # ---- preprocess + map stages
    def preprocess_depth(self, raw_ptr, cutoff, filtered_ptr, metric_ptr, metric_filtered_ptr):
        _chk(lib().ef_preprocess_depth(self.h_ctx, C.c_void_p(raw_ptr), _f(cutoff), C.c_void_p(filtered_ptr),
                                       C.c_void_p(metric_ptr), C.c_void_p(metric_filtered_ptr)))

    def map_initialise(self):
        _chk(lib().ef_map_initialise(self.h_ctx))

    def map_predict_indices(self, T_wc, time, max_depth, time_delta):
        _chk(lib().ef_map_predict_indices(self.h_ctx, _p(_T(T_wc)), int(time), _f(max_depth), int(time_delta)))

    def map_fuse(self, T_wc, time, max_depth, weighting):
        _chk(lib().ef_map_fuse(self.h_ctx, _p(_T(T_wc)), int(time), _f(max_depth), _f(weighting)))

    def map_clean(self, T_wc, time, conf_threshold, time_delta, max_depth):
        _chk(lib().ef_map_clean(self.h_ctx, _p(_T(T_wc)), int(time), _f(conf_threshold), int(time_delta), _f(max_depth)))

    def map_clean_deform(self, T_wc, time, conf_threshold, time_delta, max_depth, nodes, is_fern=False):
        nd = np.ascontiguousarray(nodes, np.float32).reshape(-1, 16)
        _chk(lib().ef_map_clean_deform(self.h_ctx, _p(_T(T_wc)), int(time), _f(conf_threshold), int(time_delta), _f(max_depth), _p(nd), len(nd),
                                       int(is_fern)))

    def map_raycast(self, T_wc, max_depth, conf_threshold, time, max_time, time_delta, mode=0):
        _chk(lib().ef_map_raycast(self.h_ctx, _p(_T(T_wc)), _f(max_depth), _f(conf_threshold), int(time), int(max_time),
                                  int(time_delta), int(mode)))

    def map_fill_in(self, passthrough_geometry=False, passthrough_image=False):
        _chk(lib().ef_map_fill_in(self.h_ctx, int(passthrough_geometry), int(passthrough_image)))

    def dense_enough(self):
        out = C.c_int32()
        _chk(lib().ef_dense_enough(self.h_ctx, C.byref(out)))
        return bool(out.value)

    def map_count(self):
        n = C.c_int32()
        _chk(lib().ef_map_count(self.h_ctx, C.byref(n)))
        return n.value

    def map_download(self):
        n = self.map_count()
        out = np.zeros((max(n, 1), 12), np.float32)
        cnt = C.c_int32()
        _chk(lib().ef_map_download(self.h_ctx, _p(out), n, C.byref(cnt)))
        return out[:n]

    def map_download_new(self):
        out = np.zeros((self.w * self.h, 12), np.float32)
        cnt = C.c_int32()
        _chk(lib().ef_map_download_new(self.h_ctx, _p(out), self.w * self.h, C.byref(cnt)))
        return out[:cnt.value].copy()

    def map_upload_range(self, surfels, first):
        s = np.ascontiguousarray(surfels, np.float32)
        _chk(lib().ef_map_upload_range(self.h_ctx, _p(s), int(first), len(s)))

    def join_lookahead(self):
        _chk(lib().ef_join_lookahead(self.h_ctx))

    def stage_ms(self):
        """EF_STAGE_TIMING=1 (set before the context is created): ms per stage of the last frame, index as in the header."""
        out = (C.c_float * 16)()
        n = lib().ef_debug_stage_ms(self.h_ctx, out)
        return [out[i] for i in range(n)]

    def map_upload(self, surfels):
        s = np.ascontiguousarray(surfels, np.float32)
        _chk(lib().ef_map_upload(self.h_ctx, _p(s), len(s)))
