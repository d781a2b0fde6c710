"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_ref/gl/libef_refgl.so: the REFERENCE's GLSL shader files executed
unmodified on Mesa llvmpipe (oracle/gl/ref_gl_harness.cpp). CPU only; needs the Mesa libGL that ships with Nsight Compute in this
image and /root/reference/Core/Shaders. The interpreter must be started with LD_LIBRARY_PATH containing oracle/_ref/gl (the libX11
stand-in) and the Mesa directory: run(), below, re-executes a function in such a child process."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GL_DIR = os.path.join(_HERE, "_ref", "gl")
SO = os.path.join(GL_DIR, "libef_refgl.so")
SHADERS = "/root/reference/Core/Shaders"
_LIB = None


def mesa_dir():
    c = sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/linux-desktop-*/Mesa"))
    return c[-1] if c else None


def available() -> bool:
    return os.path.exists(SO) and os.path.exists(os.path.join(GL_DIR, "libX11.so.6")) and mesa_dir() is not None and os.path.isdir(SHADERS)


def env():
    e = dict(os.environ)
    e["LD_LIBRARY_PATH"] = GL_DIR + ":" + mesa_dir() + ":" + e.get("LD_LIBRARY_PATH", "")
    return e


def in_gl_process() -> bool:
    return GL_DIR in os.environ.get("LD_LIBRARY_PATH", "")


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _f(x):
    return C.c_float(float(x))


class RefGL:
    """One GL context at a fixed resolution / intrinsics. Stage functions take and return the same arrays as oracle.ef_oracle's."""

    def __init__(self, K, tex_dim=1024):
        global _LIB
        assert in_gl_process(), "start the interpreter through ef_refgl.env()"
        if _LIB is None:
            _LIB = C.CDLL(SO)
            _LIB.efg_log.restype = C.c_char_p
        self.K = K
        rc = _LIB.efg_init((mesa_dir() + "/libGL.so.1").encode(), SHADERS.encode(), K.width, K.height, _f(K.fx), _f(K.fy), _f(K.cx), _f(K.cy), tex_dim)
        if rc:
            raise RuntimeError("efg_init failed:\n" + _LIB.efg_log().decode())
        self.lib = _LIB

    def log(self):
        return self.lib.efg_log().decode()

    def bilateral(self, depth, max_d):
        out = np.zeros_like(depth)
        self.lib.efg_bilateral(_p(np.ascontiguousarray(depth, np.uint16)), _f(max_d), _p(out))
        return out

    def metric(self, depth, max_d):
        out = np.zeros(depth.shape, np.float32)
        self.lib.efg_metric(_p(np.ascontiguousarray(depth, np.uint16)), _f(max_d), _p(out))
        return out

    def feedback_buffer(self, rgb, depth_metric, time, max_depth):
        out = np.zeros((self.K.width * self.K.height, 12), np.float32)
        n = self.lib.efg_feedback(_p(np.ascontiguousarray(rgb, np.uint8)), _p(np.ascontiguousarray(depth_metric, np.float32)), int(time), _f(max_depth), _p(out))
        return out[:n].copy()

    def map_initialise(self, raw_fb, filt_fb):
        out = np.zeros((self.K.width * self.K.height, 12), np.float32)
        r, f = np.ascontiguousarray(raw_fb, np.float32), np.ascontiguousarray(filt_fb, np.float32)
        n = self.lib.efg_initialise(_p(r), len(r), _p(f), len(f), _p(out))
        return out[:n].copy()

    def predict_indices(self, surfels, T_wc, time, max_depth, time_delta):
        r, c = self.K.height, self.K.width
        index = np.zeros((r, c), np.uint32)
        vc, ct, nr = (np.zeros((r, c, 4), np.float32) for _ in range(3))
        s = np.ascontiguousarray(surfels, np.float32)
        T = np.ascontiguousarray(T_wc, np.float64)
        self.lib.efg_predict_indices(_p(s), len(s), _p(T), int(time), _f(max_depth), int(time_delta), _p(index), _p(vc), _p(ct), _p(nr))
        return index, vc, ct, nr

    def fuse(self, surfels, T_wc, time, rgb, depth_raw, depth_filt, index, vc, ct, nr, max_depth, weighting):
        m = np.ascontiguousarray(surfels, np.float32).copy()
        new = np.zeros((self.K.width * self.K.height, 12), np.float32)
        T = np.ascontiguousarray(T_wc, np.float64)
        n = self.lib.efg_fuse(_p(m), len(m), _p(T), int(time), _p(np.ascontiguousarray(rgb, np.uint8)), _p(np.ascontiguousarray(depth_raw, np.float32)),
                              _p(np.ascontiguousarray(depth_filt, np.float32)), _p(index), _p(vc), _p(ct), _p(nr), _f(max_depth), _f(weighting), _p(new))
        return m, new[:n].copy()

    def clean(self, surfels, new_unstable, T_wc, time, index, vc, ct, nr, conf_threshold, time_delta, max_depth, nodes=None, depth=None, is_fern=False):
        s = np.ascontiguousarray(surfels, np.float32)
        nu = np.ascontiguousarray(new_unstable, np.float32).reshape(-1, 12)
        out = np.zeros((len(s) + len(nu) + 1, 12), np.float32)
        T = np.ascontiguousarray(T_wc, np.float64)
        nd = None if nodes is None else np.ascontiguousarray(nodes, np.float32).reshape(-1, 16)
        d = None if depth is None else np.ascontiguousarray(depth, np.float32)
        n = self.lib.efg_clean(_p(s), len(s), _p(nu), len(nu), _p(T), int(time), _p(index), _p(vc), _p(ct), _p(nr), _f(conf_threshold), int(time_delta),
                               _f(max_depth), _p(nd), 0 if nd is None else len(nd), _p(d), int(is_fern), _p(out))
        return out[:n].copy()

    def combined_predict(self, surfels, T_wc, max_depth, conf_threshold, time, max_time, time_delta, depth_only=False):
        r, c = self.K.height, self.K.width
        s = np.ascontiguousarray(surfels, np.float32)
        T = np.ascontiguousarray(T_wc, np.float64)
        if depth_only:
            d = np.zeros((r, c), np.float32)
            self.lib.efg_combined_predict(_p(s), len(s), _p(T), _f(max_depth), _f(conf_threshold), int(time), int(max_time), int(time_delta), None, None,
                                          None, None, _p(d), 1)
            return d
        image = np.zeros((r, c, 4), np.uint8)
        vertex, normal = np.zeros((r, c, 4), np.float32), np.zeros((r, c, 4), np.float32)
        tm = np.zeros((r, c), np.uint16)
        self.lib.efg_combined_predict(_p(s), len(s), _p(T), _f(max_depth), _f(conf_threshold), int(time), int(max_time), int(time_delta), _p(image),
                                      _p(vertex), _p(normal), _p(tm), None, 0)
        return image, vertex, normal, tm

    def fill_vertex(self, existing4, raw_depth, passthrough):
        out = np.zeros_like(existing4)
        self.lib.efg_fill_vertex(_p(np.ascontiguousarray(existing4, np.float32)), _p(np.ascontiguousarray(raw_depth, np.uint16)), int(passthrough), _p(out))
        return out

    def fill_normal(self, existing4, raw_depth, passthrough):
        out = np.zeros_like(existing4)
        self.lib.efg_fill_normal(_p(np.ascontiguousarray(existing4, np.float32)), _p(np.ascontiguousarray(raw_depth, np.uint16)), int(passthrough), _p(out))
        return out

    def fill_image(self, existing4, rgb, passthrough):
        out = np.zeros_like(existing4)
        self.lib.efg_fill_image(_p(np.ascontiguousarray(existing4, np.uint8)), _p(np.ascontiguousarray(rgb, np.uint8)), int(passthrough), _p(out))
        return out


def run_script(path, *args, timeout=1800):
    """Runs a Python script in a child interpreter whose dynamic loader sees the X11 stand-in and Mesa."""
    return subprocess.run([sys.executable, path, *args], env=env(), timeout=timeout, check=True)
