"""Phase timestamps of k_iter1 / k_iter2 at level 0 (library built with -DEF_PROFILE_PHASES)."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
from elasticfusion_b200 import synth, capi
K = synth.K_DEFAULT
frames = list(synth.sequence(6, K, seed=42, noise=True))
BIG = 2147483647 // 2
ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=1000000, time_delta=BIG))
for i, (rgb, d, _) in enumerate(frames):
    ctx.process_frame(rgb, d, i)
# GNState layout: find dbg by scanning for it via offset: use a helper export? read whole struct
import ctypes
lib = capi.lib()
# read the raw GNState: ef_buffer has no id for it; use cudart directly
cudart = C.CDLL('libcudart.so')
ptr = C.c_void_p(); n = C.c_size_t()
# trick: the gn pointer is not exposed; add a debug export
lib.ef_debug_gn.restype = C.c_void_p
gnp = lib.ef_debug_gn(ctx.h_ctx, 0)
sz = lib.ef_debug_gn_size()
buf = (C.c_char * sz)()
cudart.cudaMemcpy(buf, C.c_void_p(gnp), C.c_size_t(sz), 2)
off = lib.ef_debug_dbg_offset()
dbg = np.frombuffer(buf, dtype=np.int64, count=32, offset=off)
print("k_iter1 cycles: start->after cand %d, icp loop %d, block reduce %d" % (dbg[1]-dbg[0], dbg[2]-dbg[1], dbg[3]-dbg[2]))
print("k_iter2 block0: stats %d, rgb+presum %d" % (dbg[9]-dbg[8], dbg[10]-dbg[9]))
print("k_iter2 last block: final sums %d, stage+lastA %d, ldlt %d, rodrigues %d, rest %d" % (dbg[12]-dbg[11], dbg[13]-dbg[12], dbg[14]-dbg[13], dbg[15]-dbg[14], dbg[16]-dbg[15]))
