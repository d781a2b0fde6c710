"""Phase timestamps (%globaltimer) of k_iter1 / k_iter2 at level 0. Needs an instrumented build of the library:
   EF_OUT=build/libefusion_prof.so ./build.sh -DEF_PROFILE_PHASES ; EF_LIB=build/libefusion_prof.so python scripts/phase_profile.py"""
import sys, ctypes as C, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
d = dbg.astype(np.int64)
print("globaltimer ns, last level-0 iteration (block 0 / last block stamps)")
print("k_iter1: start->cand done %d, icp loop %d, block reduce %d   [k1 total %d]" % (d[1]-d[0], d[2]-d[1], d[3]-d[2], d[3]-d[0]))
print("k1.end(block0) -> k2.start(block0): %d" % (d[8]-d[3]))
print("k_iter2 block0: stats %d, rgb+presum %d" % (d[9]-d[8], d[10]-d[9]))
print("k2 block0 done -> last block has ticket: %d" % (d[11]-d[10]))
print("k_iter2 last block: final sums %d, stage+lastA %d, ldlt %d, rodrigues %d, rest %d" % (d[12]-d[11], d[13]-d[12], d[14]-d[13], d[15]-d[14], d[16]-d[15]))
print("k1.start -> k2.end: %d" % (d[16]-d[0]))
