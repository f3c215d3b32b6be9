"""Developer tool: phase timeline of one CTA of the attention forward kernel (variants PIANOBART_B200_ATTN_FWD=3 / 2).
Needs the trace build: tools/build_trace_lib.sh, PIANOBART_B200_LIB=pianobart_b200/libpianobart_b200_trace.so."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
src = open(os.path.join(os.path.dirname(__file__), 'gpu_attn_prof.py')).read()
src = src.replace('lib.pb_attn_bwd(C.byref(a), L.stream_ptr())', 'None')
exec(src)
out = np.zeros(3 * 64 * 8, dtype=np.int64)
assert lib.pb_debug_trace(out.ctypes.data_as(C.c_void_p), out.size) == 0
t = out.reshape(3, 64, 8)
t0 = t[0, 63, 0]
nb = 8
print('MMA thread (two-tile kernel): [loop top, P_a + V ok, PV_a + S_a(j+1) issued, P_b ok, PV_b + S_b(j+1) issued]   (cycles since CTA entry)')
for j in range(nb):
    print(j, [int(x - t0) if x > 0 else -1 for x in t[0, j, :5]])
for role in (1, 2):  # (fwd4: warps 2 and 10)
    print('softmax warp %d: [loop top, s_full ok, pass 1 done, rescale done, pass 2 done, arrived]' % role)
    for j in range(nb):
        print(j, [int(x - t0) if x > 0 else -1 for x in t[role, j, :6]])
print('CTA life (thread 64): [entry, prologue done, last PV done, O stored]', [int(x - t0) if x > 0 else -1 for x in t[0, 63, :4]])
