"""Developer tool: phase timeline of one CTA of the attention dQ kernel (library built with -DPB_TRACE)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
which = sys.argv[1] if len(sys.argv) > 1 else 'bwd'
src = open(os.path.join(os.path.dirname(__file__), 'gpu_attn_prof.py')).read()
if which == 'fwd':   # the trace buffer keeps the last kernel that wrote it: run the forward kernel only
    src = src.replace('lib.pb_attn_bwd(C.byref(a), L.stream_ptr())', 'None')
exec(src)
out = np.zeros(3 * 64 * 8, dtype=np.int64)
assert lib.pb_debug_trace(out.ctypes.data_as(C.c_void_p), out.size) == 0
t = out.reshape(3, 64, 8)
t0 = t[t > 0].min()
nb = 8
print('MMA warp: [k_full(j+1) ok, S(j+1) issued + dp_free(j) + v_full ok, ds_full(j) ok, dQ(j) issued]')
for j in range(nb):
    print(j, [int(x - t0) if x > 0 else -1 for x in t[0, j, :6]])
for role in (1, 2):
    print('softmax warp %d: [bar, s/dp_full ok, dP read (2nd chunk), chunk0 math done, dq_done ok, arrived]' % role)
    for j in range(nb):
        print(j, [int(x - t0) if x > 0 else -1 for x in t[role, j, :7]])
print('CTA life (thread 64): [entry, prologue done, last PV done, O stored, after dealloc]', [int(x - t0) if x > 0 else -1 for x in t[0, 63, :5]])
