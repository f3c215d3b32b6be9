"""Developer tool: phase timeline of one CTA of the attention dQ kernel (library built with -DPB_TRACE)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
exec(open(os.path.join(os.path.dirname(__file__), 'gpu_attn_prof.py')).read())
out = np.zeros(3 * 64 * 8, dtype=np.int64)
assert lib.pb_debug_trace(out.ctypes.data_as(C.c_void_p), out.size) == 0
t = out.reshape(3, 64, 8)
t0 = t[t > 0].min()
nb = 16
print('MMA warp: [kv_full(j+1) ok, sdp(j+1) issued, ds_full(j) ok, dQ(j) issued]')
for j in range(nb):
    print(j, [int(x - t0) if x > 0 else -1 for x in t[0, j, :4]])
for role in (1, 2):
    print('softmax warp %d: [bar, sdp_full ok, tmem ld done, math done, dq_done ok, arrived]' % role)
    for j in range(nb):
        print(j, [int(x - t0) if x > 0 else -1 for x in t[role, j, :6]])
