"""One 768,771 B stream decoded by the warp mapping (for an ncu capture of decode_warp_al_kernel)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import redux_b200 as rb
rb.lib()
params = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8,14,16").split(","))
ctx = rb.Context([0]); ctx.set_schedule(rb.SCHED_WARP)
n = 768771
data = rb.generate_blocks_host(1, 1, n, 0x5EED202610180000)
off = np.array([0, n], dtype=np.uint64)
model = rb.AdaptiveTreeModel(rb.Parameters(*params))
comp, coff, st = ctx.encode_batch(data, off, model)
ctx.timing_enable(True)
for _ in range(2):
    t0 = time.perf_counter(); back, rl, cons, st = ctx.decode_batch(comp, coff, off, model); t1 = time.perf_counter()
    print("decode_batch %.2f ms" % ((t1 - t0) * 1e3), ctx.timing_collect()["decode"])
assert (back[:n] == data).all()
