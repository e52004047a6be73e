"""One 768,771 B stream through the split encoder (for an ncu launch list: which phase takes the time)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import redux_b200 as rb
rb.lib()
params = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8,14,16").split(","))
ctx = rb.Context([0]); ctx.set_schedule(rb.SCHED_SPLIT)
n = 768771
data = rb.generate_blocks_host(1, 1, n, 0x5EED202610180000)
off = np.array([0, n], dtype=np.uint64)
model = rb.AdaptiveTreeModel(rb.Parameters(*params))
ctx.timing_enable(True)
for _ in range(3):
    t0 = time.perf_counter(); ctx.encode_batch(data, off, model); t1 = time.perf_counter()
    print("encode_batch %.2f ms" % ((t1 - t0) * 1e3), ctx.timing_collect())
