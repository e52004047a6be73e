"""Small mixed workload for compute-sanitizer (memcheck): both mappings, ragged/unaligned/empty blocks,
truncated streams, tight capacities."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import redux_b200 as rb
rng = np.random.default_rng(3)
base = rb.generate_blocks_host(0, 64, 3000, 0x5EED202610180000)
lens = [0, 1, 2, 3, 5, 17, 100, 2999] + [int(x) for x in rng.integers(0, 3000, size=56)]
blocks = [base[i * 3000:i * 3000 + lens[i]] for i in range(64)]
data = np.concatenate(blocks); off = np.zeros(65, dtype=np.uint64); np.cumsum(lens, out=off[1:])
for sched in (rb.SCHED_LANE, rb.SCHED_WARP):
    for params in ((8, 10, 12), (8, 14, 16), (8, 30, 32), (8, 20, 44)):
        with rb.Context([0]) as ctx:
            ctx.set_schedule(sched)
            m = rb.AdaptiveTreeModel(rb.Parameters(*params))
            comp, coff, st = ctx.encode_batch(data, off, m)
            back, ln, cons, st = ctx.decode_batch(comp, coff, off, m)
            assert (back == data).all() and (ln == np.array(lens)).all()
            # truncated streams + exact-capacity slots
            cut = coff.copy(); 
            comp2 = np.concatenate([comp[int(coff[i]):int(coff[i + 1]) - (1 if i % 3 == 0 and coff[i + 1] > coff[i] else 0)] for i in range(64)])
            l2 = [int(coff[i + 1] - coff[i]) - (1 if i % 3 == 0 and coff[i + 1] > coff[i] else 0) for i in range(64)]
            o2 = np.zeros(65, dtype=np.uint64); np.cumsum(l2, out=o2[1:])
            ctx.decode_batch(comp2, o2, off, m, check=False)
print("sanitize workload ok")
