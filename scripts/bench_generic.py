"""Generic path (redux_generic_codec.cuh): symbol widths other than 8, pre-trained models with code_bits > 32.
8,192 mixed-entropy blocks of 16 KiB; kernel times from the library's own CUDA-event brackets."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import redux_b200 as rb
import oracle_lib as o
rb.lib()
n, L = 8192, 16384
raw = rb.generate_blocks_host(0, n, L, 0x5EED202610180000)
off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
ctx = rb.Context([0])
res = {}
cases = [("s=4 (4,10,16)", (4, 10, 16), None), ("s=12 (12,22,24)", (12, 22, 24), None), ("s=16 (16,18,20)", (16, 18, 20), None),
         ("s=8 pre-trained, code_bits 34: generic kernels (8,30,34)", (8, 30, 34), [int(x) for x in raw[:4000]])]
for name, params, train in cases:
    model = rb.AdaptiveTreeModel(rb.Parameters(*params))
    if train: model.train(train)
    comp, coff, st = ctx.encode_batch(raw, off, model)
    ctx.timing_enable(True); ctx.timing_collect()
    comp, coff, st = ctx.encode_batch(raw, off, model)
    back, lens, cons, st2 = ctx.decode_batch(comp, coff, off, model)
    t = ctx.timing_collect(); ctx.timing_enable(False)
    assert (st == 0).all() and (st2 == 0).all()
    i = 4097
    want = o.compress_trained(raw[i * L:(i + 1) * L], train, o.TREE, params)[1] if train else o.compress(raw[i * L:(i + 1) * L], o.TREE, params)[1]
    assert comp[int(coff[i]):int(coff[i + 1])].tobytes() == want
    if 8 % params[0] == 0: assert (back[: n * L] == raw).all()
    res[name] = {"encode_ms": round(t["encode"][0], 2), "decode_ms": round(t["decode"][0], 2),
                 "encode_MBps": round(n * L / t["encode"][0] / 1e3, 1), "decode_MBps": round(n * L / t["decode"][0] / 1e3, 1),
                 "ratio": round(n * L / int(coff[-1]), 3)}
    print(name, json.dumps(res[name]), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_generic.json"), "w"), indent=1)
