"""Pre-trained byte models on the tuned lane kernels: device-resident encode + decode of 65,536 x 64 KiB blocks
starting from a model trained on 4,000 symbols, next to the fresh model (kernel times from CUDA events)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import redux_b200 as rb
import oracle_lib as o
rb.lib()
n, L = 65536, 65536
SEED = 0x5EED202610180000
ctx = rb.Context([0])
raw = torch.empty(n * L, dtype=torch.uint8, device="cuda")
ctx.generate_blocks_device(raw, 0, n, L, SEED, device=0, stream=None)
torch.cuda.synchronize()
# the host-buffer API is the one that takes a trained model; time its kernels on a 4,096-block slice per chunk
host = raw[: 4096 * L].cpu().numpy()
off = np.arange(4097, dtype=np.uint64) * np.uint64(L)
train = [int(x) for x in host[:4000]]
res = {}
for name, params, tr in (("fresh (8,14,16)", (8, 14, 16), None), ("trained (8,14,16)", (8, 14, 16), train),
                         ("fresh (8,30,32)", (8, 30, 32), None), ("trained (8,30,32)", (8, 30, 32), train)):
    model = rb.AdaptiveTreeModel(rb.Parameters(*params))
    if tr: model.train(tr)
    os.environ["REDUX_PIPE_CHUNKS"] = "1"
    comp, coff, st = ctx.encode_batch(host, off, model)
    ctx.timing_enable(True); ctx.timing_collect()
    comp, coff, st = ctx.encode_batch(host, off, model)
    back, lens, cons, st2 = ctx.decode_batch(comp, coff, off, model)
    t = ctx.timing_collect(); ctx.timing_enable(False)
    assert (st == 0).all() and (st2 == 0).all() and (back[: host.size] == host).all()
    i = 4095
    want = o.compress_trained(host[i * L:(i + 1) * L], tr, o.TREE, params)[1] if tr else o.compress(host[i * L:(i + 1) * L], o.TREE, params)[1]
    assert comp[int(coff[i]):int(coff[i + 1])].tobytes() == want
    res[name] = {"encode_ms_sum_of_chunks": round(t["encode"][0], 2), "decode_ms_sum_of_chunks": round(t["decode"][0], 2),
                 "chunks": t["encode"][1], "ratio": round(host.size / int(coff[-1]), 3)}
    print(name, json.dumps(res[name]), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_trained.json"), "w"), indent=1)
