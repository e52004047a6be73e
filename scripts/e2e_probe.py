"""Times the pieces of the end-to-end path separately (pinned copies, host-API encode / decode)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import redux_b200 as rb
rb.lib()   # before CUDA initialises: the library's load hook asks for 32 hardware queues
n, L = 65536, 65536
ctx = rb.Context([0])
model = rb.AdaptiveTreeModel(rb.Parameters(8, 14, 16))
raw = torch.empty(n * L, dtype=torch.uint8, device="cuda")
ctx.generate_blocks_device(raw, 0, n, L, 0x5EED202610180000, device=0, stream=None)
torch.cuda.synchronize()
h_raw = torch.empty(n * L, dtype=torch.uint8, pin_memory=True); h_raw.copy_(raw)
cap = n * L + n * 4200
h_comp = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
h_back = torch.empty(n * L, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
def t(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("H2D 4GiB pinned   %.1f ms" % t(lambda: raw.copy_(h_raw, non_blocking=True)))
print("D2H 4GiB pinned   %.1f ms" % t(lambda: h_back.copy_(raw, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): raw.copy_(h_raw, non_blocking=True)
    with torch.cuda.stream(s2): h_back.copy_(raw, non_blocking=True)
print("H2D+D2H duplex    %.1f ms" % t(both))
np_raw, np_comp, np_back = h_raw.numpy(), h_comp.numpy(), h_back.numpy()
np_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
del raw; torch.cuda.empty_cache()
res = {}
def enc():
    res["o"] = ctx.encode_batch(np_raw, np_off, model, out=np_comp)
print("encode_batch host %.1f ms" % t(enc))
out, out_off, st = res["o"]
def dec():
    ctx.decode_batch(np_comp, out_off, np_off, model, raw=np_back)
print("decode_batch host %.1f ms" % t(dec))
assert (np_back == np_raw).all()
print("comp bytes", int(out_off[-1]))
