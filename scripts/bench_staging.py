"""Pageable-memory path of the host-buffer API (redux_ctx_set_staging): end-to-end MB/s of raw input for one encode +
decode of N x 64 KiB blocks held in plain numpy buffers, against the number of host copy threads, the piece size and
the ring depth; the driver-staged path and pinned buffers beside it.  Usage: python scripts/bench_staging.py [blocks]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import redux_b200 as rb

rb.process_init()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
L = 65536
SEED = 0x5EED202610180000
corp = None
try:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from corpora_fixture import text_corpus
    corp = text_corpus()
except Exception:
    pass
raw = rb.generate_blocks_host(0, n, L, SEED, corpus=corp) if corp is not None else rb.generate_blocks_host(0, n, L, SEED)
off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
model = rb.AdaptiveTreeModel(rb.Parameters(8, 14, 16))
comp = np.zeros(int(rb.compress_bound(L, 16)) * n // 2 + (64 << 20), dtype=np.uint8)
back = np.zeros_like(raw)
res = {"blocks": n, "raw_bytes": int(raw.size), "host_threads": os.cpu_count(), "rows": []}


def run(c, src, dst_comp, dst_back, reps=2):
    out, out_off, _ = c.encode_batch(src, off, model, out=dst_comp)
    c.decode_batch(dst_comp, out_off, off, model, raw=dst_back)
    t0 = time.perf_counter()
    te = td = 0.0
    for _ in range(reps):
        a = time.perf_counter()
        out, out_off, _ = c.encode_batch(src, off, model, out=dst_comp)
        b = time.perf_counter()
        c.decode_batch(dst_comp, out_off, off, model, raw=dst_back)
        d = time.perf_counter()
        te += b - a; td += d - b
    assert (dst_back == src).all()
    return te / reps, td / reps


with rb.Context([0]) as c:
    def row(name, **kw):
        te, td = run(c, raw, comp, back)
        r = dict(name=name, encode_ms=round(te * 1e3, 1), decode_ms=round(td * 1e3, 1),
                 MBps=round(raw.size / (te + td) / 1e6, 1), **kw)
        res["rows"].append(r)
        print(json.dumps(r), flush=True)

    c.set_staging(False)
    row("driver-staged pageable")
    for threads in (1, 2, 4, 6, 8, 12):
        c.set_staging(True, threads=threads)
        row("library-staged", threads=threads, piece_MiB=8, slots=4)
    c.set_staging(True, threads=6, piece_bytes=2 << 20, slots=8)
    row("library-staged", threads=6, piece_MiB=2, slots=8)
    c.set_staging(True, threads=6, piece_bytes=16 << 20, slots=4)
    row("library-staged", threads=6, piece_MiB=16, slots=4)
    c.set_staging(True, threads=6, piece_bytes=32 << 20, slots=3)
    row("library-staged", threads=6, piece_MiB=32, slots=3)
    hb = [rb.HostBuffer(x.size) for x in (raw, comp, back)]
    hb[0].array[:] = raw
    te, td = run(c, hb[0].array, hb[1].array, hb[2].array)
    r = dict(name="pinned", encode_ms=round(te * 1e3, 1), decode_ms=round(td * 1e3, 1), MBps=round(raw.size / (te + td) / 1e6, 1))
    res["rows"].append(r); print(json.dumps(r))
    # host memcpy speed of this box, for scale
    t0 = time.perf_counter(); back[:] = raw; t1 = time.perf_counter()
    res["numpy_copy_GBps_one_thread"] = round(raw.size / (t1 - t0) / 1e9, 2)
    for b in hb: b.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_staging.json", "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "rows"}))
