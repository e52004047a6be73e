"""Latency-bound configurations (BASELINE.json configs 1, 2, 5): few long streams, on the reference's own corpora
(tests/golden/corpora/corpora.tar.xz, byte-identical to /root/reference/resources): calgary/book1 as one stream,
the 29 Calgary + Canterbury files one stream each, and resources/large (bible.txt, world192.txt, seeded E.coli
stand-in) in 1 MiB blocks with the frequency_bits / code_bits sweep.  Reports MB/s of raw input for the three
mappings (lane, warp, split encoder) and for the CPU oracle; asserts every mapping's bytes equal the oracle's."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import redux_b200 as rb
import oracle_lib as o

import corpora_fixture as cf

FILES = cf.corpora()
MIB = 1 << 20
CONFIGS = {
    "config1_book1_one_stream": [FILES["calgary/book1"]],
    "config2_calgary_canterbury_29_streams": [FILES[n] for n in sorted(FILES) if n.startswith(("calgary/", "canterbury/"))],
    "config5_large_in_1MiB_blocks": cf.blocks_of(FILES["large/bible.txt"], MIB) + cf.blocks_of(FILES["large/world192.txt"], MIB)
                                    + cf.blocks_of(cf.ecoli_stand_in(), MIB),
}

def run(ctx, sched, data, off, model, reps=3):
    ctx.set_schedule(sched)
    comp, coff, st = ctx.encode_batch(data, off, model)
    t0 = time.perf_counter()
    for _ in range(reps): comp, coff, st = ctx.encode_batch(data, off, model)
    te = (time.perf_counter() - t0) / reps
    back, lens, cons, st = ctx.decode_batch(comp, coff, off, model)
    t0 = time.perf_counter()
    for _ in range(reps): back, lens, cons, st = ctx.decode_batch(comp, coff, off, model)
    td = (time.perf_counter() - t0) / reps
    assert (back == data).all()
    return te, td, comp, coff

SWEEP = [(8, 10, 16), (8, 14, 16), (8, 16, 18), (8, 20, 22), (8, 22, 24), (8, 24, 30), (8, 30, 32)]   # config 5


def main():
    one = rb.Context([0])
    ng = torch.cuda.device_count()
    every = rb.Context(list(range(ng))) if ng > 1 else one      # config 5: the block list sharded over all GPUs of the box
    res = {}
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, blocks_b in CONFIGS.items():
        if only and not name.startswith(only):
            continue
        ctx = every if name.startswith("config5") else one
        sizes = [len(b) for b in blocks_b]
        data = np.frombuffer(b"".join(blocks_b), dtype=np.uint8)
        off = np.zeros(len(sizes) + 1, dtype=np.uint64); np.cumsum(sizes, out=off[1:])
        for params in (SWEEP if name.startswith("config5") else ((8, 14, 16), (8, 22, 24), (8, 30, 32))):
            model = rb.AdaptiveTreeModel(rb.Parameters(*params))
            row = {"devices": ng if name.startswith("config5") else 1, "streams": len(sizes), "raw_bytes": int(data.size)}
            ref = None
            for sched, label in ((rb.SCHED_LANE, "lane"), (rb.SCHED_WARP, "warp"), (rb.SCHED_SPLIT, "split")):
                te, td, comp, coff = run(ctx, sched, data, off, model)
                row[label] = {"encode_MBps": round(data.size / te / 1e6, 2), "decode_MBps": round(data.size / td / 1e6, 2)}
                if ref is None: ref = (comp.tobytes(), coff.tobytes())
                else: assert ref == (comp.tobytes(), coff.tobytes()), "mappings disagree"
            threads = len(os.sched_getaffinity(0))
            t0 = time.perf_counter(); rc, slots, so, ol, st = o.compress_batch(data, off, o.TREE, params, threads); tc = time.perf_counter() - t0
            assert rc == 0
            want = b"".join(slots[int(so[i]):int(so[i]) + int(ol[i])].tobytes() for i in range(len(sizes)))
            assert want == ref[0], "GPU bytes differ from the oracle"
            raw_off = off
            comp_a = np.frombuffer(want, dtype=np.uint8)
            coff_a = np.zeros(len(sizes) + 1, dtype=np.uint64); np.cumsum(ol, out=coff_a[1:])
            t0 = time.perf_counter(); rc, back, rl, cons, st = o.decompress_batch(comp_a, coff_a, raw_off, o.TREE, params, threads); td_cpu = time.perf_counter() - t0
            assert rc == 0 and (back == data).all()
            row["cpu_oracle"] = {"encode_MBps": round(data.size / tc / 1e6, 2), "decode_MBps": round(data.size / td_cpu / 1e6, 2),
                                 "threads": threads}
            res["%s %s" % (name, params)] = row
            print(name, params, json.dumps(row), flush=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_small%s.json" % (("_" + only + "_%dgpu" % ng) if only else "")), "w"), indent=1)

if __name__ == "__main__":
    main()
