"""Latency-bound configurations (BASELINE.json configs 1, 2, 5 in spirit): few long streams.
The reference's corpora cannot travel to the GPU box, so the streams are text-like synthetic data of the
same sizes: one 768,771 B stream (calgary/book1's size), the 29 Calgary+Canterbury file sizes, and
12 blocks of <= 1 MiB.  Reports MB/s of raw input for the three mappings (lane, warp, split encoder) and for the CPU oracle."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import redux_b200 as rb
import oracle_lib as o

CORPUS = [111261, 768771, 610856, 102400, 377109, 21504, 246814, 53161, 82199, 46526, 13286, 11954, 38105,
          513216, 39611, 71646, 49379, 93695, 148481, 125179, 24603, 11150, 3721, 1029744, 419235, 471162,
          513216, 38240, 4227]
CONFIGS = {"config1_one_stream_768771B": [768771], "config2_29_corpus_sized_streams": CORPUS,
           "config5_12_blocks_1MiB": [1048576] * 9 + [901664, 311129, 444386]}

def text_like(n, seed):
    return rb.generate_blocks_host(1 + 4 * seed, 1, n, 0x5EED202610180000)       # class 1 = text-like

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
    ctx = rb.Context([0])
    res = {}
    for name, sizes in CONFIGS.items():
        blocks = [text_like(n, i) for i, n in enumerate(sizes)]
        data = np.concatenate(blocks)
        off = np.zeros(len(sizes) + 1, dtype=np.uint64); np.cumsum(sizes, out=off[1:])
        for params in (SWEEP if name.startswith("config5") else ((8, 14, 16), (8, 22, 24), (8, 30, 32))):
            model = rb.AdaptiveTreeModel(rb.Parameters(*params))
            row = {}
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
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_small.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
