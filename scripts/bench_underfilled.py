"""Which stream-to-thread mapping for batches that do not fill the GPU?  

At 8 GPUs the headline batch leaves 8,192 blocks per GPU = 1.7 warps per SM under the lane mapping.  This measures,
device resident, encode and decode of n = 512 ... 16,384 blocks of 64 KiB under the lane mapping (32 streams per warp,
table in shared memory), the warp mapping (one stream per warp, cumulative array in registers) and the split encoder,
for the narrow and the wide class, and prints what REDUX_SCHED_AUTO picks.  Output: gpurun_out/bench_underfilled.json"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import redux_b200 as rb

SEED = 0x5EED202610180000
L = 65536


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps): fn()
    ev[1].record(); ev[1].synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def main():
    ctx = rb.Context([0])
    stream = torch.cuda.current_stream().cuda_stream
    res = {}
    for params in ((8, 14, 16), (8, 30, 32)):
        model = rb.AdaptiveTreeModel(rb.Parameters(*params))
        for n in (512, 2048, 4096, 8192, 16384):
            raw = torch.empty(n * L, dtype=torch.uint8, device="cuda")
            ctx.generate_blocks_device(raw, 0, n, L, SEED, device=0, stream=stream)
            off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
            cap = n * L + n * (L // 16) + 65536
            comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); coff = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
            st = torch.zeros(n, dtype=torch.int32, device="cuda"); back = torch.empty(n * L, dtype=torch.uint8, device="cuda")
            rl = torch.zeros(n, dtype=torch.int64, device="cuda"); cons = torch.zeros(n, dtype=torch.int64, device="cuda")
            row = {}
            ref = None
            for sched, label in ((rb.SCHED_LANE, "lane"), (rb.SCHED_WARP, "warp"), (rb.SCHED_SPLIT, "split"), (rb.SCHED_AUTO, "auto")):
                if label == "split" and n * L * 8 > (2 << 30):
                    continue                                   # the split encoder's workspace bound
                ctx.set_schedule(sched)
                te = timed(lambda: ctx.encode_batch_device(raw, off, n, L, comp, cap, coff, st, model, device=0, stream=stream))
                td = timed(lambda: ctx.decode_batch_device(comp, coff, n, L, back, off, rl, cons, st, model, device=0, stream=stream))
                torch.cuda.synchronize()
                assert torch.equal(back, raw) and int(st.abs().max()) == 0
                sig = (int(coff[-1]), int(comp[:int(coff[-1])].to(torch.int64).sum()))
                ref = ref or sig
                assert sig == ref, "mappings disagree"
                row[label] = {"encode_ms": round(te, 3), "decode_ms": round(td, 3),
                              "round_trip_GBps": round(n * L / (te + td) / 1e6, 2)}
            res["%s n=%d" % (params, n)] = row
            print(params, n, json.dumps(row), flush=True)
            del raw, comp, back
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_underfilled.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
