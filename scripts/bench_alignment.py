"""Does the lane decoder / encoder care how the streams of a warp are ALIGNED?  32 streams of 256 KiB, device resident,
(a) every raw slot and every compressed stream at a multiple of 16 bytes, (b) raw slots at byte phases 0..3 (pads of
0..3 bytes between them), (c) compressed streams at byte phases 0..3.  A large gap means lanes of a warp stay diverged
after their (different) head steps."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import redux_b200 as rb

SEED = 0x5EED202610180000
n, L = 32, 262144


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps): fn()
    ev[1].record(); ev[1].synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def main():
    ctx = rb.Context([0]); ctx.set_schedule(rb.SCHED_LANE)
    stream = torch.cuda.current_stream().cuda_stream
    res = {}
    for params in ((8, 14, 16), (8, 30, 32)):
        model = rb.AdaptiveTreeModel(rb.Parameters(*params))
        host = rb.generate_blocks_host(1, n * 4, L, SEED)[: 0]  # noqa
        blocks = [rb.generate_blocks_host(1 + 4 * i, 1, L, SEED) for i in range(n)]      # text class
        for name, raw_pad, comp_pad in (("aligned", 0, 0), ("raw slots at byte phases 0..3", 1, 0), ("streams at byte phases 0..3", 0, 1)):
            # raw layout
            roff = np.zeros(n + 1, dtype=np.int64)
            for i in range(n):
                roff[i + 1] = roff[i] + L + (((i % 4) + 1) % 4 if raw_pad else 0) + (16 if not raw_pad else 0)
            raw = torch.zeros(int(roff[-1]) + 64, dtype=torch.uint8, device="cuda")
            in_off = np.zeros(n + 1, dtype=np.int64)
            for i in range(n):
                raw[int(roff[i]):int(roff[i]) + L] = torch.from_numpy(blocks[i]).cuda()
            # encode needs back-to-back in_offsets: use a compact copy for the encoder input
            cin = torch.from_numpy(np.concatenate(blocks)).cuda()
            cin_off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
            cap = n * L * 2
            comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); coff = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
            st = torch.zeros(n, dtype=torch.int32, device="cuda")
            te = timed(lambda: ctx.encode_batch_device(cin, cin_off, n, L, comp, cap, coff, st, model, device=0, stream=stream))
            torch.cuda.synchronize()
            hc, hoff = comp.cpu().numpy(), coff.cpu().numpy()
            # compressed layout
            streams = [hc[int(hoff[i]):int(hoff[i + 1])] for i in range(n)]
            c2off = np.zeros(n + 1, dtype=np.int64)
            parts = []
            for i in range(n):
                pad = (-(int(c2off[i]) + len(streams[i])) % 16) if not comp_pad else 0
                if comp_pad:
                    pad = (4 - (int(c2off[i]) + len(streams[i])) % 4) % 4 + (i + 1) % 4      # next stream starts at phase (i+1) % 4
                parts += [streams[i], np.zeros(pad, dtype=np.uint8)]
                c2off[i + 1] = c2off[i] + len(streams[i]) + pad
            comp2 = torch.from_numpy(np.concatenate(parts + [np.zeros(64, np.uint8)])).cuda()
            # decode job wants comp offsets [start_i, start_{i+1}) = stream + pad: pads are trailing garbage, never read
            d_coff = torch.from_numpy(c2off).cuda(); d_roff = torch.from_numpy(roff).cuda()
            back = torch.zeros_like(raw)
            rl = torch.zeros(n, dtype=torch.int64, device="cuda"); cons = torch.zeros(n, dtype=torch.int64, device="cuda")
            td = timed(lambda: ctx.decode_batch_device(comp2, d_coff, n, L + 32, back, d_roff, rl, cons, st, model, device=0, stream=stream))
            torch.cuda.synchronize()
            # the slots are larger than the blocks, so the decoder stops at the EOF symbol
            ok = all(bool(torch.equal(back[int(roff[i]):int(roff[i]) + L], raw[int(roff[i]):int(roff[i]) + L])) for i in range(n))
            assert ok and int(st.abs().max()) == 0 and bool((rl == L).all())
            res["%s %s" % (params, name)] = {"encode_ms": round(te, 3), "decode_ms": round(td, 3)}
            print(params, name, res["%s %s" % (params, name)], flush=True)
        # encoder: raw blocks at byte phases 0..3 of the input buffer (lengths L - (i % 4), back to back)
        lens = np.array([L - (i % 4) for i in range(n)], dtype=np.int64)
        ioff = np.zeros(n + 1, dtype=np.int64); np.cumsum(lens, out=ioff[1:])
        cin = torch.from_numpy(np.concatenate([blocks[i][:lens[i]] for i in range(n)])).cuda()
        d_ioff = torch.from_numpy(ioff).cuda()
        cap = n * L * 2
        comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); coff = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
        st = torch.zeros(n, dtype=torch.int32, device="cuda")
        te = timed(lambda: ctx.encode_batch_device(cin, d_ioff, n, L, comp, cap, coff, st, model, device=0, stream=stream))
        res["%s encoder input at byte phases 0..3" % (params,)] = {"encode_ms": round(te, 3)}
        print(params, "encoder input at byte phases 0..3", round(te, 3), flush=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_alignment.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
