"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a B200: `pytest -m gpu`.
Bit-exact is the bar: every compressed byte, every decoded byte, every count and status."""
import io
import json
import os

import numpy as np
import pytest

import oracle_lib as o
import redux_b200 as rb

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 0x5EED202610180000
TRIPLES = [(8, 14, 16), (8, 22, 24), (8, 30, 32)]          # tests/corpora.rs:35 (bits, bits+2)
KINDS = [(rb.AdaptiveLinearModel, o.LINEAR), (rb.AdaptiveTreeModel, o.TREE)]


@pytest.fixture(scope="module", params=["lane", "warp", "split"])
def ctx(request):
    """Every parity test runs on all stream-to-thread mappings: one stream per lane (Fenwick table in
    shared memory), one stream per warp (cumulative array in registers) and the split encoder (parallel
    model phase + one coder warp per stream; its decode side is the warp mapping)."""
    c = rb.Context()
    c.set_schedule({"lane": rb.SCHED_LANE, "warp": rb.SCHED_WARP, "split": rb.SCHED_SPLIT}[request.param])
    yield c
    c.close()


def concat(blocks):
    lens = np.array([len(b) for b in blocks], dtype=np.uint64)
    off = np.zeros(len(blocks) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    data = np.frombuffer(b"".join(blocks), dtype=np.uint8) if off[-1] else np.zeros(0, dtype=np.uint8)
    return data, off


def oracle_encode_all(blocks, kind, params):
    outs = []
    for b in blocks:
        rc, out, ic, oc = o.compress(b, kind, params)
        assert rc == o.OK and ic == len(b)
        outs.append(out)
    return outs


def check_encode(ctx, blocks, model_cls, okind, params, expect=None):
    model = model_cls(rb.Parameters(*params))
    data, off = concat(blocks)
    out, out_off, status = ctx.encode_batch(data, off, model)
    assert (status == 0).all(), status[status != 0][:8]
    expect = expect if expect is not None else oracle_encode_all(blocks, okind, params)
    for i, e in enumerate(expect):
        got = out[int(out_off[i]):int(out_off[i + 1])].tobytes()
        if got != e:
            n = min(len(got), len(e))
            first = next((j for j in range(n) if got[j] != e[j]), n)
            raise AssertionError("block %d (len %d) params %s: sizes %d vs oracle %d, first diff at byte %d: %s vs %s"
                                 % (i, len(blocks[i]), params, len(got), len(e), first,
                                    got[first:first + 8].hex(), e[first:first + 8].hex()))
    return out, out_off, expect


def check_decode(ctx, comp, comp_off, blocks, model_cls, params, slack=0):
    model = model_cls(rb.Parameters(*params))
    lens = np.array([len(b) + slack for b in blocks], dtype=np.uint64)
    raw_off = np.zeros(len(blocks) + 1, dtype=np.uint64)
    np.cumsum(lens, out=raw_off[1:])
    raw, raw_lens, consumed, status = ctx.decode_batch(comp, comp_off, raw_off, model)
    assert (status == 0).all(), status[status != 0][:8]
    for i, b in enumerate(blocks):
        assert int(raw_lens[i]) == len(b), (i, int(raw_lens[i]), len(b))
        assert int(consumed[i]) == int(comp_off[i + 1] - comp_off[i]), "decoder must consume the whole stream"
        got = raw[int(raw_off[i]):int(raw_off[i]) + len(b)].tobytes()
        if got != b:
            first = next(j for j in range(len(b)) if got[j] != b[j])
            raise AssertionError("decode block %d params %s: first diff at %d" % (i, params, first))


def test_kat_vectors_single_stream(ctx):
    """Golden vectors (SURVEY B.1 + the independent second reading of tests/golden/make_golden.py, incl. odd
    symbol widths and models trained before the call) through redux_compress / redux_decompress."""
    vecs = json.load(open(os.path.join(GOLD, "kat_vectors.json")))
    assert len(vecs) >= 130
    for v in vecs:
        p = tuple(v["params"])
        data, want = bytes.fromhex(v["input"]), bytes.fromhex(v["compressed"])
        train = list(bytes.fromhex(v["train"])) if "train" in v else None
        nbytes = (len(data) * 8 // p[0]) * p[0] // 8            # decompress() never flushes partial bytes
        for cls, _ in KINDS:
            model = cls(rb.Parameters(*p))
            if train is not None:
                model.train(train)
            out, (ic, oc) = ctx.compress(data, model)
            assert out == want, (v["name"], p, out.hex()[:32], want.hex()[:32])
            assert (ic, oc) == (len(data), len(want))
            model = cls(rb.Parameters(*p))
            if train is not None:
                model.train(train)
            dec, (ic2, oc2) = ctx.decompress(out, model, len(data) + 8)
            assert dec == data[:nbytes] and (ic2, oc2) == (len(want), nbytes), v["name"]


def test_auto_schedule_gives_the_same_bytes():
    """REDUX_SCHED_AUTO picks the split encoder for small batches and the lane mapping otherwise (decode: always the
    lane mapping); the bytes never depend on the mapping."""
    n, L = 2600, 1500
    raw = rb.generate_blocks_host(0, n, L, SEED)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    model = rb.AdaptiveLinearModel(rb.Parameters(8, 22, 24))
    outs = []
    for sched, count in ((rb.SCHED_AUTO, n), (rb.SCHED_AUTO, 100), (rb.SCHED_LANE, n), (rb.SCHED_WARP, n), (rb.SCHED_SPLIT, n)):
        with rb.Context() as c:
            c.set_schedule(sched)
            comp, coff, st = c.encode_batch(raw[:count * L], off[:count + 1], model)
            back, lens, cons, st = c.decode_batch(comp, coff, off[:count + 1], model)
            assert (back == raw[:count * L]).all()
            outs.append((comp.tobytes(), coff.tobytes()))
    assert outs[0] == outs[2] == outs[3] == outs[4]
    assert outs[2][0].startswith(outs[1][0])


def test_dropin_compress_decompress_doctest():
    """The doctest of src/lib.rs:23-39 through the Python mirror of the crate API."""
    data = bytes([0x72, 0x65, 0x64, 0x75, 0x78])
    compressed = io.BytesIO()
    counts = rb.compress(io.BytesIO(data), compressed, rb.AdaptiveTreeModel.new(rb.Parameters.new(8, 14, 16)))
    assert counts == (5, 7) and compressed.getvalue().hex() == "71f23484c4c510"
    decompressed = io.BytesIO()
    counts = rb.decompress(io.BytesIO(compressed.getvalue()), decompressed,
                           rb.AdaptiveTreeModel.new(rb.Parameters.new(8, 14, 16)))
    assert counts == (7, 5) and decompressed.getvalue() == data


def test_empty_and_tiny_blocks(ctx):
    blocks = [b"", b"a", b"", b"ab", b"\x00", b"\xff" * 3, b""] + [bytes([i]) * i for i in range(1, 40)]
    for params in TRIPLES:
        for cls, ok in KINDS:
            out, out_off, _ = check_encode(ctx, blocks, cls, ok, params)
            check_decode(ctx, out, out_off, blocks, cls, params)


@pytest.mark.parametrize("params", TRIPLES + [(8, 10, 12), (8, 10, 16), (8, 12, 18), (8, 16, 18), (8, 20, 22),
                                              (8, 24, 30), (8, 17, 32), (8, 30, 34), (8, 20, 44), (8, 10, 54),
                                              (8, 10, 30), (8, 12, 32)])      # wide classes that freeze early
def test_ragged_mixed_entropy_batch(ctx, params):
    """Ragged lengths (0..5000), every generator class, both model kinds; freeze regime for small f."""
    rng = np.random.default_rng(params[1] * 100 + params[2])
    n = 300
    lens = rng.integers(0, 5000, size=n)
    lens[:8] = [0, 1, 2, 3, 15, 16, 17, 4999]
    base = rb.generate_blocks_host(0, n, 5000, SEED)
    blocks = [base[i * 5000:i * 5000 + int(lens[i])].tobytes() for i in range(n)]
    for cls, ok in KINDS[1:] if params not in TRIPLES else KINDS:
        out, out_off, _ = check_encode(ctx, blocks, cls, ok, params)
        check_decode(ctx, out, out_off, blocks, cls, params)


def test_full_size_blocks_64k(ctx):
    """The headline shape: 64 KiB blocks (u16 table, last update may wrap node counters)."""
    n = 96
    base = rb.generate_blocks_host(0, n, 65536, SEED)
    blocks = [base[i * 65536:(i + 1) * 65536].tobytes() for i in range(n)]
    blocks[5] = b"\x41" * 65536            # every update lands on one path: increments reach 65535/65536
    blocks[6] = bytes([0x7f]) * 65536
    for params in TRIPLES:
        out, out_off, _ = check_encode(ctx, blocks, rb.AdaptiveTreeModel, o.TREE, params)
        check_decode(ctx, out, out_off, blocks, rb.AdaptiveTreeModel, params)


def test_long_blocks_need_wide_table(ctx):
    """Blocks longer than 65,536 symbols with a late freeze use the 32-bit table."""
    n = 6
    L = 200000
    base = rb.generate_blocks_host(100, n, L, SEED)
    blocks = [base[i * L:(i + 1) * L].tobytes() for i in range(n)]
    blocks.append(b"\x00" * 150000)
    for params in [(8, 30, 32), (8, 20, 22), (8, 16, 18), (8, 14, 16)]:
        out, out_off, _ = check_encode(ctx, blocks, rb.AdaptiveTreeModel, o.TREE, params)
        check_decode(ctx, out, out_off, blocks, rb.AdaptiveTreeModel, params)


def test_unaligned_offsets_and_slack(ctx):
    """Blocks start at arbitrary byte offsets in both directions; raw slots have spare capacity."""
    rng = np.random.default_rng(5)
    blocks = [rng.integers(0, 256, size=int(k), dtype=np.uint8).tobytes() for k in rng.integers(1, 300, size=200)]
    out, out_off, _ = check_encode(ctx, blocks, rb.AdaptiveTreeModel, o.TREE, (8, 22, 24))
    check_decode(ctx, out, out_off, blocks, rb.AdaptiveTreeModel, (8, 22, 24), slack=0)
    check_decode(ctx, out, out_off, blocks, rb.AdaptiveTreeModel, (8, 22, 24), slack=3)


def test_truncated_and_garbage_streams(ctx):
    """Error behaviour of decompress(): Eof on truncation with the decoded prefix kept
    (src/bitio/mod.rs:106-108 via src/codec.rs:50), trailing garbage not read."""
    params = (8, 22, 24)
    model = rb.AdaptiveTreeModel(rb.Parameters(*params))
    data = rb.generate_blocks_host(1, 1, 3000, SEED).tobytes()
    rc, comp, _, _ = o.compress(data, o.TREE, params)
    streams, expect = [], []
    for cut in (0, 1, 2, 3, 4, 5, len(comp) // 3, len(comp) // 2, len(comp) - 2, len(comp) - 1):
        streams.append(comp[:cut])
        expect.append(o.decompress(comp[:cut], o.TREE, params, out_cap=4000))
    streams.append(comp + b"\xde\xad\xbe\xef\x00\x11")
    expect.append(o.decompress(streams[-1], o.TREE, params, out_cap=4000))
    cdata, coff = concat(streams)
    raw_off = np.arange(len(streams) + 1, dtype=np.uint64) * np.uint64(4000)
    raw, raw_lens, consumed, status = ctx.decode_batch(cdata, coff, raw_off, model, check=False)
    for i, (erc, edec, eic, eoc) in enumerate(expect):
        assert int(status[i]) == erc, (i, int(status[i]), erc)
        assert int(raw_lens[i]) == eoc and int(consumed[i]) == eic, (i, int(raw_lens[i]), eoc, int(consumed[i]), eic)
        assert raw[i * 4000:i * 4000 + eoc].tobytes() == edec
    assert int(status[-1]) == 0 and int(consumed[-1]) == len(comp)
    with pytest.raises(rb.Eof):
        ctx.decompress(comp[: len(comp) // 2], model, 4000)
    with pytest.raises(rb.Eof):
        rb.decompress(io.BytesIO(b""), io.BytesIO(), rb.AdaptiveTreeModel(rb.Parameters(8, 14, 16)))


def test_random_garbage_never_faults(ctx):
    """Arbitrary bytes are a valid (if meaningless) code stream: the decoder must agree with the oracle
    on status, lengths and bytes, and stop at the slot capacity."""
    rng = np.random.default_rng(11)
    params = (8, 14, 16)
    model = rb.AdaptiveTreeModel(rb.Parameters(*params))
    streams = [rng.integers(0, 256, size=int(k), dtype=np.uint8).tobytes() for k in rng.integers(0, 400, size=64)]
    cdata, coff = concat(streams)
    cap = 2048
    raw_off = np.arange(len(streams) + 1, dtype=np.uint64) * np.uint64(cap)
    raw, raw_lens, consumed, status = ctx.decode_batch(cdata, coff, raw_off, model, check=False)
    for i, s in enumerate(streams):
        erc, edec, eic, eoc = o.decompress(s, o.TREE, params, out_cap=cap)
        if erc == o.IO_ERROR:      # oracle sink full <-> device OUT_CAPACITY
            assert int(status[i]) == rb.OUT_CAPACITY and int(raw_lens[i]) == cap
            assert raw[i * cap:(i + 1) * cap].tobytes() == edec
        else:
            assert int(status[i]) == erc and int(raw_lens[i]) == eoc and int(consumed[i]) == eic
            assert raw[i * cap:i * cap + eoc].tobytes() == edec


def test_output_capacity_is_reported(ctx):
    blocks = [rb.generate_blocks_host(0, 1, 4096, SEED).tobytes()] * 4
    data, off = concat(blocks)
    model = rb.AdaptiveTreeModel(rb.Parameters(8, 14, 16))
    small = np.zeros(5000, dtype=np.uint8)
    with pytest.raises(rb.OutCapacity):
        ctx.encode_batch(data, off, model, out=small)
    out, out_off, status = ctx.encode_batch(data, off, model, out=small, check=False)
    assert int(status[0]) == 0 and rb.OUT_CAPACITY in [int(s) for s in status]


def test_invalid_and_unsupported_parameters(ctx):
    with pytest.raises(rb.InvalidInput):
        rb.Parameters(8, 9, 16)
    data, off = concat([b"abc"])
    model = rb.AdaptiveTreeModel(rb.Parameters(17, 20, 24))      # valid Parameters beyond the device scope
    with pytest.raises(rb.Unsupported):
        ctx.encode_batch(data, off, model)
    trained = rb.AdaptiveTreeModel(rb.Parameters(8, 10, 12))
    trained.freq = np.full(257, 4, dtype=np.uint32)              # total 1028 > freq_max 1023
    with pytest.raises(rb.InvalidInput):
        ctx.encode_batch(data, off, trained)


def test_generator_device_equals_host(ctx):
    import torch
    n, L = 64, 4099
    d = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    ctx.generate_blocks_device(d, 7, n, L, SEED, device=torch.cuda.current_device(),
                               stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (d.cpu().numpy() == rb.generate_blocks_host(7, n, L, SEED)).all()
    # text class cut from a corpus (what bench.py does with Calgary + Canterbury): device == host == oracle side
    import corpora_fixture as cf
    files = cf.corpora()
    corpus = np.frombuffer(b"".join(files[k] for k in sorted(files) if k.startswith("canterbury/")), dtype=np.uint8)
    ctx.set_text_corpus(corpus)
    try:
        ctx.generate_blocks_device(d, 7, n, L, SEED, device=torch.cuda.current_device(),
                                   stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        host = rb.generate_blocks_host(7, n, L, SEED, corpus=corpus)
        assert (d.cpu().numpy() == host).all() and (host == o.generate_blocks(7, n, L, SEED, corpus=corpus)).all()
        assert corpus.tobytes().find(host[2 * L:3 * L].tobytes()) >= 0          # block index 9: class 1 = a window
    finally:
        ctx.set_text_corpus(None)


def test_device_resident_batch_roundtrip_and_sampled_parity(ctx):
    """Device-resident API at a multi-wave size: 8,192 blocks x 64 KiB generated on the GPU, encoded and
    decoded without leaving HBM; a sample of blocks is memcmp'd with the oracle and the whole batch must
    round-trip (size-independent property)."""
    import torch
    dev = torch.cuda.current_device()
    st = torch.cuda.current_stream().cuda_stream
    n, L = 8192, 65536
    params = (8, 14, 16)
    model = rb.AdaptiveTreeModel(rb.Parameters(*params))
    raw = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    ctx.generate_blocks_device(raw, 0, n, L, SEED, device=dev, stream=st)
    in_off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
    cap = n * (L + L // 8)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    comp_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    status = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    ctx.encode_batch_device(raw, in_off, n, L, comp, cap, comp_off, status, model, device=dev, stream=st)
    torch.cuda.synchronize()
    assert int(status.abs().max()) == 0
    offs = comp_off.cpu().numpy().astype(np.uint64)
    assert int(offs[-1]) <= cap
    sample = [0, 1, 2, 3, 4097, 8191]
    host_raw = raw.cpu().numpy()
    host_comp = comp[: int(offs[-1])].cpu().numpy()
    for i in sample:
        rc, e, _, _ = o.compress(host_raw[i * L:(i + 1) * L], o.TREE, params)
        assert host_comp[int(offs[i]):int(offs[i + 1])].tobytes() == e, i
    back = torch.zeros(n * L, dtype=torch.uint8, device="cuda")
    raw_lens = torch.zeros(n, dtype=torch.int64, device="cuda")
    consumed = torch.zeros(n, dtype=torch.int64, device="cuda")
    status.fill_(-1)
    ctx.decode_batch_device(comp, comp_off, n, L, back, in_off, raw_lens, consumed, status, model, device=dev, stream=st)
    torch.cuda.synchronize()
    assert int(status.abs().max()) == 0
    assert bool((raw_lens == L).all())
    assert bool((consumed == (comp_off[1:] - comp_off[:-1])).all())
    assert torch.equal(back, raw)


# ------------------------------------------------------------------ SURVEY 8(f) rank 4
GENERIC_PARAMS = [(4, 10, 16), (4, 14, 16), (4, 30, 32), (12, 14, 16), (12, 22, 24), (12, 30, 32),
                  (1, 3, 5), (5, 8, 11), (7, 20, 40), (16, 18, 20)]


@pytest.mark.parametrize("params", GENERIC_PARAMS)
def test_generic_symbol_widths(ctx, params):
    """symbol_bits != 8 (the widths of src/model/tests.rs:95-251 and some odd ones): compressed bytes equal
    compress(); decoding equals decompress() including the reference's quirks when 8 % symbol_bits != 0
    (trailing partial symbol dropped on encode, no final flush on decode)."""
    rng = np.random.default_rng(sum(params))
    blocks = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in (0, 1, 2, 3, 7, 64, 333, 1000)]
    blocks += [bytes(rng.choice(np.frombuffer(b"abcd \n", dtype=np.uint8), 500)), bytes([0xFF] * 257)]
    for cls, okind in KINDS:
        model = cls(rb.Parameters(*params))
        data, off = concat(blocks)
        out, out_off, status = ctx.encode_batch(data, off, model)
        assert (status == 0).all()
        want, back = [], []
        for i, b in enumerate(blocks):
            rc, e, ic, oc = o.compress(b, okind, params)
            assert rc == o.OK
            assert out[int(out_off[i]):int(out_off[i + 1])].tobytes() == e, (params, i)
            rc, d, ic2, oc2 = o.decompress(e, okind, params, out_cap=len(b) + 8)
            assert rc == o.OK
            back.append(d)
        raw_off = np.zeros(len(blocks) + 1, dtype=np.uint64)
        np.cumsum(np.array([len(b) + 8 for b in blocks], dtype=np.uint64), out=raw_off[1:])
        raw, raw_lens, consumed, status = ctx.decode_batch(out, out_off, raw_off, model)
        assert (status == 0).all()
        for i, d in enumerate(back):
            assert raw[int(raw_off[i]):int(raw_off[i]) + int(raw_lens[i])].tobytes() == d, (params, i)
            assert int(consumed[i]) == int(out_off[i + 1] - out_off[i])


@pytest.mark.parametrize("params", [(8, 14, 16), (8, 30, 32), (4, 10, 16), (12, 22, 24)])
def test_pretrained_models(ctx, params):
    """A model trained through get_frequency() before compress()/decompress() receive it (the reference's
    Box<Model> may arrive in any state, src/lib.rs:102 + src/model/mod.rs:23-25)."""
    s, f, c = params
    rng = np.random.default_rng(77 + s)
    train = [int(x) for x in rng.integers(0, min((1 << s) + 1, 40), 1500)]
    blocks = [bytes(rng.integers(0, 40, n, dtype=np.uint8)) for n in (0, 1, 5, 100, 2000)]
    for cls, okind in KINDS:
        model = cls(rb.Parameters(*params)).train(train)
        assert (model.freq == o.trained_frequencies(train, okind, params)).all()
        assert model.total_frequency() == min((1 << s) + 1 + len(train), (1 << f) - 1)
        single = cls(rb.Parameters(*params))
        for t in train[:20]:
            single.get_frequency(t)                              # the reference's own training call
        assert (single.freq == o.trained_frequencies(train[:20], okind, params)).all()
        data, off = concat(blocks)
        out, out_off, status = ctx.encode_batch(data, off, model)
        assert (status == 0).all()
        raw_off = np.zeros(len(blocks) + 1, dtype=np.uint64)
        np.cumsum(np.array([len(b) + 4 for b in blocks], dtype=np.uint64), out=raw_off[1:])
        raw, raw_lens, consumed, status = ctx.decode_batch(out, out_off, raw_off, model)
        assert (status == 0).all()
        for i, b in enumerate(blocks):
            rc, e, ic, oc = o.compress_trained(b, train, okind, params)
            assert rc == o.OK and out[int(out_off[i]):int(out_off[i + 1])].tobytes() == e, (params, i)
            rc, d, ic2, oc2 = o.decompress_trained(e, train, okind, params, out_cap=len(b) + 4)
            assert raw[int(raw_off[i]):int(raw_off[i]) + int(raw_lens[i])].tobytes() == d
        # the drop-in pair with a trained model
        cbuf = io.BytesIO()
        counts = rb.compress(io.BytesIO(blocks[3]), cbuf, cls(rb.Parameters(*params)).train(train), context=ctx)
        assert cbuf.getvalue() == o.compress_trained(blocks[3], train, okind, params)[1] and counts[0] == len(blocks[3])
        dbuf = io.BytesIO()
        rb.decompress(io.BytesIO(cbuf.getvalue()), dbuf, cls(rb.Parameters(*params)).train(train), context=ctx)
        assert dbuf.getvalue() == o.decompress_trained(cbuf.getvalue(), train, okind, params)[1]


def test_generic_many_blocks_share_columns(ctx):
    """More blocks than the generic kernel has threads per launch slot reuse: 5,000 blocks of 12-bit symbols."""
    params = (12, 14, 16)
    n, L = 5000, 96
    raw = rb.generate_blocks_host(0, n, L, SEED)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    model = rb.AdaptiveTreeModel(rb.Parameters(*params))
    out, out_off, status = ctx.encode_batch(raw, off, model)
    assert (status == 0).all()
    for i in (0, 1, 127, 128, 4095, 4999):
        assert out[int(out_off[i]):int(out_off[i + 1])].tobytes() == o.compress(raw[i * L:(i + 1) * L], o.TREE, params)[1]
    back, lens, cons, status = ctx.decode_batch(out, out_off, off, model)
    assert (status == 0).all() and (back[: n * L] == raw).all()      # 96 bytes = 64 whole 12-bit symbols


def test_decode_time_does_not_depend_on_slot_alignment():
    """Lanes of one warp decode into slots at different byte phases (any ragged batch).  Round 1 walked each lane to a
    word boundary first and the warp never reconverged afterwards: it ran the whole stream once per phase (4.0x,
    scripts/bench_alignment.py).  The phase-free sink keeps the warp together: same bytes, and the mixed-phase launch
    takes as long as the aligned one (generous bound: 1.5x; both launches are timed back to back on one stream)."""
    import torch
    n, L = 32, 65536
    params = (8, 14, 16)
    model = rb.AdaptiveTreeModel(rb.Parameters(*params))
    blocks = [rb.generate_blocks_host(1 + 4 * i, 1, L, SEED) for i in range(n)]
    with rb.Context([0]) as c:
        c.set_schedule(rb.SCHED_LANE)
        data, off = concat([b.tobytes() for b in blocks])
        comp, coff, st = c.encode_batch(data, off, model)
        stream = torch.cuda.current_stream().cuda_stream
        d_comp = torch.from_numpy(np.concatenate([comp, np.zeros(64, np.uint8)])).cuda()
        d_coff = torch.from_numpy(coff.astype(np.int64)).cuda()
        times = {}
        for name, pads in (("aligned", [16] * n), ("mixed", [((i % 4) + 1) % 4 for i in range(n)])):
            roff = np.zeros(n + 1, dtype=np.int64)
            for i in range(n):
                roff[i + 1] = roff[i] + L + pads[i]
            d_roff = torch.from_numpy(roff).cuda()
            back = torch.zeros(int(roff[-1]) + 64, dtype=torch.uint8, device="cuda")
            rl = torch.zeros(n, dtype=torch.int64, device="cuda"); cons = torch.zeros(n, dtype=torch.int64, device="cuda")
            dst = torch.zeros(n, dtype=torch.int32, device="cuda")
            run = lambda: c.decode_batch_device(d_comp, d_coff, n, L + 16, back, d_roff, rl, cons, dst, model, device=0, stream=stream)
            run(); torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record(); run(); run(); ev[1].record(); ev[1].synchronize()
            times[name] = ev[0].elapsed_time(ev[1]) / 2
            assert int(dst.abs().max()) == 0 and bool((rl == L).all())
            host = back.cpu().numpy()
            for i in (0, 1, 2, 3, 17, 31):
                assert (host[int(roff[i]):int(roff[i]) + L] == blocks[i]).all(), (name, i)
                if pads[i]:
                    assert (host[int(roff[i]) + L:int(roff[i + 1])] == 0).all(), "bytes between the slots were written"
        assert times["mixed"] < 1.5 * times["aligned"], times


def test_pinned_host_buffers_and_registered_buffers():
    """redux_host_alloc / redux_host_register: the same bytes come out whatever kind of host memory the caller hands
    in (pageable numpy, page-locked by the library, page-locked in place)."""
    n, L = 300, 3000
    raw = rb.generate_blocks_host(0, n, L, SEED)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    model = rb.AdaptiveTreeModel(rb.Parameters(8, 14, 16))
    with rb.Context([0]) as c:
        ref, ref_off, _ = c.encode_batch(raw, off, model)
        hb = rb.HostBuffer(raw.size)
        hb.array[:] = raw
        out = rb.HostBuffer(ref.size + 4096)
        comp, coff, st = c.encode_batch(hb.array, off, model, out=out.array)
        assert comp.tobytes() == ref.tobytes() and (coff == ref_off).all() and (st == 0).all()
        reg = raw.copy()
        rb.host_register(reg)
        try:
            comp2, coff2, _ = c.encode_batch(reg, off, model)
            assert comp2.tobytes() == ref.tobytes()
            back, rl, cons, st = c.decode_batch(comp2, coff2, off, model, raw=reg)      # decode into the registered buffer
            assert (st == 0).all() and (reg == raw).all()
        finally:
            rb.host_unregister(reg)
        comp = coff = None
        hb.close(); out.close()


def test_pageable_buffers_staged_by_the_library():
    """Pageable caller memory goes through the library's pinned ring and copy threads (redux_ctx_set_staging): same
    bytes, offsets, counts and statuses as the driver-staged path and the oracle -- with pieces small enough that the
    ring wraps many times inside every chunk, for pageable input, output, or both, several chunks per call."""
    rng = np.random.default_rng(11)
    n = 6000                                                  # > 3 chunks of whole CTAs
    lens = rng.integers(0, 700, n)
    blocks = [rb.generate_blocks_host(i, 1, int(L), SEED).tobytes() if L else b"" for i, L in enumerate(lens)]
    data, off = concat(blocks)
    for params in ((8, 14, 16), (8, 30, 32)):
        model = rb.AdaptiveTreeModel(rb.Parameters(*params))
        with rb.Context([0]) as c:
            c.set_staging(False)
            ref, ref_off, ref_st = c.encode_batch(data, off, model)
            for i in (0, 1, 2, 999, n - 1):
                assert ref[int(ref_off[i]):int(ref_off[i + 1])].tobytes() == o.compress(blocks[i], o.TREE, params)[1]
            raw_off = off.copy()
            pin_in = rb.HostBuffer(data.size); pin_in.array[:] = data
            pin_out = rb.HostBuffer(ref.size + 64)
            pin_back = rb.HostBuffer(data.size)
            c.set_staging(True, min_bytes=1, piece_bytes=4096 * 3, slots=3, threads=3)
            for src, dst in ((data, None), (pin_in.array, None), (data, pin_out.array)):
                comp, coff, st = c.encode_batch(src, off, model, out=dst)
                assert (coff == ref_off).all() and (st == ref_st).all()
                assert comp[:int(coff[-1])].tobytes() == ref.tobytes()
            # decode: a truncated and a garbage stream among the good ones keep their statuses and counts
            streams = [ref[int(ref_off[i]):int(ref_off[i + 1])].tobytes() for i in range(n)]
            streams[5] = streams[5][:-3]
            streams[7] = bytes(rng.integers(0, 256, 400, dtype=np.uint8))
            bad, bad_off = concat(streams)
            c.set_staging(False)
            want = c.decode_batch(bad, bad_off, raw_off, model, check=False)
            assert want[3][5] != 0 and (np.delete(want[3], [5, 7]) == 0).all()
            c.set_staging(True, min_bytes=1, piece_bytes=4096 * 3, slots=3, threads=3)
            pin_comp = rb.HostBuffer(bad.size); pin_comp.array[:] = bad
            for src, dst in ((bad, None), (pin_comp.array, None), (bad, pin_back.array)):
                got = c.decode_batch(src, bad_off, raw_off, model, raw=dst, check=False)
                for a, b in zip(got[1:], want[1:]):
                    assert (a == b).all()
                for i in range(n):
                    lo = int(raw_off[i])
                    assert got[0][lo:lo + int(got[1][i])].tobytes() == want[0][lo:lo + int(want[1][i])].tobytes(), i
                    if i not in (5, 7):
                        assert got[0][lo:lo + int(got[1][i])].tobytes() == blocks[i]
            pin_comp.close()
            # a short capacity is reported the same way
            small = np.zeros(ref.size // 2, dtype=np.uint8)
            with pytest.raises(rb.OutCapacity):
                c.encode_batch(data, off, model, out=small)
            pin_in.close(); pin_out.close(); pin_back.close()


def test_pageable_staging_multi_megabyte_default_pieces():
    """Default staging parameters on a batch big enough to engage them (>= 8 MiB each way): round trip + oracle sample."""
    n, L = 4096, 8192                                         # 32 MiB
    raw = rb.generate_blocks_host(0, n, L, SEED)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    model = rb.AdaptiveTreeModel(rb.Parameters(8, 14, 16))
    with rb.Context([0]) as c:
        comp, coff, st = c.encode_batch(raw, off, model)
        assert (st == 0).all()
        for i in (0, 1, 2, 3, 2047, n - 1):
            assert comp[int(coff[i]):int(coff[i + 1])].tobytes() == o.compress(raw[i * L:(i + 1) * L].tobytes(), o.TREE, (8, 14, 16))[1]
        c.set_staging(False)
        comp2, coff2, _ = c.encode_batch(raw, off, model)
        assert comp2.tobytes() == comp.tobytes() and (coff2 == coff).all()
        c.set_staging(True)
        back, rl, cons, st = c.decode_batch(comp, coff, off, model)
        assert (st == 0).all() and (rl == L).all() and back.tobytes() == raw.tobytes()


def test_device_entry_points_take_a_trained_model():
    """redux_encode_batch_device_ex / redux_decode_batch_device_ex: a model trained before the call, data resident in
    HBM; bytes equal the oracle's compress() from the same trained model."""
    import torch
    n, L = 64, 5000
    raw = rb.generate_blocks_host(0, n, L, SEED)
    train = [int(x) for x in raw[:700]]
    for params in ((8, 14, 16), (8, 30, 32)):
        model = rb.AdaptiveTreeModel(rb.Parameters(*params)).train(train)
        with rb.Context([0]) as c:
            stream = torch.cuda.current_stream().cuda_stream
            d_raw = torch.from_numpy(raw).cuda()
            d_off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
            cap = 2 * n * L
            d_comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); d_coff = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
            d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
            c.encode_batch_device(d_raw, d_off, n, L, d_comp, cap, d_coff, d_st, model, device=0, stream=stream)
            d_back = torch.empty(n * L, dtype=torch.uint8, device="cuda")
            rl = torch.zeros(n, dtype=torch.int64, device="cuda"); cons = torch.zeros(n, dtype=torch.int64, device="cuda")
            c.decode_batch_device(d_comp, d_coff, n, L, d_back, d_off, rl, cons, d_st, model, device=0, stream=stream)
            torch.cuda.synchronize()
            assert int(d_st.abs().max()) == 0 and torch.equal(d_back, d_raw)
            comp, coff = d_comp.cpu().numpy(), d_coff.cpu().numpy()
            for i in (0, 1, 31, 63):
                want = o.compress_trained(raw[i * L:(i + 1) * L], train, o.TREE, params)[1]
                assert comp[int(coff[i]):int(coff[i + 1])].tobytes() == want, (params, i)


def test_device_calls_on_different_streams_do_not_share_the_workspace_unordered():
    """One workspace per device: two encodes enqueued back to back on DIFFERENT streams must not overwrite each
    other's slots; the library orders them on the device."""
    import torch
    n, L = 2048, 4096
    model = rb.AdaptiveTreeModel(rb.Parameters(8, 14, 16))
    with rb.Context([0]) as c:
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        bufs = []
        for k, st in enumerate((s1, s2)):
            d_raw = torch.from_numpy(rb.generate_blocks_host(k * n, n, L, SEED)).cuda()
            bufs.append((st, d_raw, torch.empty(2 * n * L, dtype=torch.uint8, device="cuda"),
                         torch.zeros(n + 1, dtype=torch.int64, device="cuda"), torch.zeros(n, dtype=torch.int32, device="cuda")))
        d_off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
        torch.cuda.synchronize()
        for st, d_raw, d_comp, d_coff, d_st in bufs:            # enqueued without any host synchronisation in between
            c.encode_batch_device(d_raw, d_off, n, L, d_comp, 2 * n * L, d_coff, d_st, model, device=0, stream=st.cuda_stream)
        torch.cuda.synchronize()
        for k, (st, d_raw, d_comp, d_coff, d_st) in enumerate(bufs):
            comp, coff = d_comp.cpu().numpy(), d_coff.cpu().numpy()
            raw = d_raw.cpu().numpy()
            for i in (0, 7, 1000, n - 1):
                assert comp[int(coff[i]):int(coff[i + 1])].tobytes() == o.compress(raw[i * L:(i + 1) * L], o.TREE, (8, 14, 16))[1], (k, i)


def test_two_ctas_per_sm_fit_the_shared_memory_budget():
    """The tuned kernels' shared memory (7 x 16 KiB tables + one 4-byte staging slot per thread) is sized to the
    byte for two CTAs per SM = 448 resident streams per SM, which is what puts 65,536 blocks in ONE wave on 148
    SMs.  One CTA per SM would silently halve the throughput."""
    import ctypes as C
    with rb.Context([0]):
        enc, dec = C.c_int(0), C.c_int(0)
        assert rb.lib().redux_debug_lane_occupancy(C.byref(enc), C.byref(dec)) == rb.OK
        assert (enc.value, dec.value) == (2, 2)


def test_multi_device_context_shards_by_block_ranges():
    """one context over all visible GPUs, one host thread per device, contiguous block ranges,
    no collective.  The bytes, offsets and statuses must not depend on the number of devices.  Needs >= 2 GPUs
    (skipped on a 1-GPU box; `gpurun --gpus 2` runs it)."""
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs at least 2 GPUs")
    n, L = 3001, 2500                                   # odd count: uneven shards
    raw = rb.generate_blocks_host(0, n, L, SEED)
    lens = np.full(n, L, dtype=np.uint64); lens[::7] = 0; lens[5::11] = 17      # ragged, some empty
    off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(lens, out=off[1:])
    data = np.concatenate([raw[i * L:i * L + int(lens[i])] for i in range(n)])
    for params, trained in (((8, 14, 16), False), ((8, 30, 32), False), ((12, 22, 24), False), ((8, 14, 16), True)):
        model = rb.AdaptiveTreeModel(rb.Parameters(*params))
        train = [int(x) for x in raw[:500]] if trained else None
        if trained:
            model.train(train)
        with rb.Context([0]) as one:
            ref_comp, ref_off, ref_st = one.encode_batch(data, off, model)
        with rb.Context(list(range(ng))) as many:
            assert many.device_count == ng
            comp, coff, st = many.encode_batch(data, off, model)
            assert comp.tobytes() == ref_comp.tobytes() and (coff == ref_off).all() and (st == 0).all()
            back, rl, cons, st = many.decode_batch(comp, coff, off, model)
            assert (st == 0).all() and (cons == coff[1:] - coff[:-1]).all()
            if params[0] == 8:
                assert (rl == lens).all() and (back[: int(off[-1])] == data).all()
            # the same through the pageable-memory staging (feeder thread + ring per device, shard hand-off between
            # the drainers): forced on with pieces small enough to wrap the rings many times
            many.set_staging(True, min_bytes=1, piece_bytes=4096 * 5, slots=3, threads=3)
            comp2, coff2, st2 = many.encode_batch(data, off, model)
            assert comp2.tobytes() == ref_comp.tobytes() and (coff2 == ref_off).all() and (st2 == 0).all()
            back2, rl2, cons2, st2 = many.decode_batch(comp2, coff2, off, model)
            assert (st2 == 0).all() and (rl2 == rl).all() and (cons2 == cons).all()
            if params[0] == 8:
                assert (back2[: int(off[-1])] == data).all()
        for i in (0, 1, 2, 1500, 3000):
            b = data[int(off[i]):int(off[i + 1])]
            want = o.compress_trained(b, train, o.TREE, params)[1] if trained else o.compress(b, o.TREE, params)[1]
            assert comp[int(coff[i]):int(coff[i + 1])].tobytes() == want, (params, i)


def test_frozen_decoder_next_to_threads_without_a_block():
    """The frozen narrow decoder reads "entry 256" of its cumulative array from the neighbouring warp's column (or the
    zero row after the last table).  That halfword must read zero also when the neighbouring threads have no block
    (partially filled last CTA / warp) and the shared memory still holds another kernel's data: blocks long enough to
    freeze, full of the symbols 252..255 whose group touches entry 256, in counts that leave warps and lanes idle."""
    rng = np.random.default_rng(256)
    params = (8, 10, 12)                                      # freezes after 766 symbols
    model = rb.AdaptiveTreeModel(rb.Parameters(*params))
    with rb.Context([0]) as c:
        c.set_schedule(rb.SCHED_LANE)
        for n in (1, 33, 224 + 96, 224 * 3 + 129):
            blocks = [bytes(rng.integers(248, 256, 3000, dtype=np.uint8)) for _ in range(n)]
            data, off = concat(blocks)
            comp, coff, st = c.encode_batch(data, off, model)            # leaves its tables in shared memory
            assert (st == 0).all()
            back, rl, cons, st = c.decode_batch(comp, coff, off, model)
            assert (st == 0).all() and (rl == 3000).all() and back.tobytes() == data.tobytes(), n
            for i in (0, n - 1):
                assert comp[int(coff[i]):int(coff[i + 1])].tobytes() == o.compress(blocks[i], o.TREE, params)[1]


def test_million_symbol_block_crosses_the_quotient_bound(ctx):
    """A 1.15 MB block with a late freeze: the total frequency passes 2^20, where the 64-bit-product decoders
    stop trusting the float estimate of value = X / range (lane: product-domain descent; warp: exact 64-bit
    division).  Also a slot one byte too small: OUT_CAPACITY with the prefix in place."""
    rng = np.random.default_rng(41)
    big = np.concatenate([rng.integers(0, 256, 400000, dtype=np.uint8),
                          rng.choice(np.frombuffer(b"acgt", dtype=np.uint8), 400000),
                          np.minimum(rng.geometric(0.3, 350000) - 1, 255).astype(np.uint8)]).tobytes()
    blocks = [big, b"short block"]
    for params in ((8, 22, 24), (8, 30, 32)):
        out, out_off, expect = check_encode(ctx, blocks, rb.AdaptiveTreeModel, o.TREE, params)
        check_decode(ctx, out, out_off, blocks, rb.AdaptiveTreeModel, params)
        model = rb.AdaptiveTreeModel(rb.Parameters(*params))
        raw_off = np.array([0, len(big) - 1, len(big) - 1 + len(blocks[1])], dtype=np.uint64)
        raw, raw_lens, consumed, status = ctx.decode_batch(out, out_off, raw_off, model, check=False)
        assert [int(x) for x in status] == [rb.OUT_CAPACITY, 0]
        assert int(raw_lens[0]) == len(big) - 1 and raw[: len(big) - 1].tobytes() == big[:-1]
        assert raw[len(big) - 1: len(big) - 1 + len(blocks[1])].tobytes() == blocks[1]


def test_generic_threads_reuse_their_columns():
    """16-bit symbols: a Fenwick column is 256 KiB, so 2 GB of columns hold 8,192 threads; 9,000 blocks make
    some threads code two blocks in turn and reset their column in between."""
    params = (16, 18, 20)
    n, L = 9000, 48
    raw = rb.generate_blocks_host(3, n, L, SEED)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    model = rb.AdaptiveTreeModel(rb.Parameters(*params))
    with rb.Context() as c:
        out, out_off, status = c.encode_batch(raw, off, model)
        assert (status == 0).all()
        for i in (0, 1, 8191, 8192, 8500, 8999):
            assert out[int(out_off[i]):int(out_off[i + 1])].tobytes() == o.compress(raw[i * L:(i + 1) * L], o.TREE, params)[1], i
        back, lens, cons, status = c.decode_batch(out, out_off, off, model)
        assert (status == 0).all() and (lens == L).all() and (back[: n * L] == raw).all()


def test_pretrained_model_with_totals_beyond_u16(ctx):
    """A model trained on 70,000 symbols (total 70,257 > 65,535): the tuned lane kernels start from u32 table
    entries; with freq_bits = 16 the same training freezes the model before the first coded symbol."""
    rng = np.random.default_rng(99)
    train = [int(x) for x in rng.integers(60, 70, 70000)]
    blocks = [bytes(rng.integers(55, 75, n, dtype=np.uint8)) for n in (0, 1, 1000, 20000)] + [bytes([255] * 300)]
    for params in ((8, 30, 32), (8, 22, 24), (8, 16, 18)):
        model = rb.AdaptiveTreeModel(rb.Parameters(*params)).train(train)
        assert model.total_frequency() == min(257 + len(train), (1 << params[1]) - 1)
        data, off = concat(blocks)
        out, out_off, status = ctx.encode_batch(data, off, model)
        assert (status == 0).all()
        raw_off = np.zeros(len(blocks) + 1, dtype=np.uint64)
        np.cumsum(np.array([len(b) for b in blocks], dtype=np.uint64), out=raw_off[1:])
        raw, raw_lens, consumed, status = ctx.decode_batch(out, out_off, raw_off, model)
        assert (status == 0).all()
        for i, b in enumerate(blocks):
            want = o.compress_trained(b, train, o.TREE, params)[1]
            assert out[int(out_off[i]):int(out_off[i + 1])].tobytes() == want, (params, i)
            assert raw[int(raw_off[i]):int(raw_off[i]) + len(b)].tobytes() == b and int(consumed[i]) == len(want)
