"""BASELINE.json configs 1, 2 and 5 on the reference's own corpora, on the GPU, bit-exact.  `pytest -m gpu`.

The reference's integration suite (tests/corpora.rs:87-259) round-trips every file under resources/* with
AdaptiveLinearModel and AdaptiveTreeModel at Parameters(8, bits, bits+2), bits in {14, 22, 30} (:35), asserting
the returned byte counts (:40-41) and the decoded bytes (:59,61).  Here the same files (the fixture
tests/golden/corpora/corpora.tar.xz, byte-identical to the reference tree) go through the C ABI as one batch,
one stream per file, on every stream-to-thread mapping, and on top of the reference's assertions every compressed
stream must equal the CPU oracle's byte for byte and SURVEY.md's B.2 size/SHA table."""
import hashlib
import os

import numpy as np
import pytest

import corpora_fixture as cf
import oracle_lib as o
import redux_b200 as rb

pytestmark = pytest.mark.gpu
TRIPLES = [(8, 14, 16), (8, 22, 24), (8, 30, 32)]          # tests/corpora.rs:35
KINDS = [(rb.AdaptiveLinearModel, o.LINEAR), (rb.AdaptiveTreeModel, o.TREE)]
SWEEP = [(10, 16), (14, 16), (16, 18), (20, 22), (22, 24), (24, 30), (30, 32)]
MIB = 1 << 20
_oracle_cache = {}


def host_threads():
    return max(1, min(len(os.sched_getaffinity(0)), 64))


def batch_of(blocks):
    lens = [len(b) for b in blocks]
    off = np.zeros(len(blocks) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return np.frombuffer(b"".join(blocks), dtype=np.uint8), off


def oracle_streams(tag, data, off, okind, params):
    """The oracle's compressed stream of every block (cached per input set / model kind / parameters)."""
    key = (tag, okind, params)
    if key not in _oracle_cache:
        rc, slots, slot_off, out_len, status = o.compress_batch(data, off, okind, params, host_threads())
        assert rc == 0 and (status == 0).all()
        _oracle_cache[key] = [slots[int(slot_off[i]):int(slot_off[i]) + int(out_len[i])].tobytes()
                              for i in range(len(off) - 1)]
    return _oracle_cache[key]


@pytest.fixture(scope="module", params=["lane", "warp", "split"])
def ctx(request):
    c = rb.Context()
    c.set_schedule({"lane": rb.SCHED_LANE, "warp": rb.SCHED_WARP, "split": rb.SCHED_SPLIT}[request.param])
    yield c
    c.close()


@pytest.mark.parametrize("kind", KINDS, ids=["linear", "tree"])
@pytest.mark.parametrize("params", TRIPLES, ids=lambda p: "%d-%d-%d" % p)
def test_every_corpus_file_as_one_stream(ctx, params, kind):
    """Configs 1 + 2 (and the artificial / large / misc corpora of tests/corpora.rs): 36 files, one stream each."""
    model_cls, okind = kind
    files = cf.corpora()
    names = sorted(files)
    data, off = batch_of([files[n] for n in names])
    table = {r["file"]: r for r in cf.corpus_table()}
    triple = "%d,%d,%d" % params
    comp, comp_off, status = ctx.encode_batch(data, off, model_cls(rb.Parameters(*params)))
    assert (status == 0).all()
    want = oracle_streams("corpora", data, off, okind, params)
    for i, n in enumerate(names):
        got = comp[int(comp_off[i]):int(comp_off[i + 1])].tobytes()
        assert len(got) == table[n][triple][0], (n, len(got))                    # second element of compress()'s tuple
        assert hashlib.sha256(got).hexdigest()[:16] == table[n][triple][1], n
        assert got == want[i], n                                                 # the oracle, every byte
    back, raw_lens, consumed, status = ctx.decode_batch(comp, comp_off, off, model_cls(rb.Parameters(*params)))
    assert (status == 0).all()
    assert (consumed == comp_off[1:] - comp_off[:-1]).all()      # tests/corpora.rs:40: the whole stream is consumed
    assert (raw_lens == off[1:] - off[:-1]).all()                # :41
    assert back[:data.size].tobytes() == data.tobytes()          # :59, :61


def test_book1_with_the_cli_parameters():
    """Config 1 as the reference's CLI runs it: calgary/book1, AdaptiveTreeModel, Parameters(8,30,32)
    (src/main.rs:108), through the single-stream drop-in pair; sizes from BASELINE.md section 2."""
    book1 = cf.corpora()["calgary/book1"]
    with rb.Context() as c:
        comp, (ic, oc) = c.compress(book1, rb.AdaptiveTreeModel(rb.Parameters(8, 30, 32)))
        assert (ic, oc) == (768771, 435400)
        assert comp == o.compress(book1, o.TREE, (8, 30, 32))[1]
        raw, (ic, oc) = c.decompress(comp, rb.AdaptiveTreeModel(rb.Parameters(8, 30, 32)), len(book1) + 64)
        assert (ic, oc) == (435400, 768771) and raw == book1


def config5_blocks():
    files = cf.corpora()
    blocks = cf.blocks_of(files["large/bible.txt"], MIB) + cf.blocks_of(files["large/world192.txt"], MIB) \
        + cf.blocks_of(cf.ecoli_stand_in(), MIB)
    assert [len(cf.blocks_of(files["large/bible.txt"], MIB)), len(cf.blocks_of(files["large/world192.txt"], MIB)),
            len(cf.blocks_of(cf.ecoli_stand_in(), MIB))] == [4, 3, 5]
    return blocks


@pytest.mark.parametrize("fc", SWEEP, ids=lambda fc: "f%d-c%d" % fc)
def test_large_corpus_in_1mib_blocks_over_all_devices(fc):
    """Config 5: resources/large (bible.txt, world192.txt, seeded E.coli stand-in) in 1 MiB blocks, the
    frequency_bits / code_bits sweep, the 12 blocks sharded over every GPU of the box by ONE context
    (SURVEY.md 8(e)); every block's stream equals the oracle's, every block round-trips."""
    import torch
    params = (8,) + fc
    blocks = config5_blocks()
    data, off = batch_of(blocks)
    want = oracle_streams("config5", data, off, o.TREE, params)
    with rb.Context(list(range(torch.cuda.device_count()))) as c:
        model = rb.AdaptiveTreeModel(rb.Parameters(*params))
        comp, comp_off, status = c.encode_batch(data, off, model)
        assert (status == 0).all()
        for i in range(len(blocks)):
            assert comp[int(comp_off[i]):int(comp_off[i + 1])].tobytes() == want[i], (params, i)
        back, raw_lens, consumed, status = c.decode_batch(comp, comp_off, off, model)
        assert (status == 0).all() and (raw_lens == off[1:] - off[:-1]).all()
        assert (consumed == comp_off[1:] - comp_off[:-1]).all()
        assert back[:data.size].tobytes() == data.tobytes()
