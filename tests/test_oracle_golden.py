"""Oracle vs committed golden vectors (CPU).  See tests/golden/make_golden.py for provenance."""
import json
import os

import pytest

import oracle_lib as o

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = json.load(open(os.path.join(GOLD, "kat_vectors.json")))


@pytest.mark.parametrize("kind", [o.LINEAR, o.TREE])
def test_kat_vectors_compress_and_roundtrip(kind):
    """Every golden vector: compress() bytes equal, (in,out) counts equal the real lengths
    (tests/corpora.rs:40-41), decompress() consumes the whole stream and returns the input."""
    for v in KAT:
        s, f, c = v["params"]
        data = bytes.fromhex(v["input"])
        want = bytes.fromhex(v["compressed"])
        train = list(bytes.fromhex(v["train"])) if "train" in v else None
        if train is None:
            rc, out, ic, oc = o.compress(data, kind, (s, f, c))
        else:   # the caller trained the model first (src/model/mod.rs:23-25)
            rc, out, ic, oc = o.compress_trained(data, train, kind, (s, f, c))
        assert rc == o.OK, v["name"]
        assert out == want, (v["name"], v["params"])
        assert oc == len(want)
        assert ic == len(data)
        if train is None:
            rc, dec, ic2, oc2 = o.decompress(out, kind, (s, f, c), out_cap=len(data) + 16)
        else:
            rc, dec, ic2, oc2 = o.decompress_trained(out, train, kind, (s, f, c), out_cap=len(data) + 16)
        assert rc == o.OK
        assert ic2 == len(out), "decoder must consume exactly the compressed stream"
        nbits = (len(data) * 8 // s) * s  # trailing partial symbol is dropped
        if s == 8:
            assert dec == data and oc2 == len(data)
        else:
            # decompress never flushes (src/codec.rs:164-176): only whole output bytes appear
            assert oc2 == nbits // 8
            assert dec == data[:nbits // 8]


def test_doctest_redux_roundtrip():
    """src/lib.rs:23-39."""
    data = bytes([0x72, 0x65, 0x64, 0x75, 0x78])
    rc, comp, _, _ = o.compress(data, o.TREE, (8, 14, 16))
    assert rc == o.OK
    rc, dec, _, _ = o.decompress(comp, o.TREE, (8, 14, 16))
    assert rc == o.OK and dec == data


def test_decompress_empty_is_eof():
    rc, dec, ic, oc = o.decompress(b"", o.TREE, (8, 14, 16))
    assert rc == o.EOF and dec == b"" and ic == 0 and oc == 0


def test_truncated_stream_is_eof():
    data = bytes(range(256)) * 4
    rc, comp, _, _ = o.compress(data, o.TREE, (8, 22, 24))
    assert rc == o.OK
    for cut in (1, 2, 3, len(comp) // 2, len(comp) - 1):
        rc, dec, ic, oc = o.decompress(comp[:cut], o.TREE, (8, 22, 24), out_cap=len(data) + 16)
        assert rc == o.EOF
        assert ic == cut
        assert data.startswith(dec)


def test_trailing_garbage_not_read():
    data = b"hello hello hello"
    rc, comp, _, _ = o.compress(data, o.TREE, (8, 30, 32))
    rc, dec, ic, oc = o.decompress(comp + b"\xde\xad\xbe\xef", o.TREE, (8, 30, 32))
    assert rc == o.OK and dec == data and ic == len(comp)


def test_output_full_is_io_error():
    data = bytes(range(200))
    rc, out, ic, oc = o.compress(data, o.TREE, (8, 14, 16), out_cap=10)
    assert rc == o.IO_ERROR and oc == 10


@pytest.mark.parametrize("sfc,ok", [
    ((8, 14, 16), True), ((8, 30, 32), True), ((8, 10, 12), True), ((8, 31, 33), True), ((8, 30, 34), True),
    ((0, 14, 16), False), ((8, 9, 16), False), ((8, 14, 15), False), ((8, 32, 34), False), ((8, 31, 34), False),
    ((1, 3, 5), True), ((1, 2, 5), False), ((12, 14, 16), True), ((12, 13, 16), False),
])
def test_parameters_new_validation(sfc, ok):
    """src/model/mod.rs:64."""
    rc, p = o.params_new(*sfc)
    assert (rc == o.OK) == ok
    if ok:
        s, f, c = sfc
        assert p.symbol_eof == 1 << s and p.symbol_count == (1 << s) + 1
        assert p.freq_max == (1 << f) - 1
        assert p.code_one_fourth == 1 << (c - 2) and p.code_half == 2 << (c - 2)
        assert p.code_three_fourths == 3 << (c - 2) and p.code_max == (1 << c) - 1 and p.code_min == 0
