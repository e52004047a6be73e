"""The reference's byte-exact bit I/O vectors (src/bitio/tests.rs:8-218) replayed on the oracle."""
import ctypes as C
import json
import os

import numpy as np

import oracle_lib as o

V = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bitio_vectors.json")))


def test_writer_vectors():
    L = o.lib()
    for case in V["writer"]:
        buf = np.zeros(16, dtype=np.uint8)
        w = L.oracle_bitwriter_new(buf.ctypes.data, buf.size)
        for op, cnt in zip(case["ops"], case["counts"]):
            rc = L.oracle_bitwriter_flush_bits(w) if op[0] == "f" else L.oracle_bitwriter_write_bits(w, op[1], op[2])
            assert rc == o.OK, case["name"]
            assert L.oracle_bitwriter_get_count(w) == cnt, case["name"]
        n = L.oracle_bitwriter_get_count(w)
        assert buf[:n].tobytes().hex() == case["bytes"], case["name"]
        L.oracle_bitwriter_free(w)


def test_reader_vectors():
    L = o.lib()
    for case in V["reader"]:
        data = np.frombuffer(bytes.fromhex(case["bytes"]) + b"\0", dtype=np.uint8).copy()
        r = L.oracle_bitreader_new(data.ctypes.data, data.size - 1)
        assert L.oracle_bitreader_get_count(r) == 0
        for op, cnt in zip(case["ops"], case["counts"]):
            out = C.c_uint64()
            rc = L.oracle_bitreader_read_bits(r, op[1], C.byref(out))
            if op[2] == "eof":
                assert rc == o.EOF, case["name"]
            else:
                assert rc == o.OK and out.value == op[2], case["name"]
            assert L.oracle_bitreader_get_count(r) == cnt, case["name"]
        L.oracle_bitreader_free(r)


def test_write_rejects_oversized_symbol():
    """src/bitio/mod.rs:149-151."""
    L = o.lib()
    buf = np.zeros(4, dtype=np.uint8)
    w = L.oracle_bitwriter_new(buf.ctypes.data, buf.size)
    assert L.oracle_bitwriter_write_bits(w, 2, 1) == o.INVALID_INPUT
    assert L.oracle_bitwriter_write_bits(w, 256, 8) == o.INVALID_INPUT
    assert L.oracle_bitwriter_write_bits(w, 255, 8) == o.OK
    L.oracle_bitwriter_free(w)
