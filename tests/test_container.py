"""Blocked container: header handling on CPU; on the GPU, round trip, and every
block's bytes inside the container equal the oracle's compress() of that block alone."""
import io
import struct

import numpy as np
import pytest

import oracle_lib as o
import redux_b200 as rb
from redux_b200 import container as ct


def test_header_roundtrip_and_rejection():
    buf = io.BytesIO()
    ct.write_header(buf, rb.AdaptiveTreeModel(rb.Parameters(8, 22, 24)), 4096)
    model, block_len = ct.read_header(io.BytesIO(buf.getvalue()))
    assert (model.kind, model.params.freq_bits, model.params.code_bits, block_len) == (rb.MODEL_TREE, 22, 24, 4096)
    with pytest.raises(ct.ContainerError):
        ct.read_header(io.BytesIO(b"XXXX" + buf.getvalue()[4:]))
    with pytest.raises(ct.ContainerError):
        ct.read_header(io.BytesIO(buf.getvalue()[:7]))
    bad = bytearray(buf.getvalue()); bad[7] = 9          # freq_bits 9 < symbol_bits + 2
    with pytest.raises(rb.InvalidInput):
        ct.read_header(io.BytesIO(bytes(bad)))


@pytest.mark.gpu
def test_container_roundtrip_and_block_parity():
    params, L = (8, 14, 16), 5000
    data = rb.generate_blocks_host(0, 41, L, 0x5EED202610180000).tobytes()[: 40 * L + 1234]   # ragged tail
    model = rb.AdaptiveLinearModel(rb.Parameters(*params))
    out = io.BytesIO()
    raw_n, out_n = ct.pack_stream(io.BytesIO(data), out, model, block_len=L, batch_blocks=16)  # 3 segments
    blob = out.getvalue()
    assert raw_n == len(data) and out_n == len(blob)
    assert ct.unpack(blob) == data
    # walk the container by hand and compare each block's stream with the oracle
    pos, blk = ct.HEADER.size, 0
    while True:
        n, last = struct.unpack_from("<II", blob, pos); pos += 8
        if n == 0:
            break
        sizes = np.frombuffer(blob, dtype="<u4", count=n, offset=pos); pos += 4 * n
        for i in range(n):
            raw_block = data[blk * L:(blk + 1) * L]
            rc, want, _, _ = o.compress(raw_block, o.LINEAR, params)
            assert blob[pos:pos + int(sizes[i])] == want, blk
            pos += int(sizes[i]); blk += 1
    assert blk == 41 and pos == len(blob)
    with pytest.raises(rb.Eof):
        ct.unpack(blob[: len(blob) // 2])
    assert ct.unpack(ct.pack(b"", model)) == b""


def test_pack_refuses_what_the_header_cannot_describe():
    """A trained model or a symbol width other than 8 would write a container unpack_stream() cannot decode (the
    header holds kind + Parameters only; the index holds whole bytes): rejected when packing, before any device call."""
    trained = rb.AdaptiveTreeModel(rb.Parameters(8, 14, 16)).train([1, 2, 3])
    with pytest.raises(ct.ContainerError):
        ct.pack(b"abc", trained)
    with pytest.raises(ct.ContainerError):
        ct.pack(b"abc", rb.AdaptiveTreeModel(rb.Parameters(12, 14, 16)))
