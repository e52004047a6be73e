"""Blocked container.  The reference's stream is headerless (src/lib.rs:102-120:
no magic, no parameters, no length), so a batch of independently coded blocks needs an index to be decodable.
This container is NEW (no reference counterpart) and lives strictly OUTSIDE the per-block bytes: every block's
stream inside it is byte-for-byte what redux::compress() writes for that block alone.

    file    := header segment*
    header  := "RDXB" u8 version(1) u8 model_kind u8 symbol_bits u8 freq_bits u8 code_bits u8[3] zero
               u32 block_len
    segment := u32 n_blocks  u32 last_block_raw_len  u32 comp_size[n_blocks]  stream[n_blocks]
               (blocks 0..n-2 hold block_len raw bytes; a segment with n_blocks == 0 ends the file)
All integers little-endian.  Segments let a file be written and read as a stream of batches."""
import io
import struct

import numpy as np

import redux_b200 as rb

MAGIC = b"RDXB"
HEADER = struct.Struct("<4sBBBBB3xI")


class ContainerError(rb.InvalidInput):
    pass


def write_header(f, model, block_len):
    p = model.params
    f.write(HEADER.pack(MAGIC, 1, model.kind, p.symbol_bits, p.freq_bits, p.code_bits, block_len))


def read_header(f):
    raw = f.read(HEADER.size)
    if len(raw) != HEADER.size:
        raise ContainerError("container header truncated")
    magic, ver, kind, s, fb, c, block_len = HEADER.unpack(raw)
    if magic != MAGIC or ver != 1 or kind not in (rb.MODEL_LINEAR, rb.MODEL_TREE) or block_len == 0:
        raise ContainerError("not an RDXB v1 container")
    cls = rb.AdaptiveTreeModel if kind == rb.MODEL_TREE else rb.AdaptiveLinearModel
    return cls(rb.Parameters(s, fb, c)), block_len        # Parameters() re-validates (src/model/mod.rs:64)


def pack_stream(fin, fout, model, block_len=65536, batch_blocks=16384, context=None):
    """Reads fin to the end, codes it in batches of `batch_blocks` blocks, writes the container to fout.
    Returns (raw_bytes, container_bytes)."""
    # The header records the model KIND and Parameters only, and unpack_stream() rebuilds a fresh model from it: a
    # model trained before the call would produce streams nothing in the file says how to decode.  Likewise the
    # segment index stores whole bytes per block: with symbol_bits != 8 the reference drops a trailing partial
    # symbol (src/bitio/mod.rs:94-108), so decoded lengths would not match.  Refuse both at pack time.
    if model.freq is not None:
        raise ContainerError("the container stores fresh models only (a trained model is not recorded in the header)")
    if model.params.symbol_bits != 8:
        raise ContainerError("the container stores byte-symbol streams only (symbol_bits must be 8)")
    ctx = context or rb._ctx()
    write_header(fout, model, block_len)
    raw_total, out_total = 0, HEADER.size
    while True:
        chunk = fin.read(block_len * batch_blocks)
        if not chunk:
            break
        data = np.frombuffer(chunk, dtype=np.uint8)
        n = (data.size + block_len - 1) // block_len
        off = np.minimum(np.arange(n + 1, dtype=np.uint64) * np.uint64(block_len), np.uint64(data.size))
        comp, coff, _ = ctx.encode_batch(data, off, model)
        sizes = (coff[1:] - coff[:-1]).astype("<u4")
        fout.write(struct.pack("<II", n, data.size - (n - 1) * block_len))
        fout.write(sizes.tobytes())
        fout.write(comp.tobytes())
        raw_total += data.size
        out_total += 8 + sizes.nbytes + comp.size
    fout.write(struct.pack("<II", 0, 0))
    return raw_total, out_total + 8


def unpack_stream(fin, fout, context=None):
    """Inverse of pack_stream. Returns (container_bytes, raw_bytes). Raises Eof on a truncated file."""
    ctx = context or rb._ctx()
    model, block_len = read_header(fin)
    consumed, raw_total = HEADER.size, 0
    while True:
        head = fin.read(8)
        if len(head) != 8:
            raise rb.Eof("container ended inside a segment header")
        n, last = struct.unpack("<II", head)
        consumed += 8
        if n == 0:
            return consumed, raw_total
        if last == 0 or last > block_len:
            raise ContainerError("bad segment")
        raw_sizes = fin.read(4 * n)
        if len(raw_sizes) != 4 * n:
            raise rb.Eof("container ended inside a segment index")
        sizes = np.frombuffer(raw_sizes, dtype="<u4").astype(np.uint64)
        coff = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(sizes, out=coff[1:])
        blob = fin.read(int(coff[-1]))
        if len(blob) != int(coff[-1]):
            raise rb.Eof("container ended inside a segment's streams")
        lens = np.full(n, block_len, dtype=np.uint64)
        lens[-1] = last
        roff = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(lens, out=roff[1:])
        raw, raw_lens, used, status = ctx.decode_batch(np.frombuffer(blob, dtype=np.uint8), coff, roff, model)
        if not (raw_lens == lens).all() or not (used == sizes).all():
            raise ContainerError("a block decoded to an unexpected length")
        fout.write(raw[: int(roff[-1])].tobytes())
        consumed += 4 * n + len(blob)
        raw_total += int(roff[-1])


def pack(data, model, block_len=65536, context=None):
    out = io.BytesIO()
    pack_stream(io.BytesIO(bytes(data)), out, model, block_len, context=context)
    return out.getvalue()


def unpack(blob, context=None):
    out = io.BytesIO()
    unpack_stream(io.BytesIO(bytes(blob)), out, context=context)
    return out.getvalue()
