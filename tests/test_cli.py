"""The CLI driver: argument handling mirrors src/main.rs:36-61 (CPU); the files it
writes equal the oracle's compress() with the CLI's fixed model (GPU)."""
import io

import pytest

import oracle_lib as o
import redux_b200 as rb
from redux_b200 import cli


def test_option_parsing_like_reference():
    assert cli.parse(["-c"]) == (True, None, None)
    assert cli.parse(["-d", "-i", "a", "-o", "b"]) == (False, "a", "b")
    assert cli.parse(["-c", "-d"]) == (False, None, None)          # last one wins, as in the reference loop
    assert cli.parse([]) is None and cli.parse(["-i", "a"]) is None
    assert cli.parse(["-c", "-i"]) is None and cli.parse(["-c", "-x"]) is None
    err = io.StringIO()
    assert cli.main([], stderr=err) == 1 and err.getvalue().startswith("Usage: redux (-c | -d)")
    err = io.StringIO()
    assert cli.main(["-c", "-i", "/nonexistent/input"], stderr=err) == 2
    assert err.getvalue().startswith("Error while opening input file")


@pytest.mark.gpu
def test_cli_files_equal_oracle(tmp_path):
    data = rb.generate_blocks_host(1, 1, 300000, 0x5EED202610180000).tobytes()      # text-like, one stream
    src, comp, back = tmp_path / "in.bin", tmp_path / "out.rdx", tmp_path / "back.bin"
    src.write_bytes(data)
    err = io.StringIO()
    assert cli.main(["-c", "-i", str(src), "-o", str(comp)], stderr=err) == 0
    rc, want, ic, oc = o.compress(data, o.TREE, (8, 30, 32))                       # src/main.rs:108
    assert comp.read_bytes() == want
    assert err.getvalue().strip() == "Compressed %d bytes into %d bytes, ratio: %.3f" % (ic, oc, ic / oc)
    err = io.StringIO()
    assert cli.main(["-d", "-i", str(comp), "-o", str(back)], stderr=err) == 0
    assert back.read_bytes() == data
    assert err.getvalue().strip() == "Decompressed %d bytes from %d bytes, ratio: %.3f" % (ic, oc, ic / oc)
    # stdin/stdout plumbing and the codec-error exit code
    out = io.BytesIO()
    assert cli.main(["-c"], stdin=io.BytesIO(b"redux"), stdout=out, stderr=io.StringIO()) == 0
    assert out.getvalue().hex() == "71f2a770a4a0f10a00"                           # SURVEY B.1 at (8,30,32)
    err = io.StringIO()
    assert cli.main(["-d"], stdin=io.BytesIO(want[: len(want) // 2]), stdout=io.BytesIO(), stderr=err) == 3
    assert err.getvalue().startswith("Decompression error: Unexpected end of file")
