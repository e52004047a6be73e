#!/usr/bin/env python3
"""Generates tests/golden/*.json.  Run from the repo root:  python tests/golden/make_golden.py

The reference (peterbudai/redux) is a Rust crate and no Rust toolchain exists in the build
container, so golden COMPRESSED bytes cannot be produced by running the reference.  What this
script commits instead:

  kat_vectors.json    - small inputs -> compressed bytes, produced by `PyCoder` below: an
                        INDEPENDENT second reading of the reference semantics (naive per-symbol
                        frequency list + O(n) prefix sums instead of either reference table layout,
                        bits kept as a Python list, Python big ints).  The first 9 entries are the
                        hand-derived known answers of SURVEY.md Appendix B.1 and are asserted here.
                        tests/test_oracle_golden.py checks the C oracle against every entry, so a
                        slip in either reading shows up as a mismatch.
  bitio_vectors.json  - the reference's own byte-exact bit-packing vectors
                        (src/bitio/tests.rs:21-128 writer, :131-218 reader), transcribed as
                        operation lists.  These ARE reference goldens.
  corpus_table.json   - SURVEY.md Appendix B.2 (size + SHA-256 prefix of compress() of every corpus
                        file at three parameter triples), parsed from SURVEY.md.  Used only where
                        /root/reference/resources is present (CPU container), never on the GPU box.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


class PyCoder:
    """Second reading of src/codec.rs + src/model (observational semantics)."""

    def __init__(self, s, f, c):
        assert s >= 1 and f >= s + 2 and c >= f + 2 and c + f <= 64  # src/model/mod.rs:64
        self.s, self.f, self.c = s, f, c
        self.eof = 1 << s
        self.nsym = self.eof + 1
        self.fmax = (1 << f) - 1
        self.q = 1 << (c - 2)
        self.half = 2 * self.q
        self.q3 = 3 * self.q
        self.max = (1 << c) - 1
        self.freq = [1] * self.nsym
        self.total = self.nsym

    def _lookup_then_update(self, sym):
        lo = sum(self.freq[:sym])
        hi = lo + self.freq[sym]
        if self.total < self.fmax:  # freeze rule: adaptive_tree.rs:84 / adaptive_linear.rs:34
            self.freq[sym] += 1
            self.total += 1
        return lo, hi

    def encode(self, data):
        # symbols: MSB-first s-bit groups; trailing partial group is dropped (src/bitio/mod.rs:94-108)
        bits_in = []
        for b in data:
            bits_in.extend((b >> (7 - k)) & 1 for k in range(8))
        nsyms = len(bits_in) // self.s
        syms = [int("".join(map(str, bits_in[i * self.s:(i + 1) * self.s])), 2) for i in range(nsyms)]
        syms.append(self.eof)
        low, high, pending, extra = 0, self.max, 0, self.c
        out = []

        def put(bit):
            nonlocal pending
            out.append(bit)
            out.extend([1 - bit] * pending)
            pending = 0

        for sym in syms:
            count = self.total
            cl, ch = self._lookup_then_update(sym)
            rng = high - low + 1
            high = low + rng * ch // count - 1
            low = low + rng * cl // count
            while True:
                if high < self.half:
                    put(0)
                elif low >= self.half:
                    put(1)
                elif low >= self.q and high < self.q3:
                    pending += 1
                    low -= self.q
                    high -= self.q
                else:
                    break
                if sym == self.eof:
                    extra -= 1
                high = ((high << 1) + 1) & self.max
                low = (low << 1) & self.max
        while extra > 0:
            put(1 if low & self.half else 0)
            low = (low << 1) & self.max
            extra -= 1
        while len(out) % 8:
            out.append(0)
        return bytes(int("".join(map(str, out[i:i + 8])), 2) for i in range(0, len(out), 8))


def splitmix64(state):
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)


def gen(kind, n, seed):
    st = seed
    out = bytearray()
    while len(out) < n:
        st, r = splitmix64(st)
        if kind == "uniform":
            out.append(r >> 56)
        elif kind == "skew":  # geometric-ish
            v = 0
            x = r | (1 << 63)
            while not x & 1:
                v += 1
                x >>= 1
            out.append(min(255, v) + 97)
        elif kind == "runs":
            out.extend([(r >> 56) & 3] * (1 + (r & 31)))
        elif kind == "const":
            out.append(0x41)
    return bytes(out[:n])


B1 = [  # SURVEY.md Appendix B.1 (hand-derived known answers)
    (b"", (8, 14, 16), "ff00"), (b"", (8, 22, 24), "ff00ff"), (b"", (8, 30, 32), "ff00ff00"),
    (b"a", (8, 14, 16), "619d02"), (b"a", (8, 22, 24), "619d63f8"), (b"a", (8, 30, 32), "619d64970e"),
    (b"redux", (8, 14, 16), "71f23484c4c510"), (b"redux", (8, 22, 24), "71f2a6e1ec64a5fe"),
    (b"redux", (8, 30, 32), "71f2a770a4a0f10a00"),
]


def make_kat():
    vecs = []
    for data, p, hexout in B1:
        got = PyCoder(*p).encode(data).hex()
        assert got == hexout, (data, p, got, hexout)
        vecs.append({"name": "B1:%r" % data.decode(), "params": list(p), "input": data.hex(), "compressed": got,
                     "source": "SURVEY.md Appendix B.1 (hand-derived) == PyCoder"})
    cases = [
        ("uniform", 64, 1), ("uniform", 700, 2), ("skew", 300, 3), ("skew", 1500, 4), ("runs", 900, 5),
        ("const", 1200, 6), ("uniform", 1, 7), ("skew", 2, 8), ("runs", 3000, 9),
    ]
    triples = [(8, 10, 12), (8, 10, 16), (8, 14, 16), (8, 16, 18), (8, 22, 24), (8, 24, 30), (8, 30, 32),
               (8, 20, 44), (8, 30, 34)]
    for kind, n, seed in cases:
        data = gen(kind, n, 0x5EED0000 + seed)
        for p in triples:
            vecs.append({"name": "%s-%d" % (kind, n), "params": list(p), "input": data.hex(),
                         "compressed": PyCoder(*p).encode(data).hex(), "source": "PyCoder"})
    # non-byte symbol widths
    for s, f, c in [(4, 10, 16), (4, 14, 16), (12, 14, 16), (12, 24, 30), (3, 5, 7), (1, 3, 5)]:
        data = gen("uniform", 150, 0x5EED1000 + s)
        vecs.append({"name": "sym%d" % s, "params": [s, f, c], "input": data.hex(),
                     "compressed": PyCoder(s, f, c).encode(data).hex(), "source": "PyCoder"})
    # more widths (the device's generic path): lengths that leave a partial trailing symbol, frozen regimes
    for s, f, c in [(4, 10, 16), (12, 14, 16), (12, 30, 32), (5, 8, 11), (7, 20, 40), (16, 18, 20), (2, 4, 6), (9, 11, 13)]:
        for kind, n in (("skew", 301), ("runs", 1000), ("uniform", 7)):
            data = gen(kind, n, 0x5EED2000 + 16 * s + n)
            vecs.append({"name": "sym%d-%s-%d" % (s, kind, n), "params": [s, f, c], "input": data.hex(),
                         "compressed": PyCoder(s, f, c).encode(data).hex(), "source": "PyCoder"})
    # models trained before the call: get_frequency(sym) for sym in train, then compress (src/model/mod.rs:23-25)
    for (s, f, c), ntrain in [((8, 14, 16), 40), ((8, 10, 12), 2000), ((8, 30, 32), 500), ((4, 10, 16), 100),
                              ((12, 22, 24), 300)]:
        st = 0x5EED3000 + s * 1000 + ntrain
        train = []
        for _ in range(ntrain):
            st, r = splitmix64(st)
            train.append((r >> 40) % min((1 << s) + 1, 48))       # skewed towards small symbols, may include EOF-range
        for kind, n in (("skew", 400), ("runs", 700), ("uniform", 0)):
            data = gen(kind, n, 0x5EED4000 + s + n)
            coder = PyCoder(s, f, c)
            for sym in train:
                coder._lookup_then_update(sym)
            vecs.append({"name": "trained%d-%s-%d" % (ntrain, kind, n), "params": [s, f, c], "input": data.hex(),
                         "train": bytes(train).hex(), "compressed": coder.encode(data).hex(),
                         "source": "PyCoder, pre-trained (train = one byte per get_frequency() call)"})
    # adversarial: always code the symbol whose interval contains the midpoint, so that E3 shifts pile up into
    # pending runs far beyond 32 bits (src/codec.rs:75-83); then end with ordinary symbols.  Byte-aligned symbol
    # widths only (the input is built symbol by symbol).
    for (s, f, c), nstr in [((8, 14, 16), 12), ((8, 14, 16), 40), ((8, 22, 24), 25), ((8, 30, 32), 30), ((8, 10, 12), 60),
                            ((8, 30, 34), 20), ((4, 10, 16), 40), ((16, 18, 20), 8)]:
        syms, longest = straddle_symbols(s, f, c, nstr)
        tail = [3, 1, 4, 1, 5, 9, 2, 6]
        allsyms = syms + [t % (1 << s) for t in tail]
        bits = "".join(format(x, "0%db" % s) for x in allsyms)
        assert len(bits) % 8 == 0
        data = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
        assert longest > 32, (s, f, c, longest)
        vecs.append({"name": "straddle%d-pending%d" % (nstr, longest), "params": [s, f, c], "input": data.hex(),
                     "compressed": PyCoder(s, f, c).encode(data).hex(),
                     "source": "PyCoder, midpoint-straddling input (longest pending run %d)" % longest})
    with open(os.path.join(HERE, "kat_vectors.json"), "w") as fh:
        json.dump(vecs, fh, indent=0)
    print("kat_vectors.json:", len(vecs), "vectors")


def straddle_symbols(s, f, c, n):
    """n symbols, each chosen so that its interval contains the midpoint of the code range: the interval keeps
    straddling one half and every renormalisation is an E3 shift.  Returns (symbols, longest pending run)."""
    coder = PyCoder(s, f, c)
    low, high, pending, longest = 0, coder.max, 0, 0
    out = []
    for _ in range(n):
        count = coder.total
        rng = high - low + 1
        # the symbol with low' <= half - 1 and high' >= half, if there is one
        pick = None
        cum = 0
        for sym in range(coder.eof):          # data symbols only
            lo_, hi_ = cum, cum + coder.freq[sym]
            cum = hi_
            l2 = low + rng * lo_ // count
            h2 = low + rng * hi_ // count - 1
            if l2 < coder.half <= h2:
                pick = sym
                break
        if pick is None:
            pick = 0
        cl, ch = coder._lookup_then_update(pick)
        high = low + rng * ch // count - 1
        low = low + rng * cl // count
        while True:
            if high < coder.half or low >= coder.half:
                pending = 0
            elif low >= coder.q and high < coder.q3:
                pending += 1
                longest = max(longest, pending)
                low -= coder.q
                high -= coder.q
            else:
                break
            high = ((high << 1) + 1) & coder.max
            low = (low << 1) & coder.max
        out.append(pick)
    return out, longest


def make_bitio():
    # Transcription of src/bitio/tests.rs. op = ["w", symbol, bits] | ["f"]; "count" after each op.
    one = lambda bits: [["w", int(b), 1] for b in bits]
    writer = [
        {"name": "write_empty", "ref": "src/bitio/tests.rs:8-18", "ops": [["f"]], "counts": [0], "bytes": ""},
        {"name": "write_bytes", "ref": "src/bitio/tests.rs:20-34",
         "ops": [["w", 1, 8], ["w", 2, 8], ["w", 3, 8]], "counts": [1, 2, 3], "bytes": "010203"},
        {"name": "write_bits", "ref": "src/bitio/tests.rs:36-66",
         "ops": one("1010101000001111"), "counts": [0] * 7 + [1] * 8 + [2], "bytes": "aa0f"},
        {"name": "write_mixed", "ref": "src/bitio/tests.rs:68-102",
         "ops": one("10101010") + [["w", 0, 8]] + one("00001111") + [["w", 0xF0, 8]],
         "counts": [0] * 7 + [1] + [2] + [2] * 7 + [3] + [4], "bytes": "aa000ff0"},
        {"name": "write_flush", "ref": "src/bitio/tests.rs:104-128",
         "ops": [["f"]] + one("1010") + [["f"]] + one("0") + [["f"], ["f"]],
         "counts": [0, 0, 0, 0, 0, 1, 1, 2, 2], "bytes": "a000"},
    ]
    rd1 = lambda bits: [["r", 1, int(b)] for b in bits]
    reader = [  # op = ["r", bits, expected or "eof"]
        {"name": "read_eof", "ref": "src/bitio/tests.rs:130-141", "bytes": "",
         "ops": [["r", 1, "eof"], ["r", 8, "eof"], ["r", 1, "eof"], ["r", 8, "eof"]], "counts": [0, 0, 0, 0]},
        {"name": "read_bytes", "ref": "src/bitio/tests.rs:143-156", "bytes": "010203",
         "ops": [["r", 8, 1], ["r", 8, 2], ["r", 8, 3], ["r", 8, "eof"]], "counts": [1, 2, 3, 3]},
        {"name": "read_bits", "ref": "src/bitio/tests.rs:158-185", "bytes": "aa0f",
         "ops": rd1("1010101000001111") + [["r", 8, "eof"]], "counts": [1] * 8 + [2] * 8 + [2]},
        {"name": "read_mixed", "ref": "src/bitio/tests.rs:187-218", "bytes": "aa000ff0",
         "ops": rd1("10101010") + [["r", 8, 0]] + rd1("00001111") + [["r", 8, 0xF0], ["r", 8, "eof"]],
         "counts": [1] * 8 + [2] + [3] * 8 + [4, 4]},
    ]
    with open(os.path.join(HERE, "bitio_vectors.json"), "w") as fh:
        json.dump({"writer": writer, "reader": reader}, fh, indent=0)
    print("bitio_vectors.json written")


def make_corpus_table():
    rows = []
    pat = re.compile(r"^\| ([a-z]+/[\w.]+)(?: \([^)]*\))? \| (\d+) \| (\d+) · ([0-9a-f]{16}) \| (\d+) · ([0-9a-f]{16}) \| (\d+) · ([0-9a-f]{16}) \|")
    with open(os.path.join(ROOT, "SURVEY.md")) as fh:
        for line in fh:
            m = pat.match(line)
            if m:
                rows.append({"file": m.group(1), "raw": int(m.group(2)),
                             "8,14,16": [int(m.group(3)), m.group(4)],
                             "8,22,24": [int(m.group(5)), m.group(6)],
                             "8,30,32": [int(m.group(7)), m.group(8)]})
    assert len(rows) == 36, len(rows)
    with open(os.path.join(HERE, "corpus_table.json"), "w") as fh:
        json.dump(rows, fh, indent=0)
    print("corpus_table.json:", len(rows), "rows")


if __name__ == "__main__":
    make_kat()
    make_bitio()
    make_corpus_table()
    sys.exit(0)
