"""corpora_fixture.py -- the reference's test corpora for tests that run where /root/reference does not exist
(the GPU box).  TEST INFRASTRUCTURE.

`corpora()` unpacks tests/golden/corpora/corpora.tar.xz (made by tests/golden/corpora/make_corpora.py from
/root/reference/resources, 36 files, byte-identical) once per process and returns {"calgary/book1": bytes, ...}.
`ecoli_stand_in()` is the seeded 4-symbol replacement for resources/large/E.coli, which the reference tree lacks
(/root/reference/.MISSING_LARGE_BLOBS:1): 4,638,690 symbols of "acgt" (the canonical file's length and alphabet),
symbol = top two bits of splitmix64 with state 0x5EED202610180005 + i."""
import io
import json
import lzma
import os
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "golden", "corpora", "corpora.tar.xz")
MANIFEST = os.path.join(HERE, "golden", "corpora", "MANIFEST.json")
TABLE = os.path.join(HERE, "golden", "corpus_table.json")
ECOLI_LEN = 4638690
ECOLI_SEED = 0x5EED202610180005
_cache = None


def manifest():
    return json.load(open(MANIFEST))


def corpus_table():
    """per file, raw size and (compressed size, first 16 hex of SHA-256) per parameter triple."""
    return json.load(open(TABLE))


def corpora():
    global _cache
    if _cache is None:
        raw = lzma.decompress(open(ARCHIVE, "rb").read())
        out = {}
        with tarfile.open(fileobj=io.BytesIO(raw), mode="r:") as tar:
            for m in tar.getmembers():
                out[m.name] = tar.extractfile(m).read()
        _cache = out
    return _cache


def splitmix64(state):
    """numpy uint64 vector of splitmix64 outputs for the given states (state += GAMMA already applied by caller)."""
    z = state
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def ecoli_stand_in(n=ECOLI_LEN, seed=ECOLI_SEED):
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        z = splitmix64(np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15))
    return np.frombuffer(b"acgt", dtype=np.uint8)[(z >> np.uint64(62)).astype(np.intp)].tobytes()


def blocks_of(data, block_len):
    """1 MiB blocks as config 5 cuts them: whole blocks, the last one short."""
    return [data[i:i + block_len] for i in range(0, len(data), block_len)]
