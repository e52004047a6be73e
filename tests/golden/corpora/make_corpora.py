#!/usr/bin/env python3
"""make_corpora.py -- packs the reference's test corpora into ONE compressed fixture for the GPU parity tests.

  python tests/golden/corpora/make_corpora.py        (CPU container only: reads /root/reference/resources)

The reference's integration suite (tests/corpora.rs:87-259) round-trips every file under resources/*; the GPU box
has no /root/reference, so the same bytes travel as tests/golden/corpora/corpora.tar.xz (test data, not source:
the Calgary, Canterbury, Large, Artificial and Misc corpora as they lie in the reference tree, unmodified, 36
files, 13,796,125 bytes raw).  MANIFEST.json beside it lists path, size and SHA-256 of every member so that the
fixture can be checked without the reference (tests/test_corpora_fixture.py).  resources/large/E.coli is absent
from the reference tree (.MISSING_LARGE_BLOBS:1); its stand-in is generated, not stored (tests/corpora_fixture.py).
"""
import hashlib
import io
import json
import os
import tarfile

SRC = "/root/reference/resources"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    members = []
    for corpus in sorted(os.listdir(SRC)):
        d = os.path.join(SRC, corpus)
        if not os.path.isdir(d):
            continue
        for name in sorted(os.listdir(d)):
            members.append("%s/%s" % (corpus, name))
    buf = io.BytesIO()
    manifest = []
    with tarfile.open(fileobj=buf, mode="w:xz", preset=9) as tar:
        for m in members:
            data = open(os.path.join(SRC, m), "rb").read()
            ti = tarfile.TarInfo(m)
            ti.size = len(data)
            ti.mtime = 0
            ti.mode = 0o644
            tar.addfile(ti, io.BytesIO(data))
            manifest.append({"file": m, "raw": len(data), "sha256": hashlib.sha256(data).hexdigest()})
    open(os.path.join(HERE, "corpora.tar.xz"), "wb").write(buf.getvalue())
    json.dump(manifest, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1)
    print("%d files, %d bytes raw -> %d bytes" % (len(manifest), sum(r["raw"] for r in manifest), len(buf.getvalue())))


if __name__ == "__main__":
    main()
