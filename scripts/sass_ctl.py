#!/usr/bin/env python3
"""sass_ctl.py -- SASS of one kernel of libredux_b200.so with the scheduling control fields decoded.

  python scripts/sass_ctl.py <mangled-name-regex> [--so PATH] [--grep REGEX] > listing.txt

`cuobjdump -sass` prints every sm_100 instruction as two 64-bit words; the upper word carries the
per-instruction control fields (same layout since Volta): stall count [41:44], yield [45], write
scoreboard [46:48], read scoreboard [49:51], wait mask [52:57], reuse [58:61].  A variable-latency
instruction (LDG, LDS, MUFU ...) names the scoreboard it will release (W=n); a later instruction waits on
the scoreboards of its wait mask.  The listing therefore shows WHERE a load's latency is paid -- which the
plain SASS text does not -- and is what the profiles/*_sass_*.txt extracts were made with.
"""
import argparse
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur is not None:
            out[cur].append(line)
    return out


def decode(lines):
    """[(addr, text, ctl dict)]"""
    res, pend = [], None
    for line in lines:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", line)
        if m:
            pend = (m.group(1), m.group(2).strip())
            continue
        m = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", line)
        if m and pend:
            hi = int(m.group(1), 16)
            ctl = {"stall": (hi >> 41) & 0xF, "yield": (hi >> 45) & 1, "wr": (hi >> 46) & 7, "rd": (hi >> 49) & 7,
                   "wait": (hi >> 52) & 0x3F, "reuse": (hi >> 58) & 0xF}
            res.append((pend[0], pend[1], ctl))
            pend = None
    return res


def fmt(ins):
    addr, text, c = ins
    wait = "".join(str(i) for i in range(6) if c["wait"] & (1 << i)) or "-"
    wr = str(c["wr"]) if c["wr"] != 7 else "-"
    rd = str(c["rd"]) if c["rd"] != 7 else "-"
    return "/*%s*/ wait[%-6s] W%s R%s st%-2d %s%s" % (addr, wait, wr, rd, c["stall"], "Y " if c["yield"] else "  ", text)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel")
    ap.add_argument("--so", default=os.path.join(ROOT, "redux_b200", "libredux_b200.so"))
    ap.add_argument("--grep", default=None, help="only print instructions matching this regex (plus 0 context)")
    ap.add_argument("--range", default=None, help="hex address range a:b")
    a = ap.parse_args()
    ks = kernels(a.so)
    names = [k for k in ks if re.search(a.kernel, k)]
    if len(names) != 1:
        sys.exit("kernel regex matches %d functions:\n  %s" % (len(names), "\n  ".join(names or ks)))
    ins = decode(ks[names[0]])
    print("// %s: %d instructions" % (names[0], len(ins)))
    lo, hi = (0, 1 << 62)
    if a.range:
        lo, hi = (int(x, 16) for x in a.range.split(":"))
    for i in ins:
        if not lo <= int(i[0], 16) <= hi:
            continue
        if a.grep and not re.search(a.grep, i[1]):
            continue
        print(fmt(i))
    return 0


if __name__ == "__main__":
    sys.exit(main())
