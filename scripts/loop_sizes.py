#!/usr/bin/env python3
"""Static size of the two main loops (adaptive, frozen) of the hot kernels of a built library.
Usage: python scripts/loop_sizes.py [--so PATH] [--dump DIR]"""
import argparse, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser(); ap.add_argument("--so", default=os.path.join(ROOT, "redux_b200", "libredux_b200.so")); ap.add_argument("--dump")
a = ap.parse_args()
KERNS = (("decode_lane_al_kernelItLi0ELb1ELb0ELb1", "decode_narrow"), ("encode_lane_al_kernelItLi0ELb1ELb0", "encode_narrow"),
         ("decode_lane_al_kernelItLi3ELb0ELb1ELb0", "decode_wide_d_c32"), ("encode_lane_al_kernelItLi3ELb0ELb1", "encode_wide_d_c32"))
for kern, name in KERNS:
    lst = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_ctl.py"), kern, "--so", a.so], capture_output=True, text=True).stdout
    lines = [l for l in lst.splitlines() if l.startswith("/*")]
    addr = [int(l[2:l.index("*/")], 16) for l in lines]
    idx = {x: i for i, x in enumerate(addr)}
    lp = []
    for i, l in enumerate(lines):
        m = re.search(r"BRA(?:\.\w+)* (?:.*)?0x([0-9a-f]+)", l)
        if m and int(m.group(1), 16) in idx and idx[int(m.group(1), 16)] < i:
            lp.append((idx[int(m.group(1), 16)], i))
    big = sorted((b - x, x, b) for x, b in lp if b - x > 200)[-2:]
    print(name, "total", len(lines), "loops", [(n + 1, round((n + 1) / 4.0, 1)) for n, x, b in sorted(big, key=lambda t: t[1])])
    if a.dump:
        os.makedirs(a.dump, exist_ok=True)
        with open(os.path.join(a.dump, name + ".txt"), "w") as f:
            for n, x, b in sorted(big, key=lambda t: t[1]):
                f.write("\n// ---- loop of %d instructions\n" % (n + 1)); f.write("\n".join(lines[x:b + 1]) + "\n")
