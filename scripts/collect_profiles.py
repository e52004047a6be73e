#!/usr/bin/env python3
"""Turns the files a scripts/final_round.sh visit left in gpurun_out/ into the tracked evidence under profiles/.
Usage: scripts/collect_profiles.py <gpurun tag> <profiles prefix>      e.g.  r02_final2 r02"""
import json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, pre = sys.argv[1], sys.argv[2]
G = lambda name: os.path.join(ROOT, "gpurun_out", "%s_%s" % (tag, name))
P = lambda name: os.path.join(ROOT, "profiles", "%s_%s" % (pre, name))


def last_json(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


def run(*cmd):
    return subprocess.run(list(cmd), cwd=ROOT, capture_output=True, text=True)


# ---- ncu: narrow (headline) and wide captures
r = run(sys.executable, "scripts/summarize_ncu.py", G("prof.ncu-rep"), pre)
print(r.stdout[-600:], r.stderr[-300:])
bench = last_json(G("bench_default.json"))
traffic = json.load(open("/tmp/%s_traffic.json" % pre))
json.dump({"source": "profiles/%s_ncu_full_summary.md (ncu --set full --clock-control none, one capture per kernel)" % pre,
           "workload": bench["config"]["workload"], "kernels": traffic}, open(P("ncu_traffic.json"), "w"), indent=1)
r = run(sys.executable, "scripts/summarize_ncu.py", G("wide_prof.ncu-rep"), pre + "_wide_8_30_32")
print(r.stdout[-400:], r.stderr[-300:])
shutil.copy(G("launches.csv"), P("ncu_launches.csv"))
open(P("shared_memory_wavefronts.md"), "w").write(run(sys.executable, "scripts/ncu_shared_wavefronts.py", G("prof.ncu-rep")).stdout)

# ---- bench records
# (a scripts/final_short.sh visit leaves only the first two)
for src, dst in (("bench_default.json", "bench_default.json"), ("bench_reference.json", "bench_reference_arm.json"),
                 ("bench_huge.json", "bench_code_bits_34_16384_blocks.json")):
    if os.path.exists(G(src)):
        json.dump(last_json(G(src)), open(P(dst), "w"), indent=1)
for src, dst in (("generic.json", "generic_path.json"), ("underfilled.json", "underfilled.json"), ("small.json", "small_batches_corpora.json")):
    if os.path.exists(G(src)):
        shutil.copy(G(src), P(dst))
if os.path.exists(os.path.join(ROOT, "gpurun_out", "bench_alignment.json")):
    shutil.copy(os.path.join(ROOT, "gpurun_out", "bench_alignment.json"), P("alignment.json"))
shutil.copy(G("gpu.txt"), P("box.txt"))
open(P("pytest_gpu.log"), "w").write(open(G("pytest.log")).read()[-600:])

# ---- SASS of the hot loops with scheduling control fields (scripts/sass_ctl.py), built from the in-tree library
def loops(listing):
    lines = [l for l in listing.splitlines() if l.startswith("/*")]
    addr = [int(l[2:l.index("*/")], 16) for l in lines]
    idx = {a: i for i, a in enumerate(addr)}
    out = []
    for i, l in enumerate(lines):
        m = re.search(r"BRA(?:\.\w+)* (?:.*)?0x([0-9a-f]+)", l)
        if m and int(m.group(1), 16) in idx and idx[int(m.group(1), 16)] < i:
            out.append((idx[int(m.group(1), 16)], i))
    return lines, out


for kern, name in (("decode_lane_al_kernelItLi0ELb1ELb0ELb1", "decode_narrow"), ("encode_lane_al_kernelItLi0ELb1ELb0", "encode_narrow"),
                   ("decode_lane_al_kernelItLi3ELb0ELb1ELb0", "decode_wide_d_c32"), ("encode_lane_al_kernelItLi3ELb0ELb1", "encode_wide_d_c32")):
    lst = run(sys.executable, "scripts/sass_ctl.py", kern).stdout
    lines, lp = loops(lst)
    big = sorted((b - a, a, b) for a, b in lp if b - a > 200)[-2:]        # the adaptive and the frozen main loops
    with open(P("sass_%s_main_loops.txt" % name), "w") as f:
        f.write("// %s -- the two main loops (adaptive phase, frozen phase) of %s, four symbol steps per iteration.\n"
                "// cuobjdump -sass of redux_b200/libredux_b200.so through scripts/sass_ctl.py: wait[..] = scoreboards waited for,\n"
                "// W/R = scoreboard set on completion / on operand read, st = stall count, Y = yield.\n" % (lst.splitlines()[0], name))
        for n_, a, b in sorted(big, key=lambda x: x[1]):
            f.write("\n// ---- loop of %d instructions (%.1f per symbol step)\n" % (n_ + 1, (n_ + 1) / 4.0))
            f.write("\n".join(lines[a:b + 1]) + "\n")
    print(name, [(n_ + 1) for n_, a, b in big])
print("done")
