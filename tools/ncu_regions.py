"""Bucket an ncu SASS source page into regions of consecutive instructions labelled by their dominant .cu source line.
usage: python tools/ncu_regions.py <report.ncu-rep> <lib.so> <kernel substring> [chunk] [n_env]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib, kern = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2]), sys.argv[3]
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 50
n_env = int(sys.argv[5]) if len(sys.argv) > 5 else 65536
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "kernels.sm" in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of = {}; cur = None; in_fn = False
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: in_fn = kern in m.group(1); continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = os.path.basename(m.group(1))
        cur = (f, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "Address" in r]
hdr = rows[hi[0]]
ia, isamp, iins = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in (rows[hi[0] + 1:hi[1]] if len(hi) > 1 else rows[hi[0] + 1:]) if len(r) > iins]
base = int(body[0][ia], 16)
tot_s = sum(float(r[isamp]) for r in body); tot_i = sum(float(r[iins]) for r in body)
print(f"{kern}: {len(body)} SASS instructions, {tot_i / n_env:.0f} warp-instr/env, {tot_s:.0f} samples")
src = open(os.path.join(os.path.dirname(lib), "csrc", "avg_kernels.cu")).read().splitlines()
for c in range(0, len(body), chunk):
    part = body[c:c + chunk]
    s_ = sum(float(r[isamp]) for r in part); i_ = sum(float(r[iins]) for r in part)
    if s_ / tot_s < 0.004: continue
    cnt = collections.Counter()
    for r in part:
        k = line_of.get(int(r[ia], 16) - base)
        if k and k[0] == "avg_kernels.cu": cnt[k[1]] += float(r[isamp]) + 1
    top = cnt.most_common(1)[0][0] if cnt else 0
    print(f"{c:5d} {s_ / tot_s * 100:5.1f}% smp {i_ / n_env:7.1f} ins/env  L{top:<5d} {src[top - 1].strip()[:100] if top else ''}")
