"""Aggregate an ncu SASS source page by CUDA source line.
usage: python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel substring> [top]"""
import csv, os, re, subprocess, sys, tempfile
rep, lib, kern = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "kernels.sm" in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# map instruction offset -> line within the kernel's function
line_of = {}
cur_line = None; in_fn = False
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        in_fn = kern in m.group(1); continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur_line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "Address" in r][0]
hdr = rows[hi]
ia, isamp, iins = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
base = None
agg = {}
tot_s = tot_i = 0
for r in rows[hi + 1:]:
    if len(r) <= iins: continue
    if r[ia] == "Address": break          # the page repeats itself per matching result: keep the first
    a = int(r[ia], 16)
    if base is None: base = a
    key = line_of.get(a - base)
    s = float(r[isamp] or 0); n = float(r[iins] or 0)
    tot_s += s; tot_i += n
    d = agg.setdefault(key, {"s": 0, "i": 0, "st": {}})
    d["s"] += s; d["i"] += n
    for c in stalls:
        v = float(r[hdr.index(c)] or 0)
        if v: d["st"][c] = d["st"].get(c, 0) + v
src = {}
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.3e}")
for key, d in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:top]:
    text = ""
    if key:
        path = [p for p in (os.path.join(os.path.dirname(lib), "csrc", key[0]),) if os.path.exists(p)]
        if path:
            if path[0] not in src: src[path[0]] = open(path[0]).read().splitlines()
            text = src[path[0]][key[1] - 1].strip()[:90]
    st = sorted(d["st"].items(), key=lambda kv: -kv[1])[:2]
    print(f"{d['s']/tot_s*100:5.1f}% smp {d['i']/tot_i*100:5.1f}% ins {str(key):28s} {' '.join(f'{k[6:]}={v:.0f}' for k,v in st):32s} | {text}")
