#!/usr/bin/env python
"""Correlate an ncu report's per-SASS-instruction samples with CUDA source lines.
usage: ncu_lines.py REPORT.ncu-rep KERNEL_REGEX LIB.so [topN]
(ncu's CSV source page only exports the SASS view with metrics; line info comes from nvdisasm -g.)"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kre, lib = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel instance only
kname = rows[0][1]
hdr = rows[1]
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    body.append(dict(zip(hdr, r)))
base = int(body[0]["Address"], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
mangled = None
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur = None, ("?", 0)
    for ln in sass.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m and cur_fn:
            line_of.setdefault(cur_fn, {})[int(m.group(1), 16)] = (cur, m.group(2))
short = re.sub(r"\(.*", "", kname).split("::")[-1].split("<")[0]
cands = [fn for fn in line_of if short in fn]
best = max(cands, key=lambda fn: len(line_of[fn])) if cands else None
for fn in cands:
    if len(line_of[fn]) == len(body):
        best = fn
lines = line_of[best]
num = lambda x: float(x.replace(",", "")) if x not in ("", "-") else 0.0
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
tot_s = tot_i = 0.0
for d in body:
    off = int(d["Address"], 16) - base
    key = lines.get(off, (("?", 0), ""))[0]
    s, i, t = num(d["# Samples"]), num(d["Instructions Executed"]), num(d["Thread Instructions Executed"])
    agg[key][0] += s; agg[key][1] += i; agg[key][2] += t
    tot_s += s; tot_i += i
print(f"kernel {kname[:80]}  sass={len(body)} fn={best} samples={tot_s:.0f} warp-instr={tot_i:.0f}")
src_cache = {}
def src(fname, ln):
    for root in ("self6dpp_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, fname)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
    return ""
for key, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][1 if "--by-inst" in sys.argv else 0])[:topn]:
    print(f"{key[0]:20s}:{key[1]:4d} samp {100*s/max(tot_s,1):5.1f}% inst {100*i/max(tot_i,1):5.1f}% thr/inst {t/max(i,1):5.1f} | {src(*key)}")

# ---- per-phase view for the forward kernel: helper lines (inlined .cuh code) inherit the phase of the nearest
# preceding dibr_forward.cu line in address order
if "--phases" in sys.argv:
    # phases are declared in the source: a "// @phase NAME" comment opens a phase that lasts until the next marker
    srcfile = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "self6dpp_b200", "csrc", "dibr_forward.cu")
    marks = [(n + 1, l.split("@phase", 1)[1].strip()) for n, l in enumerate(open(srcfile).read().splitlines()) if "// @phase" in l]
    def phase_of(ln):
        name = "other"
        for a, n in marks:
            if a <= ln:
                name = n
        return name
    cur = "other"
    pagg = collections.defaultdict(lambda: [0.0, 0.0])
    for d in body:
        off = int(d["Address"], 16) - base
        (fname, ln), _ = lines.get(off, (("?", 0), ""))
        if fname == "dibr_forward.cu":
            cur = phase_of(ln)
        pagg[cur][0] += num(d["# Samples"]); pagg[cur][1] += num(d["Instructions Executed"])
    print("phase view:")
    for k, (s, i) in sorted(pagg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:18s} samples {100*s/max(tot_s,1):5.1f}%  warp-instr {100*i/max(tot_i,1):5.1f}%")
