"""Attribute ncu warp-stall samples of one kernel to CUDA source lines (SASS address -> line via nvdisasm)."""
import csv, re, subprocess, sys, collections
rep, so, func, topn = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows[:10]) if "Source" in r)
hdr = rows[h]; ci = {n: k for k, n in enumerate(hdr)}
samples = []
for r in rows[h + 1:]:
    if len(r) < len(hdr): continue
    try: samples.append((float(r[ci["# Samples"]] or 0), float(r[ci["Instructions Executed"]] or 0), float(r[ci["stall_long_sb"]] or 0), r[ci["Source"]]))
    except ValueError: pass
import os
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd="/tmp", capture_output=True)
import glob
pat = sys.argv[5] if len(sys.argv) > 5 else "uwbgo_kernels"
cubin = sorted(glob.glob("/tmp/*" + pat + "*.cubin"))[-1]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
cur, infn, lines = None, False, []
for ln in dis.splitlines():
    m = re.match(r"^\.text\.(\S+):", ln)
    if m: infn = func in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', ln)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln) and ";" in ln: lines.append(cur)
print("sass instrs:", len(samples), "disasm instrs:", len(lines))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for (s, ex, lsb, _), loc in zip(samples, lines):
    a = agg[loc]; a[0] += s; a[1] += ex; a[2] += lsb
tot = sum(a[0] for a in agg.values()) or 1
srcs = {}
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    text = ""
    if loc:
        f = "/root/repo/localization_b200/csrc/" + loc[0]
        if f not in srcs:
            try: srcs[f] = open(f).read().splitlines()
            except OSError: srcs[f] = []
        text = srcs[f][loc[1] - 1].strip()[:90] if loc[1] - 1 < len(srcs[f]) else ""
    print(f"{a[0]/tot*100:5.1f}%  lsb={a[2]/max(a[0],1)*100:3.0f}%  execs={a[1]:12.0f}  {loc}  {text}")
