"""Static view of register spills: LDL / STL instructions of one kernel per CUDA source line (nvdisasm line info)."""
import collections, glob, os, re, subprocess, sys
so, func = sys.argv[1], sys.argv[2]
pat = sys.argv[3] if len(sys.argv) > 3 else ""
for f in glob.glob("/tmp/spill_*.cubin"): os.remove(f)
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd="/tmp", capture_output=True)
cubins = [c for c in glob.glob("/tmp/*.cubin") if pat in c]
for cubin in cubins:
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    cur, infn = None, False
    agg = collections.Counter()
    for ln in dis.splitlines():
        m = re.match(r"^\.text\.(\S+):", ln)
        if m: infn = func in m.group(1); continue
        if not infn: continue
        m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', ln)
        if m: cur = (m.group(1), int(m.group(2))); continue
        if re.search(r"\b(LDL|STL)\b", ln): agg[(cur, "LDL" if "LDL" in ln else "STL")] += 1
    if agg:
        print(cubin, sum(agg.values()))
        for (loc, k), n in sorted(agg.items(), key=lambda kv: -kv[1])[:40]:
            print(f"  {n:4d} {k} {loc}")
