"""Attribute an ncu capture's per-SASS-instruction counts to CUDA source lines, using nvdisasm's line info of the
same kernel (no GPU needed).  Usage: python tools/sass_by_line.py <rep> <mangled kernel name> [top]"""
import collections, csv, io, re, subprocess, sys, os, tempfile
rep, fun = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(here, "deepjoin_b200/lib/libdeepjoin_b200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith("//---") and (".text." + fun + " ") in l)
lines = []   # (offset, file, line)
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
idx = {x: i for i, x in enumerate(rows[h])}
data = []
for r in rows[h + 1:]:
    if r and r[0] == "Address":
        break
    if len(r) == len(rows[h]):
        data.append(r)
print("nvdisasm instructions %d, ncu instructions %d" % (len(lines), len(data)))
base = int(data[0][idx["Address"]], 16) if data[0][idx["Address"]].startswith("0x") else int(data[0][idx["Address"]])
agg = collections.defaultdict(lambda: [0, 0])
n = min(len(lines), len(data))
for (off, cur, txt), r in zip(lines[:n], data[:n]):
    agg[cur][0] += int(r[idx["Instructions Executed"]])
    agg[cur][1] += int(r[idx["# Samples"]])
tot = sum(v[0] for v in agg.values())
print("total warp instructions", tot)
src_cache = {}
def src(f, l):
    for root in ("deepjoin_b200/csrc",):
        p = os.path.join(here, root, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            if 0 < l <= len(src_cache[p]):
                return src_cache[p][l - 1].strip()[:90]
    return ""
for (f, l), (ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% %9d inst %6d smp  %s:%d  %s" % (100.0 * ins / tot, ins, smp, f, l, src(f, l)))
