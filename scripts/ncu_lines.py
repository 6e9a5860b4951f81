"""Aggregate the warp-stall samples of an ncu report per CUDA source line.
usage: ncu_lines.py report.ncu-rep object.o mangled-kernel-substring [top]
(ncu's source page is per SASS address; nvdisasm --print-line-info of the same object maps addresses to lines)"""
import csv, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
amap, cur, infn = {}, None, False
for ln in dis:
    if ln.startswith(".text."):
        infn = kern in ln
    if not infn:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        amap[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]
si = H.index("# Samples")
stall = [i for i, h in enumerate(H) if h.startswith("stall_")]
base = None
agg, tot = {}, 0
for r in rows[hi + 1:]:
    if len(r) <= si or not r[0]:
        continue
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    if base is None:
        base = a
    n = int(r[si] or 0)
    tot += n
    key = amap.get(a - base, ((None, 0), ""))[0]
    d = agg.setdefault(key, [0, {}])
    d[0] += n
    for i in stall:
        v = int(r[i] or 0)
        if v:
            d[1][H[i][6:]] = d[1].get(H[i][6:], 0) + v
src = {}
print("total samples", tot)
for key, (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = key if key else (None, 0)
    text = ""
    if f:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
        if os.path.exists(p):
            src.setdefault(f, open(p).read().splitlines())
            text = src[f][l - 1].strip()[:90] if l - 1 < len(src[f]) else ""
    s3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% %6d %s:%d  %s   %s" % (100.0 * n / max(tot, 1), n, f, l, text, s3))
