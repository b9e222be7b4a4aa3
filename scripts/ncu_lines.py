"""Attribute an ncu capture's per-SASS-instruction stall samples / instruction counts to CUDA source lines.

usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel-substring> [top]
Needs the matching libzstdb200.so (built with -lineinfo); uses nvdisasm -g for the line table."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kname = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("ZB_LIB", os.path.join(root, "zstd_jni_b200/lib/libzstdb200.so"))], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
funcs = {}; cur = None; fil = None; line = None
for l in dis.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: cur = funcs.setdefault(m.group(1), []); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: fil = m.group(1).split("/")[-1]; line = int(m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", l)
    if m and cur is not None: cur.append((re.sub(r"\s+", " ", m.group(2)).strip(), fil, line))
cands = [k for k in funcs if kname in k and len(funcs[k]) > 50]
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.split("\n")))
# a report may hold several launches: one section per launch, each headed by a "Kernel Name" row; take the first that matches
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
pick = [i for i in starts if kname in rows[i][1]]
if pick:
    a = pick[0]; b = min([x for x in starts if x > a] + [len(rows)])
    rows = rows[a:b]
hdr = rows[1]; si = hdr.index("Warp Stall Sampling (All Samples)"); ii = hdr.index("Instructions Executed")
ncu = [(re.sub(r"\s+", " ", r[1]).strip(), int(r[si] or 0), int(r[ii] or 0)) for r in rows[2:] if len(r) > ii and r[si].isdigit()]
match = [k for k in cands if len(funcs[k]) == len(ncu)]
assert match, "no function named *%s* has %d instructions (have %s): rebuild the .so that was profiled" % (kname, len(ncu), [len(funcs[k]) for k in cands])
ins = funcs[match[0]]
agg = collections.Counter(); aggi = collections.Counter()
for (t, s, i), (t2, fl, ln) in zip(ncu, ins):
    agg[(fl, ln)] += s; aggi[(fl, ln)] += i
tot = sum(agg.values()); toti = sum(aggi.values())
print(f"total stall samples {tot}, warp instructions {toti}")
src = {}
def text(fl, ln):
    p = os.path.join(root, "zstd_jni_b200/csrc", fl or "")
    if fl not in src: src[fl] = open(p).read().split("\n") if os.path.exists(p) else []
    return src[fl][ln - 1].strip()[:110] if ln and ln <= len(src[fl]) else ""
for (fl, ln), v in agg.most_common(top):
    print(f"{v / tot * 100:5.1f}% stall {aggi[(fl, ln)] / toti * 100:5.1f}% inst  {fl}:{ln}  {text(fl, ln)}")
