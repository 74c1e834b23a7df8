"""Attribute executed SASS instructions of one kernel (ncu --page source --csv, SASS view) to CUDA
source lines using nvdisasm -g line annotations of the cubin.  Usage:
   python tools/ncu_lines.py <ncu-rep> <kernel-regex> <launch-skip> <mangled-function-name>"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kre, skip, mangled = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; data = rows[2:]
iA, iS, iE, iT, iSm = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = int(data[0][iA], 16)
execd = {}
for r in data:
    try:
        execd[int(r[iA], 16) - base] = (int(r[iE]), int(r[iT]), int(r[iSm]), r[iS].strip())
    except (ValueError, IndexError):
        pass
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("PTB_SO", os.path.join(ROOT, "project3-cuda-path-tracer_b200", "libptb200.so"))], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("ptb200.sm")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(mangled + ":")][0]
cur = ("?", 0)
byline = collections.Counter(); bylineT = collections.Counter(); bysm = collections.Counter()
total = 0
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m:
        off = int(m.group(1), 16)
        if off in execd:
            e, t, s, _ = execd[off]
            byline[cur] += e; bylineT[cur] += t; bysm[cur] += s; total += e
print("total warp instructions", total)
srcs = {}
for (f, ln), n in byline.most_common(45):
    if f not in srcs:
        for d in ("project3-cuda-path-tracer_b200/csrc", "/usr/local/cuda/include", "/usr/local/cuda/include/crt"):
            p = os.path.join(ROOT, d, f) if not d.startswith("/") else os.path.join(d, f)
            if os.path.exists(p):
                srcs[f] = open(p, errors="ignore").read().splitlines(); break
        else:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print("%5.1f%%  thr/inst %4.1f  samples %5d  %s:%d  %s" % (100.0 * n / total, bylineT[(f, ln)] / max(n, 1), bysm[(f, ln)], f, ln, text))
