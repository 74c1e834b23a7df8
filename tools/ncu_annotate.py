"""Per-instruction listing of one launch from an ncu report, each SASS instruction labelled with the device
function chain it was inlined from (nvdisasm -gi) -- the instruction-level view behind ncu_regions.py.
   python tools/ncu_annotate.py <ncu-rep> <kernel-regex> <launch-skip> <mangled-function-name> > listing.txt"""
import csv, os, re, subprocess, sys, tempfile
rep, kre, skip, mangled = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "project3-cuda-path-tracer_b200", "csrc")
tmp = tempfile.mkdtemp()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; data = rows[2:]
iA, iE, iT, iSm, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
base = int(data[0][iA], 16)
execd = {}
for r in data:
    try:
        execd[int(r[iA], 16) - base] = (int(r[iE]), int(r[iT]), int(r[iSm]), r[iS].strip())
    except (ValueError, IndexError):
        pass
funcs = {}
for f in ("ptb_math.cuh", "ptb_kernels.cuh"):
    lines = open(os.path.join(CSRC, f)).read().splitlines()
    starts = []
    for i, l in enumerate(lines):
        m = re.match(r"^(?:PTB_DEV|__global__|template.*__global__)?.*?\b([A-Za-z_0-9]+)\s*\([^;]*$", l)
        if (l.startswith("PTB_DEV") or l.startswith("__global__")) and m:
            starts.append((i + 1, m.group(1)))
    funcs[f] = starts
def func_of(f, ln):
    name = "?"
    for s, n in funcs.get(f, []):
        if s <= ln: name = n
        else: break
    return name
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("PTB_SO", os.path.join(ROOT, "project3-cuda-path-tracer_b200", "libptb200.so"))], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("ptb200.sm")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(mangled + ":")][0]
chain = []; pending = []
for l in dis[start + 1:]:
    if l.startswith("//---------------------"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m:
        if pending: chain = pending; pending = []
        off = int(m.group(1), 16)
        if off in execd:
            e, t, s, sass = execd[off]
            names = [func_of(f, ln) for f, ln in chain]
            label = "/".join(reversed(names[:3])) + ":%d" % chain[0][1]
            print("%05x %10d %5.1f %6d  %-44s %s" % (off, e, t / max(e, 1), s, label[-44:], sass))
