"""Print the SASS instructions (with executed counts) whose inline path matches a given bucket substring."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kre, skip, mangled, want = sys.argv[1:6]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]; data = rows[2:]
iA, iE, iT = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = int(data[0][iA], 16); execd = {}
for r in data:
    try: execd[int(r[iA], 16) - base] = (int(r[iE]), int(r[iT]))
    except (ValueError, IndexError): pass
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("PTB_SO", os.path.join(ROOT, "project3-cuda-path-tracer_b200", "libptb200.so"))], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("ptb200.sm")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(mangled + ":")][0]
chain = []; pending = []
for l in dis[start + 1:]:
    if l.startswith("//---------------------"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: pending.append("%s:%s" % (os.path.basename(m.group(1)).replace("ptb_", "").replace(".cuh", ""), m.group(2))); continue
    if re.match(r"^\.L_", l): print(l); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m:
        if pending: chain = pending; pending = []
        off = int(m.group(1), 16)
        desc = " < ".join(chain)
        if want in desc and off in execd:
            e, t = execd[off]
            print("%6x %10d %5.1f  %-60s %s" % (off, e, t / max(e, 1), m.group(2)[:60], desc[:80]))
