"""Basic-block view of one kernel from an ncu report: consecutive SASS instructions with the same execution count, with the inline chain
(function names, outermost first) and source-line span of each block.
   python tools/ncu_blocks.py <ncu-rep> <kernel-regex> <launch-skip> <mangled-function-name> [min-percent]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kre, skip, mangled = sys.argv[1:5]
minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "project3-cuda-path-tracer_b200", "csrc")
tmp = tempfile.mkdtemp()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; data = rows[2:]
iA, iE, iT, iSm = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = int(data[0][iA], 16)
execd = {}
for r in data:
    try: execd[int(r[iA], 16) - base] = (int(r[iE]), int(r[iT]), int(r[iSm]))
    except (ValueError, IndexError): pass
funcs = {}
for f in ("ptb_math.cuh", "ptb_kernels.cuh", "ptb_trace.cuh"):
    lines = open(os.path.join(CSRC, f)).read().splitlines()
    starts = []
    for i, l in enumerate(lines):
        m = re.match(r"^(?:PTB_DEV|__global__|template.*__global__)?.*?\b([A-Za-z_0-9]+)\s*\([^;]*$", l)
        if (l.startswith("PTB_DEV") or l.startswith("__global__")) and m: starts.append((i + 1, m.group(1)))
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
LEAF = {"mul", "add", "sub", "fdiv", "fsqrt", "mk3", "add3", "sub3", "mul3", "scale3", "neg3", "dot3", "cross3", "normalize3", "length3", "mvRow", "mulMV",
        "rcpApprox", "rsqrtApprox", "rcpApproxRaw", "rsqrtApproxRaw", "rcpRefined", "divRefined", "sqrtRefined", "moderate", "mvRowPoint", "mulMVPoint", "pk2", "bc2", "upk2", "add2", "sub2", "mul2", "?"}
chain = []; pending = []
blocks = []   # [count, n, thr, samples, Counter(keys), set(lines), first opcode list]
total = 0
for l in dis[start + 1:]:
    if l.startswith("//---------------------"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if not m: continue
    if pending: chain = pending; pending = []
    off = int(m.group(1), 16)
    if off not in execd: continue
    e, t, s = execd[off]
    names = [func_of(f, ln) for f, ln in chain]
    path = [n for n in reversed(names) if n not in LEAF]
    key = "/".join(path[1:5]) if len(path) > 1 else (path[0] if path else "other")
    outer = [(f, ln) for f, ln in chain if func_of(f, ln) not in LEAF]
    ln = outer[0] if outer else (chain[0] if chain else ("?", 0))
    if not blocks or blocks[-1][0] != e:
        blocks.append([e, 0, 0, 0, collections.Counter(), [], off])
    b = blocks[-1]; b[1] += 1; b[2] += t; b[3] += s; b[4][key] += 1; b[5].append("%s:%d" % (ln[0].replace("ptb_", "").replace(".cuh", ""), ln[1]))
    total += e
print("total warp instructions", total)
for e, n, t, s, keys, lines, off in blocks:
    if e * n < total * minpct / 100.0: continue
    ls = sorted(set(lines), key=lambda x: (x.split(":")[0], int(x.split(":")[1])))
    print("%5.2f%%  @%05x  %4d instr x %10d  lanes %4.1f  samples %6d  %s  [%s .. %s]" % (100.0 * e * n / total, off, n, e, t / max(1, e * n), s,
          ", ".join("%s(%d)" % kv for kv in keys.most_common(3)), ls[0], ls[-1]))
